// TEST INFRASTRUCTURE (oracle build only): the subset of google-glog the
// reference uses (LOG, CHECK*, CHECK_NEAR, CHECK_NOTNULL, FLAGS_logtostderr,
// InitGoogleLogging, InstallFailureSignalHandler).  Logging only.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace google {
enum LogSeverity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };

class LogMessage {
 public:
  LogMessage(const char* file, int line, int severity) : severity_(severity) {
    static const char kTag[] = {'I', 'W', 'E', 'F'};
    const char* base = file;
    for (const char* p = file; *p; ++p) {
      if (*p == '/') base = p + 1;
    }
    stream_ << kTag[severity] << ' ' << base << ':' << line << "] ";
  }
  ~LogMessage() {
    stream_ << '\n';
    std::cerr << stream_.str();
    std::cerr.flush();
    if (severity_ == GLOG_FATAL) std::abort();
  }
  std::ostream& stream() { return stream_; }
 private:
  int severity_;
  std::ostringstream stream_;
};

// for `cond ? (void)0 : Voidify() & stream`
struct LogMessageVoidify {
  void operator&(std::ostream&) {}
};

inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}

template <typename T>
T CheckNotNull(const char* file, int line, const char* what, T&& t) {
  if (t == nullptr) {
    LogMessage(file, line, GLOG_FATAL).stream() << what;
  }
  return std::forward<T>(t);
}
}  // namespace google

static bool FLAGS_logtostderr __attribute__((unused)) = true;
static int FLAGS_minloglevel __attribute__((unused)) = 0;
static int FLAGS_v __attribute__((unused)) = 0;

#define POVAR_GLOG_SEV_INFO ::google::GLOG_INFO
#define POVAR_GLOG_SEV_WARNING ::google::GLOG_WARNING
#define POVAR_GLOG_SEV_ERROR ::google::GLOG_ERROR
#define POVAR_GLOG_SEV_FATAL ::google::GLOG_FATAL

#define LOG(sev) \
  ::google::LogMessage(__FILE__, __LINE__, POVAR_GLOG_SEV_##sev).stream()

#define LOG_IF(sev, cond) \
  !(cond) ? (void)0 : ::google::LogMessageVoidify() & LOG(sev)

#define CHECK(cond)                                      \
  (cond) ? (void)0                                       \
         : ::google::LogMessageVoidify() &               \
               LOG(FATAL) << "Check failed: " #cond " "

#define POVAR_CHECK_OP(a, b, op)                                          \
  ((a)op(b)) ? (void)0                                                    \
             : ::google::LogMessageVoidify() &                            \
                   LOG(FATAL) << "Check failed: " #a " " #op " " #b " (" \
                              << (a) << " vs. " << (b) << ") "

#define CHECK_EQ(a, b) POVAR_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) POVAR_CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) POVAR_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) POVAR_CHECK_OP(a, b, <)
#define CHECK_GE(a, b) POVAR_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) POVAR_CHECK_OP(a, b, >)

#define CHECK_NEAR(a, b, margin)                                    \
  (std::abs((a) - (b)) <= (margin))                                 \
      ? (void)0                                                     \
      : ::google::LogMessageVoidify() &                             \
            LOG(FATAL) << "Check failed: |" #a " - " #b "| <= " #margin " "

#define CHECK_NOTNULL(p) \
  ::google::CheckNotNull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))

#define DCHECK(cond) CHECK(cond)
#define VLOG(n) LOG_IF(INFO, false)
