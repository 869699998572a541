// Host-side problem plumbing behind the C ABI: BAL text reader, canonical ordering, sharding.
// No CUDA in this file.
//
// Replaces BalProblem::load_bal_eccv (/root/reference/src/rootba_povar/bal/bal_problem.cpp:182-303)
// for the 15-parameter files written by --create-dataset (bal_problem.cpp:306-471):
//   C L N / N x (cam lm x y) / C x 15 / L x 3
// The reference keeps a landmark's observations in a std::map<cam, obs> (bal_problem.hpp:226), so
// its canonical order is landmark index, then camera index ascending, whatever the file order;
// the loader flips the image y axis (bal_problem.cpp:240) and treats a repeated (cam, lm) pair as
// fatal (bal_problem.cpp:227).  The random landmark draw of the loader (bal_problem.cpp:255-268) is
// not reproduced: iteration 0 of step 1 overwrites all landmarks.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <random>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "../../../include/povar_b200.h"

namespace {

void set_err(char* err, size_t len, const std::string& msg) {
  if (err && len > 0) {
    std::snprintf(err, len, "%s", msg.c_str());
  }
}

// minimal fast tokenizer over a whole-file buffer
struct Cursor {
  const char* p;
  const char* end;
  void skip_ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool next_int(long long* out) {
    skip_ws();
    if (p >= end) return false;
    char* q = nullptr;
    errno = 0;
    const long long v = std::strtoll(p, &q, 10);
    if (q == p || errno != 0) return false;
    p = q;
    *out = v;
    return true;
  }
  bool next_double(double* out) {
    skip_ws();
    if (p >= end) return false;
    char* q = nullptr;
    errno = 0;
    const double v = std::strtod(p, &q);
    if (q == p) return false;
    p = q;
    *out = v;
    return true;
  }
};

}  // namespace

extern "C" int povar_canonical_order(int32_t num_cams, int32_t num_lms, int64_t num_obs,
                                     const int32_t* cam, const int32_t* lm, int64_t* perm,
                                     int64_t* lm_ptr) {
  if (num_cams <= 0 || num_lms < 0 || num_obs < 0 || !perm || !lm_ptr) return POVAR_ERR_INVALID;
  if (num_obs > 0 && (!cam || !lm)) return POVAR_ERR_INVALID;
  // counting sort by landmark (stable), then sort each landmark's run by camera
  std::vector<int64_t> count(static_cast<size_t>(num_lms) + 1, 0);
  for (int64_t i = 0; i < num_obs; ++i) {
    if (lm[i] < 0 || lm[i] >= num_lms || cam[i] < 0 || cam[i] >= num_cams) return POVAR_ERR_INVALID;
    count[static_cast<size_t>(lm[i]) + 1]++;
  }
  for (int32_t l = 0; l < num_lms; ++l) count[l + 1] += count[l];
  for (int32_t l = 0; l <= num_lms; ++l) lm_ptr[l] = count[l];
  std::vector<int64_t> fill(count.begin(), count.end() - 1);
  for (int64_t i = 0; i < num_obs; ++i) perm[fill[lm[i]]++] = i;
  for (int32_t l = 0; l < num_lms; ++l) {
    int64_t* b = perm + lm_ptr[l];
    int64_t* e = perm + lm_ptr[l + 1];
    std::sort(b, e, [cam](int64_t a, int64_t c) { return cam[a] < cam[c]; });
    for (int64_t* q = b + 1; q < e; ++q) {
      if (cam[*q] == cam[*(q - 1)]) return POVAR_ERR_INVALID;  // duplicate observation
    }
  }
  return POVAR_OK;
}

extern "C" int povar_partition_landmarks(int32_t num_lms, const int64_t* lm_ptr, int32_t world_size,
                                         int32_t* bounds) {
  if (num_lms < 0 || world_size <= 0 || !lm_ptr || !bounds) return POVAR_ERR_INVALID;
  // contiguous landmark ranges with (nearly) equal observation counts: rank r ends at the first
  // landmark whose cumulative observation count reaches (r+1)/world of the total
  const int64_t total = lm_ptr[num_lms];
  bounds[0] = 0;
  for (int32_t r = 1; r < world_size; ++r) {
    const int64_t target = (total * r + world_size / 2) / world_size;
    // first landmark boundary at or after the target, or the one before it if that is closer
    int64_t l = std::lower_bound(lm_ptr, lm_ptr + num_lms + 1, target) - lm_ptr;
    if (l > num_lms) l = num_lms;
    if (l > 0 && (target - lm_ptr[l - 1]) <= (lm_ptr[l] - target)) --l;
    if (l < bounds[r - 1]) l = bounds[r - 1];
    bounds[r] = static_cast<int32_t>(l);
  }
  bounds[world_size] = num_lms;
  return POVAR_OK;
}

extern "C" void povar_bal_free(povar_bal_data* data) {
  if (!data) return;
  std::free(data->lm_ptr);
  std::free(data->obs_cam);
  std::free(data->obs_uv);
  std::free(data->cam_params);
  std::memset(data, 0, sizeof(*data));
}

// Number parsing: std::from_chars is correctly rounded like strtod (the values must be the ones the
// reference's fscanf("%lf") reads, bit for bit) and several times faster; anything it does not accept
// (a leading '+', "inf", hex floats) goes through strtod.
static inline bool is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

static bool token_to_double(const char* b, const char* e, double* out) {
  const auto r = std::from_chars(b, e, *out);
  if (r.ec == std::errc() && r.ptr == e) return true;
  std::string tmp(b, e);
  char* q = nullptr;
  errno = 0;
  const double v = std::strtod(tmp.c_str(), &q);
  if (q == tmp.c_str() || *q != '\0') return false;
  *out = v;
  return true;
}

static bool token_to_int(const char* b, const char* e, long long* out) {
  if (b < e && *b == '+') ++b;
  const auto r = std::from_chars(b, e, *out);
  return r.ec == std::errc() && r.ptr == e;
}

// whole file into memory (+ a terminating NUL); false on any I/O error (ftell can fail on pipes, fread can be short)
static bool slurp(FILE* f, std::vector<char>* buf) {
  if (std::fseek(f, 0, SEEK_END) != 0) return false;
  const long size = std::ftell(f);
  if (size < 0 || std::fseek(f, 0, SEEK_SET) != 0) return false;
  buf->resize(static_cast<size_t>(size) + 1);
  const size_t got = std::fread(buf->data(), 1, static_cast<size_t>(size), f);
  if (got != static_cast<size_t>(size) || std::ferror(f)) return false;
  (*buf)[got] = '\0';
  return true;
}

static int reader_threads(size_t bytes) {
  static const int forced = getenv("POVAR_HOST_THREADS") ? atoi(getenv("POVAR_HOST_THREADS")) : 0;
  if (forced > 0) return forced;
  if (bytes < (1u << 20)) return 1;
  const unsigned hw = std::thread::hardware_concurrency();
  return static_cast<int>(std::min<unsigned>(hw == 0 ? 1 : hw, 16));
}

// The file is a flat list of whitespace-separated tokens: 3 header, 4 N observation, 15 C camera, 3 L landmark.
// Two parallel passes over chunks cut at whitespace: count the tokens of each chunk, then (their global numbers
// known by a prefix sum) convert every token straight into its destination.
static int bal_read_impl(const char* path, povar_bal_data* out, char* err, size_t err_len) {
  if (!path || !out) return POVAR_ERR_INVALID;
  std::memset(out, 0, sizeof(*out));
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    set_err(err, err_len, std::string("Could not open '") + path + "'");
    return POVAR_ERR_IO;
  }
  std::vector<char> buf;
  if (!slurp(f, &buf)) {
    std::fclose(f);
    set_err(err, err_len, std::string("Could not read '") + path + "'");
    return POVAR_ERR_IO;
  }
  std::fclose(f);
  const size_t got = buf.size() - 1;
  const char* base = buf.data();
  const char* end = base + got;
  Cursor cur{base, end};

  long long C = 0, L = 0, N = 0;
  if (!cur.next_int(&C) || !cur.next_int(&L) || !cur.next_int(&N) || C <= 0 || L <= 0 || N <= 0) {
    set_err(err, err_len, std::string("Failed to parse header of '") + path + "'");
    return POVAR_ERR_IO;
  }
  // the header is untrusted: indices are int32 on the device, and a file with N observations, C cameras and L
  // landmarks has at least 4 N + 15 C + 3 L tokens of at least two bytes each -- checked before anything is
  // allocated from these numbers
  if (C > INT32_MAX || L > INT32_MAX || N > INT32_MAX ||
      2 * (4 * N + 15 * C + 3 * L) > static_cast<long long>(got) + 1) {
    set_err(err, err_len, std::string("Header of '") + path + "' does not fit the file (" + std::to_string(C) +
                              " cameras, " + std::to_string(L) + " landmarks, " + std::to_string(N) +
                              " observations in " + std::to_string(got) + " bytes)");
    return POVAR_ERR_IO;
  }
  const long long n_obs_tok = 4 * N, n_cam_tok = 15 * C, n_lm_tok = 3 * L;
  std::vector<int32_t> cam(static_cast<size_t>(N)), lm(static_cast<size_t>(N));
  std::vector<double> xy(2 * static_cast<size_t>(N));
  out->cam_params = static_cast<double*>(std::malloc(sizeof(double) * 15 * static_cast<size_t>(C)));

  // chunks of the remainder, each starting at the beginning of a token
  const char* body = cur.p;
  const int T = reader_threads(static_cast<size_t>(end - body));
  std::vector<const char*> cut(T + 1);
  cut[0] = body;
  cut[T] = end;
  for (int t = 1; t < T; ++t) {
    const char* p = body + (end - body) * t / T;
    while (p < end && !is_ws(*p)) ++p;   // finish the token the cut fell into
    cut[t] = p;
  }
  std::vector<long long> first(T + 1, 0);
  std::vector<std::thread> pool;
  auto run = [&](auto&& fn) {
    pool.clear();
    for (int t = 1; t < T; ++t) pool.emplace_back(fn, t);
    fn(0);
    for (auto& th : pool) th.join();
  };
  run([&](int t) {
    long long n = 0;
    const char* p = cut[t];
    const char* e = cut[t + 1];
    while (p < e) {
      while (p < e && is_ws(*p)) ++p;
      if (p >= e) break;
      ++n;
      while (p < e && !is_ws(*p)) ++p;
    }
    first[t + 1] = n;
  });
  for (int t = 0; t < T; ++t) first[t + 1] += first[t];
  const long long total = first[T];
  // 0 ok, 1 bad observation token, 2 index out of range, 3 bad camera token, 4 bad landmark token
  std::vector<int> status(T, 0);
  run([&](int t) {
    long long k = first[t];
    const char* p = cut[t];
    const char* e = cut[t + 1];
    while (p < e) {
      while (p < e && is_ws(*p)) ++p;
      if (p >= e) break;
      const char* b = p;
      while (p < e && !is_ws(*p)) ++p;
      if (k < n_obs_tok) {
        const long long i = k >> 2;
        const int field = static_cast<int>(k & 3);
        if (field < 2) {
          long long v = 0;
          if (!token_to_int(b, p, &v)) {
            status[t] = 1;
            return;
          }
          if (v < 0 || v >= (field == 0 ? C : L)) {
            status[t] = 2;
            return;
          }
          (field == 0 ? cam : lm)[i] = static_cast<int32_t>(v);
        } else {
          double v = 0;
          if (!token_to_double(b, p, &v)) {
            status[t] = 1;
            return;
          }
          xy[2 * i + (field - 2)] = field == 3 ? -v : v;   // invert y axis, bal_problem.cpp:240
        }
      } else if (k < n_obs_tok + n_cam_tok) {
        if (!token_to_double(b, p, &out->cam_params[k - n_obs_tok])) {
          status[t] = 3;
          return;
        }
      } else if (k < n_obs_tok + n_cam_tok + n_lm_tok) {
        // the L x 3 landmark block must be present and numeric (the reference's loader reads it) but is not used
        double v;
        if (!token_to_double(b, p, &v)) {
          status[t] = 4;
          return;
        }
      }
      ++k;
    }
  });
  int bad = 0;
  for (int t = 0; t < T && bad == 0; ++t) bad = status[t];
  if (bad == 0 && total < n_obs_tok) bad = 1;
  else if (bad == 0 && total < n_obs_tok + n_cam_tok) bad = 3;
  else if (bad == 0 && total < n_obs_tok + n_cam_tok + n_lm_tok) bad = 4;
  if (bad != 0) {
    const char* what = bad == 1 ? "Failed to parse observations of '" : bad == 2 ? "Index out of range in '"
                       : bad == 3 ? "Failed to parse cameras of '" : "Failed to parse landmarks of '";
    set_err(err, err_len, std::string(what) + path + "'");
    povar_bal_free(out);
    return bad == 2 ? POVAR_ERR_INVALID : POVAR_ERR_IO;
  }
  std::vector<int64_t> perm(static_cast<size_t>(N));
  out->lm_ptr = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * (static_cast<size_t>(L) + 1)));
  const int rc = povar_canonical_order(static_cast<int32_t>(C), static_cast<int32_t>(L), N, cam.data(),
                                       lm.data(), perm.data(), out->lm_ptr);
  if (rc != POVAR_OK) {
    set_err(err, err_len, std::string("Invalid file '") + path + "' (duplicate observation)");
    povar_bal_free(out);
    return rc;
  }
  out->obs_cam = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * static_cast<size_t>(N)));
  out->obs_uv = static_cast<double*>(std::malloc(sizeof(double) * 2 * static_cast<size_t>(N)));
  for (long long i = 0; i < N; ++i) {
    const int64_t s = perm[i];
    out->obs_cam[i] = cam[s];
    out->obs_uv[2 * i] = xy[2 * s];
    out->obs_uv[2 * i + 1] = xy[2 * s + 1];
  }
  out->num_cams = static_cast<int32_t>(C);
  out->num_lms = static_cast<int32_t>(L);
  out->num_obs = N;
  return POVAR_OK;
}

// --create-dataset: BalProblem::load_bal_varproj_space_matrix_write (bal_problem.cpp:306-471).  Reads an
// original BAL file (C L N / N x (cam lm x y) / C x 9 / L x 3) and writes the 15-parameter file the
// solver loads: same header and observation lines (`%d %d %lf %lf`, i.e. six decimals), per camera the
// first two rows of the 3x4 matrix drawn from N(0,1) and the third row `0 0 0 1`, then f k1 k2 of the
// input, then the landmark block copied through; one number per line, like the reference's fprintf
// sequence.  The reference seeds std::mt19937 from std::random_device (seed < 0 here); with a seed the
// output is reproducible.  Like the reference it draws 15 variates per camera from a fresh
// std::normal_distribution and uses the first 8.
// no exception crosses the C ABI: allocation failures on hostile sizes come back as POVAR_ERR_IO
extern "C" int povar_bal_read(const char* path, povar_bal_data* out, char* err, size_t err_len) {
  try {
    return bal_read_impl(path, out, err, err_len);
  } catch (const std::exception& e) {
    set_err(err, err_len, std::string("Could not load '") + (path ? path : "") + "': " + e.what());
    if (out) povar_bal_free(out);
    return POVAR_ERR_IO;
  }
}

static int create_dataset_impl(const char* input, const char* output, int64_t seed, char* err, size_t err_len);
extern "C" int povar_bal_create_dataset(const char* input, const char* output, int64_t seed, char* err,
                                        size_t err_len) {
  try {
    return create_dataset_impl(input, output, seed, err, err_len);
  } catch (const std::exception& e) {
    set_err(err, err_len, std::string("--create-dataset failed: ") + e.what());
    return POVAR_ERR_IO;
  }
}

static int create_dataset_impl(const char* input, const char* output, int64_t seed, char* err, size_t err_len) {
  if (!input || !output) return POVAR_ERR_INVALID;
  FILE* f = std::fopen(input, "rb");
  if (!f) {
    set_err(err, err_len, std::string("Could not open '") + input + "'");
    return POVAR_ERR_IO;
  }
  std::vector<char> buf;
  if (!slurp(f, &buf)) {
    std::fclose(f);
    set_err(err, err_len, std::string("Could not read '") + input + "'");
    return POVAR_ERR_IO;
  }
  std::fclose(f);
  const size_t got = buf.size() - 1;
  Cursor cur{buf.data(), buf.data() + got};
  long long C = 0, L = 0, N = 0;
  if (!cur.next_int(&C) || !cur.next_int(&L) || !cur.next_int(&N) || C <= 0 || L <= 0 || N <= 0) {
    set_err(err, err_len, std::string("Failed to parse header of '") + input + "'");
    return POVAR_ERR_IO;
  }
  FILE* out = std::fopen(output, "w");
  if (!out) {
    set_err(err, err_len, std::string("Could not open '") + output + "' for writing");
    return POVAR_ERR_IO;
  }
  std::vector<char> obuf(1 << 22);
  std::setvbuf(out, obuf.data(), _IOFBF, obuf.size());
  auto bail = [&](const std::string& msg, int code) {
    std::fclose(out);
    set_err(err, err_len, msg);
    return code;
  };
  std::fprintf(out, "%lld %lld %lld", C, L, N);
  // duplicate (cam, lm) pairs are fatal in the reference (CHECK(inserted), bal_problem.cpp:366)
  std::vector<int32_t> cams(static_cast<size_t>(N)), lms(static_cast<size_t>(N));
  for (long long i = 0; i < N; ++i) {
    long long c = 0, l = 0;
    double x = 0, y = 0;
    if (!cur.next_int(&c) || !cur.next_int(&l) || !cur.next_double(&x) || !cur.next_double(&y)) {
      return bail(std::string("Failed to parse observations of '") + input + "'", POVAR_ERR_IO);
    }
    if (c < 0 || c >= C || l < 0 || l >= L) {
      return bail(std::string("Index out of range in '") + input + "'", POVAR_ERR_INVALID);
    }
    cams[i] = static_cast<int32_t>(c);
    lms[i] = static_cast<int32_t>(l);
    std::fprintf(out, "\n%lld %lld %lf %lf", c, l, x, y);
  }
  {
    std::vector<int64_t> perm(static_cast<size_t>(N)), lm_ptr(static_cast<size_t>(L) + 1);
    if (povar_canonical_order(static_cast<int32_t>(C), static_cast<int32_t>(L), N, cams.data(), lms.data(),
                              perm.data(), lm_ptr.data()) != POVAR_OK) {
      return bail(std::string("Invalid file '") + input + "' (duplicate observation)", POVAR_ERR_INVALID);
    }
  }
  std::mt19937 gen;
  if (seed < 0) {
    std::random_device rd;
    gen.seed(rd());
  } else {
    gen.seed(static_cast<std::mt19937::result_type>(seed));
  }
  for (long long i = 0; i < C; ++i) {
    double p9[9];
    for (double& v : p9) {
      if (!cur.next_double(&v)) return bail(std::string("Failed to parse cameras of '") + input + "'", POVAR_ERR_IO);
    }
    std::normal_distribution<double> d(0, 1);
    double draw[15];
    for (double& v : draw) v = d(gen);
    for (int m = 0; m < 8; ++m) std::fprintf(out, "\n%lf", draw[m]);
    std::fprintf(out, "\n%lf\n%lf\n%lf\n%lf", 0.0, 0.0, 0.0, 1.0);
    for (int m = 6; m < 9; ++m) std::fprintf(out, "\n%lf", p9[m]);
  }
  for (long long i = 0; i < 3 * L; ++i) {
    double v;
    if (!cur.next_double(&v)) return bail(std::string("Failed to parse landmarks of '") + input + "'", POVAR_ERR_IO);
    std::fprintf(out, "\n%lf", v);
  }
  if (std::fclose(out) != 0) {
    set_err(err, err_len, std::string("Failed to write '") + output + "'");
    return POVAR_ERR_IO;
  }
  return POVAR_OK;
}
