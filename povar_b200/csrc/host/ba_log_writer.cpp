// ba_log.json writer with the reference's complete column set, so that the reference's own tooling
// (/root/reference/python/rootba/log.py, metric.py: BaLog, l._static.solver.*, per-iteration arrays) loads
// our logs unchanged.  Keys and their meaning follow
//   /root/reference/src/rootba_povar/bal/ba_log.hpp:85-245 (BaLog::Static, BaLog::Iteration),
//   bal/ba_log_utils.cpp:100-175 (which summary field goes where; failed trials repeat the previous cost),
//   solver/bal_bundle_adjustment.cpp:61-150 (finish_iteration / finish_solve: derived sums).
// No CUDA in this file.  Columns the GPU path has no counterpart for are written as the reference writes
// them when it does not compute them (0 / false / ""): gradient norms, step_norm, logging_time,
// perform_qr_time, compute_gradient_time, compute_preconditioner_time, grouping / merge fields.
#include <sys/resource.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/povar_b200.h"

namespace {

const char* step1_name(int t) {   // finish_solve, bal_bundle_adjustment.cpp:98-113
  switch (t) {
    case POVAR_PCG: return "bal_pcg";
    case POVAR_POWER_SCHUR_COMPLEMENT: return "bal_power_sc";
    case POVAR_POWER_VARPROJ: return "power_variable_projection";
    default: return "variable_projection";
  }
}

std::string json_escape(const char* s) {
  std::string out;
  for (const char* p = s ? s : ""; *p; ++p) {
    const unsigned char c = static_cast<unsigned char>(*p);
    if (c == '"' || c == '\\') {
      out += '\\';
      out += static_cast<char>(c);
    } else if (c == '\n') {
      out += "\\n";
    } else if (c < 0x20) {
      char buf[8];
      std::snprintf(buf, sizeof(buf), "\\u%04x", c);
      out += buf;
    } else {
      out += static_cast<char>(c);
    }
  }
  return out;
}

void num(FILE* f, double v) {
  if (std::isfinite(v)) std::fprintf(f, "%.17g", v);
  else std::fprintf(f, "null");
}

struct Columns {
  FILE* f;
  int n;
  bool first = true;
  void key(const char* k) {
    std::fprintf(f, "%s\n    \"%s\": [", first ? "" : ",", k);
    first = false;
  }
  template <typename F>
  void doubles(const char* k, F get) {
    key(k);
    for (int i = 0; i < n; ++i) {
      if (i) std::fprintf(f, ", ");
      num(f, get(i));
    }
    std::fprintf(f, "]");
  }
  template <typename F>
  void ints(const char* k, F get) {
    key(k);
    for (int i = 0; i < n; ++i) std::fprintf(f, "%s%lld", i ? ", " : "", static_cast<long long>(get(i)));
    std::fprintf(f, "]");
  }
  template <typename F>
  void bools(const char* k, F get) {
    key(k);
    for (int i = 0; i < n; ++i) std::fprintf(f, "%s%s", i ? ", " : "", get(i) ? "true" : "false");
    std::fprintf(f, "]");
  }
  template <typename F>
  void strings(const char* k, F get) {
    key(k);
    for (int i = 0; i < n; ++i) std::fprintf(f, "%s\"%s\"", i ? ", " : "", get(i));
    std::fprintf(f, "]");
  }
};

}  // namespace

extern "C" int povar_write_ba_log(const char* path, const povar_ba_log_info* info, const povar_options* opt,
                                  const povar_iteration* its, int32_t n, const povar_solve_summary* s) {
  if (!path || !info || !opt || !s || n < 0 || (n > 0 && !its)) return POVAR_ERR_INVALID;
  FILE* f = std::fopen(path, "w");
  if (!f) return POVAR_ERR_IO;
  // per-landmark observation statistics (DatasetSummary, bal/bal_problem.cpp: per_lm_obs)
  double lm_min = 0, lm_max = 0, lm_mean = 0, lm_std = 0;
  if (info->lm_ptr && info->num_lms > 0) {
    lm_min = 1e300;
    double sum = 0, sum2 = 0;
    for (int32_t l = 0; l < info->num_lms; ++l) {
      const double d = static_cast<double>(info->lm_ptr[l + 1] - info->lm_ptr[l]);
      lm_min = std::min(lm_min, d);
      lm_max = std::max(lm_max, d);
      sum += d;
      sum2 += d * d;
    }
    lm_mean = sum / info->num_lms;
    lm_std = std::sqrt(std::max(0.0, sum2 / info->num_lms - lm_mean * lm_mean));
  }
  long long rss_peak = 0;
  {
    struct rusage ru;
    if (getrusage(RUSAGE_SELF, &ru) == 0) rss_peak = static_cast<long long>(ru.ru_maxrss) * 1024;
  }
  // sums of finish_solve (bal_bundle_adjustment.cpp:138-150)
  double t_lin = 0, t_res = 0, t_jac = 0;
  int n_lin = 0, n_jac = 0, n_res = 0;
  for (int i = 0; i < n; ++i) {
    t_lin += its[i].solve_reduced_system_time + its[i].back_substitution_time + its[i].prepare_time;
    t_res += its[i].residual_evaluation_time;
    t_jac += its[i].jacobian_evaluation_time;
    if (its[i].iteration > 0) ++n_lin;
    if (its[i].jacobian_evaluation_time > 0) ++n_jac;
    if (std::isfinite(its[i].trial_cost)) ++n_res;
  }
  const unsigned hw = std::thread::hardware_concurrency();

  std::fprintf(f, "{\n    \"_type\": \"rootba_povar\",\n    \"_static\": {\n");
  std::fprintf(f, "        \"problem_info\": {\"type\": \"bal\", \"input_path\": \"%s\", \"num_cameras\": %d, "
                  "\"num_landmarks\": %d, \"num_observations\": %lld, \"rcs_sparsity\": 0.0,\n",
               json_escape(info->input_path).c_str(), info->num_cams, info->num_lms,
               static_cast<long long>(info->num_obs));
  std::fprintf(f, "            \"per_lm_obs\": {\"min\": %.17g, \"max\": %.17g, \"mean\": %.17g, \"stddev\": %.17g},\n",
               lm_min, lm_max, lm_mean, lm_std);
  std::fprintf(f, "            \"per_host_lms\": {\"min\": 0.0, \"max\": 0.0, \"mean\": 0.0, \"stddev\": 0.0}},\n");
  std::fprintf(f, "        \"timing\": {\"load\": %.9g, \"preprocess\": 0.0, \"optimize\": %.9g, \"postprocess\": 0.0, "
                  "\"total\": %.9g},\n",
               info->load_time, s->total_time, info->load_time + s->total_time);
  std::fprintf(f, "        \"solver\": {\"solver_type\": \"%s\", \"termination_type\": %d, \"termination_type_step_1\": %d, "
                  "\"message\": \"%s\",\n",
               step1_name(opt->solver_type_step_1), s->termination_type_step_2, s->termination_type_step_1,
               json_escape(s->message).c_str());
  std::fprintf(f, "            \"num_successful_steps\": %d, \"num_unsuccessful_steps\": %d, \"num_linear_solves\": %d, "
                  "\"num_residual_evaluations\": %d, \"num_jacobian_evaluations\": %d,\n",
               s->num_successful_steps, s->num_unsuccessful_steps, n_lin, n_res, n_jac);
  std::fprintf(f, "            \"total_time_in_seconds\": %.9g, \"minimizer_time_in_seconds\": %.9g, "
                  "\"preprocessor_time_in_seconds\": 0.0, \"postprocessor_time_in_seconds\": 0.0, "
                  "\"logging_time_in_seconds\": 0.0,\n",
               s->total_time, s->total_time);
  std::fprintf(f, "            \"linear_solver_time_in_seconds\": %.9g, \"residual_evaluation_time_in_seconds\": %.9g, "
                  "\"jacobian_evaluation_time_in_seconds\": %.9g,\n",
               t_lin, t_res, t_jac);
  std::fprintf(f, "            \"fraction_grouped\": 0.0, \"grouping_time_in_seconds\": 0.0, \"merge_factor\": true, "
                  "\"num_threads_given\": 0, \"num_threads_used\": 1, \"num_threads_available\": %u, "
                  "\"resident_memory_peak\": %lld,\n",
               hw, rss_peak);
  // ours, beyond the reference's keys
  std::fprintf(f, "            \"step_1_time_in_seconds\": %.9g, \"step_2_time_in_seconds\": %.9g, "
                  "\"power_series_terms\": %lld, \"power_series_time_in_seconds\": %.9g, \"num_gpus\": %d,\n",
               s->step1_time, s->step2_time, static_cast<long long>(s->power_terms), s->power_series_time,
               info->num_gpus);
  std::fprintf(f, "            \"initial_cost\": ");
  num(f, s->initial_cost);
  std::fprintf(f, ", \"final_cost\": ");
  num(f, s->final_cost);
  std::fprintf(f, "}\n    }");

  // change = previous - current (bal/residual_info.cpp:43-53) against the previous summary (ba_log_utils.cpp:106-141: only for successful
  // trials with iteration > 0; the previous summary's own cost is its trial cost when it failed)
  auto prev_raw_cost = [&](int i) {
    const povar_iteration& p = its[i - 1];
    return (!p.step_is_successful && std::isfinite(p.trial_cost)) ? p.trial_cost : p.cost;
  };
  auto changes = [&](int i) { return i > 0 && its[i].step_is_successful && its[i].iteration > 0; };
  auto avg_valid = [&](int i) {
    return its[i].num_obs_valid > 0 ? its[i].cost_valid / static_cast<double>(its[i].num_obs_valid) : 0.0;
  };
  const char* lin1 = (opt->solver_type_step_1 == POVAR_PCG || opt->solver_type_step_1 == POVAR_CHOLESKY) ? "bal_sc" : "bal_power_sc";
  const char* lin2 = opt->solver_type_step_2 == POVAR_RIPCG ? "bal_sc" : "bal_power_sc";
  std::fprintf(f, ",");
  Columns c{f, n};
  c.first = true;
  // the leading comma was written above; Columns::key adds one between columns only
  c.ints("iteration", [&](int i) { return its[i].iteration; });
  c.ints("step", [&](int i) { return its[i].step; });
  c.bools("step_is_valid", [&](int i) { return its[i].step_is_valid != 0; });
  c.bools("step_is_nonmonotonic", [&](int) { return false; });
  c.bools("step_is_successful", [&](int i) { return its[i].step_is_successful != 0; });
  c.ints("num_obs", [&](int) { return info->num_obs; });
  c.ints("num_obs_valid", [&](int i) { return its[i].num_obs_valid; });
  c.ints("num_obs_valid_change", [&](int i) { return changes(i) ? its[i - 1].num_obs_valid - its[i].num_obs_valid : 0; });
  c.doubles("cost", [&](int i) { return its[i].cost; });
  c.doubles("cost_change", [&](int i) { return changes(i) ? prev_raw_cost(i) - its[i].cost : 0.0; });
  c.doubles("cost_valid", [&](int i) { return its[i].cost_valid; });
  c.doubles("cost_valid_change", [&](int i) { return changes(i) ? its[i - 1].cost_valid - its[i].cost_valid : 0.0; });
  c.doubles("cost_avg_valid", avg_valid);
  c.doubles("cost_avg_valid_change", [&](int i) { return changes(i) ? avg_valid(i - 1) - avg_valid(i) : 0.0; });
  c.doubles("residual_block_mean", [&](int i) { return its[i].residual_mean; });
  c.doubles("residual_block_valid_mean", [&](int i) { return its[i].residual_valid_mean; });
  c.doubles("grad_max_norm", [&](int) { return 0.0; });
  c.doubles("grad_norm", [&](int) { return 0.0; });
  c.doubles("grad_projected_max_norm", [&](int) { return 0.0; });
  c.doubles("grad_projected_norm", [&](int) { return 0.0; });
  c.doubles("step_norm", [&](int) { return 0.0; });
  c.doubles("relative_decrease", [&](int i) { return std::isfinite(its[i].relative_decrease) ? its[i].relative_decrease : 0.0; });
  c.doubles("trust_region_radius", [&](int i) { return its[i].trust_region_radius; });
  c.ints("linear_solver_iterations", [&](int i) { return its[i].linear_solver_iterations; });
  c.strings("linear_solver_type", [&](int i) { return its[i].iteration == 0 ? "" : (its[i].step == 2 ? lin2 : lin1); });
  c.doubles("iteration_time", [&](int i) { return its[i].iteration_time; });
  c.doubles("cumulative_time", [&](int i) { return its[i].cumulative_time; });
  c.doubles("logging_time", [&](int) { return 0.0; });
  // finish_iteration (bal_bundle_adjustment.cpp:61-71): scale_landmark_jacobian + perform_qr + stage2 +
  // solve_reduced_system + back_substitution; stage 2 of the reference (prepare_Hb) is our prepare phase
  c.doubles("step_solver_time", [&](int i) {
    return its[i].prepare_time + its[i].solve_reduced_system_time + its[i].back_substitution_time;
  });
  c.doubles("residual_evaluation_time", [&](int i) { return its[i].residual_evaluation_time; });
  c.doubles("jacobian_evaluation_time", [&](int i) { return its[i].jacobian_evaluation_time; });
  c.doubles("scale_landmark_jacobian_time", [&](int) { return 0.0; });   // inside jacobian_evaluation_time here
  c.doubles("perform_qr_time", [&](int) { return 0.0; });
  c.doubles("stage1_time", [&](int i) { return its[i].jacobian_evaluation_time; });
  c.doubles("scale_pose_jacobian_time", [&](int) { return 0.0; });
  c.doubles("landmark_damping_time", [&](int) { return 0.0; });
  c.doubles("compute_preconditioner_time", [&](int) { return 0.0; });
  c.doubles("compute_gradient_time", [&](int) { return 0.0; });
  c.doubles("stage2_time", [&](int i) { return its[i].prepare_time; });
  c.doubles("prepare_time", [&](int i) { return its[i].prepare_time; });
  c.doubles("solve_reduced_system_time", [&](int i) { return its[i].solve_reduced_system_time; });
  c.doubles("back_substitution_time", [&](int i) { return its[i].back_substitution_time; });
  c.doubles("update_cameras_time", [&](int) { return 0.0; });            // inside back_substitution_time here
  c.ints("resident_memory", [&](int) { return rss_peak; });
  c.ints("resident_memory_peak", [&](int) { return rss_peak; });
  std::fprintf(f, "\n}\n");
  if (std::fclose(f) != 0) return POVAR_ERR_IO;
  return POVAR_OK;
}
