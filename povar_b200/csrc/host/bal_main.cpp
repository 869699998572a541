// `bal`: command-line front end with the reference's flag names
// (/root/reference/src/app/bal.cpp:44-103; flags generated from bal/solver_options.hpp:88-307,
// bal/bal_dataset_options.hpp, bal/ba_log_options.hpp by cli/cli_options.cpp:61) on top of
// libpovar_b200.so.  Writes ba_log.json with the per-iteration keys the reference writes
// (bal/ba_log.hpp:147-245, bal/ba_log.cpp:63-150).
//
// --num-gpus N forks one process per GPU (landmarks sharded by observation count, cameras
// replicated); rank 0 hands the NCCL id to the others through pipes.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "../../../include/povar_b200.h"

namespace {

struct AppOptions {
  std::string input;
  std::string log_path = "ba_log.json";
  int num_gpus = 1;
  std::vector<int> devices;      // --devices: CUDA ordinal of every rank (default: rank r on device r)
  bool create_dataset = false;
  long long dataset_seed = -1;   // < 0: std::random_device, like the reference
  povar_options solver;
};

[[noreturn]] void die(const std::string& msg) {
  std::fprintf(stderr, "bal: %s\n", msg.c_str());
  std::exit(1);
}

int parse_enum(const std::string& flag, const std::string& v, const std::map<std::string, int>& table) {
  auto it = table.find(v);
  if (it == table.end()) die("Could not convert value " + v + " for " + flag + " to enum");
  return it->second;
}

void usage() {
  std::printf(
      "Solve BAL problem with solver determined by config (B200 build).\n\n"
      "  -C, --directory <DIR>  change to DIR first\n  --config <PATH>  config file (default rootba_config.toml)\n"
      "  --dump-config  print the effective config and exit\n"
      "  --input <STR>\n  --create-dataset  write data_custom/<name> with randomised camera matrices and exit\n"
      "  --create-dataset-seed <INT> (default: std::random_device, like the reference)\n"
      "  --num-gpus <INT>  ranks (one forked process each; landmarks sharded, cameras replicated)\n"
      "  --devices <a,b,...>  CUDA ordinal of each rank (default 0,1,...); a repeated ordinal puts several\n"
      "                       ranks on one GPU and swaps the peer handles through shared memory instead of NCCL\n"
      "  --num-threads <INT> (ignored)\n"
      "  --solver-type-step-1 {POWER_VARPROJ,POWER_SCHUR_COMPLEMENT,POWER_BUNDLE_ADJUSTMENT,PCG,CHOLESKY}\n"
      "  --solver-type-step-2 {RIPOBA,RIPCG}\n  --power-sc-iterations <INT>\n"
      "  --residual-robust-norm {NONE,HUBER,CAUCHY}\n  --residual-huber-parameter <FLOAT>\n  --alpha <FLOAT>\n"
      "  --max-num-iterations-step-1 <INT>\n  --max-num-iterations-step-2 <INT>\n  --eta <FLOAT>\n"
      "  --r-tolerance <FLOAT>\n  --function-tolerance <FLOAT>\n  --initial-trust-region-radius <FLOAT>\n"
      "  --min-trust-region-radius <FLOAT>\n  --max-trust-region-radius <FLOAT>\n  --initial-vee <FLOAT>\n"
      "  --vee-factor <FLOAT>\n  --min-relative-decrease <FLOAT>\n  --optimized-cost {ERROR,ERROR_VALID,ERROR_VALID_AVG}\n"
      "  --jacobi-scaling-epsilon <FLOAT>\n  --min-linear-solver-iterations <INT>\n"
      "  --max-linear-solver-iterations <INT>\n  --preconditioner-type {JACOBI,SCHUR_JACOBI}\n"
      "  --verbosity-level <INT>\n  --log-log-path <STR>\n");
}

// one option, by its command-line name; `val` yields its value.  false: not an option of this program
bool apply_option(AppOptions& o, const std::string& f, const std::function<std::string()>& val) {
  const std::map<std::string, int> step1 = {{"PCG", POVAR_PCG},
                                            {"POWER_SCHUR_COMPLEMENT", POVAR_POWER_SCHUR_COMPLEMENT},
                                            // README name; the reference's own enum rejects it
                                            {"POWER_BUNDLE_ADJUSTMENT", POVAR_POWER_SCHUR_COMPLEMENT},
                                            {"POWER_VARPROJ", POVAR_POWER_VARPROJ},
                                            {"CHOLESKY", POVAR_CHOLESKY}};
  const std::map<std::string, int> step2 = {{"RIPOBA", POVAR_RIPOBA}, {"RIPCG", POVAR_RIPCG}};
  const std::map<std::string, int> norm = {{"NONE", POVAR_NORM_NONE}, {"HUBER", POVAR_NORM_HUBER}, {"CAUCHY", POVAR_NORM_CAUCHY}};
  const std::map<std::string, int> oc = {{"ERROR", POVAR_COST_ERROR}, {"ERROR_VALID", POVAR_COST_ERROR_VALID},
                                         {"ERROR_VALID_AVG", POVAR_COST_ERROR_VALID_AVG}};
  {
    if (f == "--input") o.input = val();
    else if (f == "--create-dataset") o.create_dataset = true;
    else if (f == "--no-create-dataset") o.create_dataset = false;
    else if (f == "--create-dataset-seed") o.dataset_seed = std::atoll(val().c_str());
    else if (f == "--num-gpus") o.num_gpus = std::atoi(val().c_str());
    else if (f == "--devices") {
      o.devices.clear();
      const std::string v = val();
      for (size_t p = 0; p < v.size();) {
        const size_t q = v.find(',', p);
        o.devices.push_back(std::atoi(v.substr(p, q == std::string::npos ? std::string::npos : q - p).c_str()));
        if (q == std::string::npos) break;
        p = q + 1;
      }
    }
    else if (f == "--num-threads") val();
    else if (f == "--solver-type-step-1") o.solver.solver_type_step_1 = parse_enum(f, val(), step1);
    else if (f == "--solver-type-step-2") o.solver.solver_type_step_2 = parse_enum(f, val(), step2);
    else if (f == "--power-sc-iterations") o.solver.power_sc_iterations = std::atoi(val().c_str());
    else if (f == "--residual-robust-norm") o.solver.robust_norm = parse_enum(f, val(), norm);
    else if (f == "--residual-huber-parameter") o.solver.huber_parameter = std::atof(val().c_str());
    else if (f == "--alpha") o.solver.alpha = std::atof(val().c_str());
    else if (f == "--max-num-iterations-step-1") o.solver.max_num_iterations_step_1 = std::atoi(val().c_str());
    else if (f == "--max-num-iterations-step-2") o.solver.max_num_iterations_step_2 = std::atoi(val().c_str());
    else if (f == "--eta") o.solver.eta = std::atof(val().c_str());
    else if (f == "--r-tolerance") o.solver.r_tolerance = std::atof(val().c_str());
    else if (f == "--function-tolerance") o.solver.function_tolerance = std::atof(val().c_str());
    else if (f == "--initial-trust-region-radius") o.solver.initial_trust_region_radius = std::atof(val().c_str());
    else if (f == "--min-trust-region-radius") o.solver.min_trust_region_radius = std::atof(val().c_str());
    else if (f == "--max-trust-region-radius") o.solver.max_trust_region_radius = std::atof(val().c_str());
    else if (f == "--initial-vee") o.solver.initial_vee = std::atof(val().c_str());
    else if (f == "--vee-factor") o.solver.vee_factor = std::atof(val().c_str());
    else if (f == "--min-relative-decrease") o.solver.min_relative_decrease = std::atof(val().c_str());
    else if (f == "--optimized-cost") o.solver.optimized_cost = parse_enum(f, val(), oc);
    else if (f == "--jacobi-scaling-epsilon") o.solver.jacobi_scaling_epsilon = std::atof(val().c_str());
    else if (f == "--min-linear-solver-iterations") o.solver.min_linear_solver_iterations = std::atoi(val().c_str());
    else if (f == "--max-linear-solver-iterations") o.solver.max_linear_solver_iterations = std::atoi(val().c_str());
    else if (f == "--preconditioner-type") {
      const std::string v = val();
      if (v != "SCHUR_JACOBI" && v != "JACOBI") die("predonditioner " + v + " not implemented");
    } else if (f == "--verbosity-level") o.solver.verbosity_level = std::atoi(val().c_str());
    else if (f == "--log-log-path") o.log_path = val();
    else if (f == "--quiet" || f == "--no-quiet" || f == "--normalize" || f == "--no-normalize" ||
             f == "--debug" || f == "--no-debug") {
      // accepted for command-line compatibility; no effect on this path
    } else {
      return false;
    }
  }
  return true;
}

// rootba_config.toml (options/options_interface.cpp:255-302, cli/bal_cli_utils.cpp:105-115): the sections
// [dataset], [solver], [solver.residual], [solver.log] hold the same options as the command line
// (--<key>, --residual-<key>, --log-<key>, '_' -> '-'); the file is read first, the command line overrides it,
// a missing file means defaults.  The subset of TOML the reference's own --dump-config writes is understood:
// tables, `key = "string" | number | true | false`, arrays (skipped), comments.
void load_config(AppOptions& o, const std::string& path, bool verbose) {
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) {
    if (verbose) std::printf("Config file %s doesn't exist. Loading defaults.\n", path.c_str());
    return;
  }
  std::string section;
  char line[4096];
  int loaded = 0, depth = 0;
  while (std::fgets(line, sizeof(line), f)) {
    std::string t(line);
    // strip comments outside strings
    bool in_str = false;
    for (size_t i = 0; i < t.size(); ++i) {
      if (t[i] == '"') in_str = !in_str;
      if (t[i] == '#' && !in_str) {
        t.erase(i);
        break;
      }
    }
    auto trim = [](std::string v) {
      const size_t b = v.find_first_not_of(" \t\r\n");
      if (b == std::string::npos) return std::string();
      return v.substr(b, v.find_last_not_of(" \t\r\n") - b + 1);
    };
    t = trim(t);
    if (t.empty()) continue;
    if (depth > 0) {   // inside a multi-line array
      for (char ch : t) depth += (ch == '[') - (ch == ']');
      continue;
    }
    if (t.front() == '[') {
      section = trim(t.substr(1, t.find(']') - 1));
      continue;
    }
    const size_t eq = t.find('=');
    if (eq == std::string::npos) continue;
    const std::string key = trim(t.substr(0, eq));
    std::string value = trim(t.substr(eq + 1));
    if (!value.empty() && value.front() == '[') {
      for (char ch : value) depth += (ch == '[') - (ch == ']');
      continue;   // arrays (save_log_flags): nothing this program uses
    }
    if (value.size() >= 2 && value.front() == '"' && value.back() == '"') value = value.substr(1, value.size() - 2);
    std::string prefix;
    if (section == "solver.residual") prefix = "residual-";
    else if (section == "solver.log") prefix = "log-";
    else if (section != "solver" && section != "dataset") continue;   // /batch_run, /slurm, ...
    std::string name = key;
    std::replace(name.begin(), name.end(), '_', '-');
    bool known;
    if (value == "true" || value == "false") {
      known = apply_option(o, (value == "true" ? "--" : "--no-") + prefix + name, [] { return std::string(); });
    } else {
      known = apply_option(o, "--" + prefix + name, [&] { return value; });
    }
    if (known) ++loaded;
  }
  std::fclose(f);
  if (verbose) std::printf("Loaded %d items from config file %s\n", loaded, path.c_str());
}

const char* enum_name(const std::map<std::string, int>& table, int v) {
  for (const auto& kv : table) {
    if (kv.second == v && kv.first != "POWER_BUNDLE_ADJUSTMENT") return kv.first.c_str();
  }
  return "?";
}

// --dump-config: the effective options in the reference's layout (only keys both programs know)
void dump_config(const AppOptions& o) {
  const std::map<std::string, int> step1 = {{"PCG", POVAR_PCG}, {"POWER_SCHUR_COMPLEMENT", POVAR_POWER_SCHUR_COMPLEMENT},
                                            {"POWER_VARPROJ", POVAR_POWER_VARPROJ}, {"CHOLESKY", POVAR_CHOLESKY}};
  const std::map<std::string, int> step2 = {{"RIPOBA", POVAR_RIPOBA}, {"RIPCG", POVAR_RIPCG}};
  const std::map<std::string, int> norm = {{"NONE", POVAR_NORM_NONE}, {"HUBER", POVAR_NORM_HUBER}, {"CAUCHY", POVAR_NORM_CAUCHY}};
  const std::map<std::string, int> oc = {{"ERROR", POVAR_COST_ERROR}, {"ERROR_VALID", POVAR_COST_ERROR_VALID},
                                         {"ERROR_VALID_AVG", POVAR_COST_ERROR_VALID_AVG}};
  const povar_options& s = o.solver;
  std::printf("\n[dataset]\ninput = \"%s\"\ncreate_dataset = %s\n", o.input.c_str(), o.create_dataset ? "true" : "false");
  std::printf("\n[solver]\nsolver_type_step_1 = \"%s\"\nsolver_type_step_2 = \"%s\"\nverbosity_level = %d\n",
              enum_name(step1, s.solver_type_step_1), enum_name(step2, s.solver_type_step_2), s.verbosity_level);
  std::printf("alpha = %.17g\noptimized_cost = \"%s\"\nmax_num_iterations_step_1 = %d\nmax_num_iterations_step_2 = %d\n",
              s.alpha, enum_name(oc, s.optimized_cost), s.max_num_iterations_step_1, s.max_num_iterations_step_2);
  std::printf("min_relative_decrease = %.17g\ninitial_trust_region_radius = %.17g\nmin_trust_region_radius = %.17g\n"
              "max_trust_region_radius = %.17g\n",
              s.min_relative_decrease, s.initial_trust_region_radius, s.min_trust_region_radius,
              s.max_trust_region_radius);
  std::printf("min_linear_solver_iterations = %d\nmax_linear_solver_iterations = %d\neta = %.17g\nr_tolerance = %.17g\n",
              s.min_linear_solver_iterations, s.max_linear_solver_iterations, s.eta, s.r_tolerance);
  std::printf("jacobi_scaling_epsilon = %.17g\npreconditioner_type = \"SCHUR_JACOBI\"\nfunction_tolerance = %.17g\n"
              "power_sc_iterations = %d\ninitial_vee = %.17g\nvee_factor = %.17g\n",
              s.jacobi_scaling_epsilon, s.function_tolerance, s.power_sc_iterations, s.initial_vee, s.vee_factor);
  std::printf("\n[solver.residual]\nrobust_norm = \"%s\"\nhuber_parameter = %.17g\n", enum_name(norm, s.robust_norm),
              s.huber_parameter);
  std::printf("\n[solver.log]\nlog_path = \"%s\"\n", o.log_path.c_str());
}

AppOptions parse(int argc, char** argv) {
  AppOptions o;
  povar_options_default(&o.solver);
  // -C / --config / --dump-config first: the config file is read before the other arguments (bal_cli_utils.cpp:96-115)
  std::string config_path = "rootba_config.toml", working_dir;
  bool dump = false, quiet_cli = false;
  for (int i = 1; i < argc; ++i) {
    const std::string f = argv[i];
    if ((f == "--config" || f == "-C" || f == "--directory") && i + 1 >= argc) die("missing value for " + f);
    if (f == "--config") config_path = argv[++i];
    else if (f == "-C" || f == "--directory") working_dir = argv[++i];
    else if (f == "--dump-config") dump = true;
    else if (f == "--verbosity-level" && i + 1 < argc) quiet_cli = std::atoi(argv[i + 1]) == 0;
    else if (f == "--help" || f == "-h") {
      usage();
      std::exit(0);
    }
  }
  if (!working_dir.empty() && chdir(working_dir.c_str()) != 0) die("cannot change to directory " + working_dir);
  load_config(o, config_path, !quiet_cli && !dump);
  for (int i = 1; i < argc; ++i) {
    const std::string f = argv[i];
    if (f == "--config" || f == "-C" || f == "--directory") {
      ++i;
      continue;
    }
    if (f == "--dump-config") continue;
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) die("missing value for " + f);
      return argv[++i];
    };
    if (!apply_option(o, f, val)) die("unknown argument " + f);
  }
  if (dump) {
    dump_config(o);
    std::exit(1);   // like the reference: parse_bal_app_arguments returns false after printing (app/bal.cpp:53-57)
  }
  if (o.input.empty()) die("--input is required");
  if (o.num_gpus < 1) die("--num-gpus must be >= 1");
  if (!o.devices.empty() && static_cast<int>(o.devices.size()) != o.num_gpus) die("--devices needs one ordinal per rank");
  return o;
}

void write_log(const AppOptions& o, const povar_bal_data& data, const std::vector<povar_iteration>& its,
               const povar_solve_summary& s, double load_time) {
  povar_ba_log_info info;
  info.input_path = o.input.c_str();
  info.num_cams = data.num_cams;
  info.num_lms = data.num_lms;
  info.num_obs = data.num_obs;
  info.lm_ptr = data.lm_ptr;
  info.load_time = load_time;
  info.num_gpus = o.num_gpus;
  if (povar_write_ba_log(o.log_path.c_str(), &info, &o.solver, its.data(), static_cast<int32_t>(its.size()), &s) !=
      POVAR_OK) {
    std::fprintf(stderr, "bal: Could not save BA log to %s.\n", o.log_path.c_str());
  }
}

}  // namespace

int main(int argc, char** argv) {
  AppOptions o = parse(argc, argv);
  if (o.create_dataset) {
    // bal_problem.cpp:306-319, 899-903: data_custom/<basename of the input>, then exit(0)
    mkdir("data_custom", 0777);
    const size_t slash = o.input.find_last_of('/');
    const std::string out = "data_custom/" + (slash == std::string::npos ? o.input : o.input.substr(slash + 1));
    char cerr[512] = {0};
    if (povar_bal_create_dataset(o.input.c_str(), out.c_str(), o.dataset_seed, cerr, sizeof(cerr)) != POVAR_OK) die(cerr);
    if (o.solver.verbosity_level >= 1) std::printf("Wrote '%s'\n", out.c_str());
    return 0;
  }

  // fork the other ranks BEFORE any CUDA / NCCL call
  int rank = 0;
  std::vector<int> write_fds;
  int read_fd = -1;
  std::vector<pid_t> children;
  for (int r = 1; r < o.num_gpus; ++r) {
    int fds[2];
    if (pipe(fds) != 0) die("pipe failed");
    const pid_t pid = fork();
    if (pid < 0) die("fork failed");
    if (pid == 0) {
      rank = r;
      read_fd = fds[0];
      close(fds[1]);
      for (int w : write_fds) close(w);
      write_fds.clear();
      children.clear();
      break;
    }
    close(fds[0]);
    write_fds.push_back(fds[1]);
    children.push_back(pid);
  }

  const auto t_load = std::chrono::steady_clock::now();
  povar_bal_data data;
  char err[512] = {0};
  int rc = povar_bal_read(o.input.c_str(), &data, err, sizeof(err));
  if (rc != POVAR_OK) die(err);
  const double load_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
  if (rank == 0 && o.solver.verbosity_level >= 1) {
    std::printf("Loaded BAL problem (%d cams, %d lms, %lld obs) from '%s'\n", data.num_cams, data.num_lms,
                static_cast<long long>(data.num_obs), o.input.c_str());
  }

  povar_comm_desc comm;
  std::memset(&comm, 0, sizeof(comm));
  comm.rank = rank;
  comm.world_size = o.num_gpus;
  comm.device = o.devices.empty() ? rank : o.devices[rank];
  if (o.num_gpus > 1) {
    if (rank == 0) {
      // ranks that share a device cannot form an NCCL communicator: host rendezvous (povar_comm_host_id)
      std::vector<int> sorted = o.devices;
      std::sort(sorted.begin(), sorted.end());
      const bool shared_device = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
      rc = shared_device ? povar_comm_host_id(comm.nccl_id) : povar_comm_unique_id(comm.nccl_id);
      if (rc != POVAR_OK) die(std::string("NCCL: ") + povar_last_error(nullptr));
      for (int w : write_fds) {
        if (write(w, comm.nccl_id, 128) != 128) die("pipe write failed");
        close(w);
      }
    } else {
      if (read(read_fd, comm.nccl_id, 128) != 128) die("pipe read failed");
      close(read_fd);
    }
  }

  std::vector<int32_t> bounds(o.num_gpus + 1);
  povar_partition_landmarks(data.num_lms, data.lm_ptr, o.num_gpus, bounds.data());
  const int32_t lb = bounds[rank], le = bounds[rank + 1];
  std::vector<int64_t> lm_ptr(le - lb + 1);
  for (int32_t l = lb; l <= le; ++l) lm_ptr[l - lb] = data.lm_ptr[l] - data.lm_ptr[lb];
  std::vector<double> cam_P(static_cast<size_t>(data.num_cams) * 12);
  for (int c = 0; c < data.num_cams; ++c) std::memcpy(&cam_P[12 * c], &data.cam_params[15 * c], 12 * sizeof(double));

  povar_problem_desc desc;
  desc.num_cams = data.num_cams;
  desc.num_lms = le - lb;
  desc.num_obs = data.lm_ptr[le] - data.lm_ptr[lb];
  desc.lm_ptr = lm_ptr.data();
  desc.obs_cam = data.obs_cam + data.lm_ptr[lb];
  desc.obs_uv = data.obs_uv + 2 * data.lm_ptr[lb];
  desc.cam_P = cam_P.data();

  povar_handle* h = nullptr;
  rc = povar_create(&desc, &o.solver, o.num_gpus > 1 ? &comm : nullptr, &h);
  if (rc != POVAR_OK) die(std::string("povar_create: ") + povar_last_error(nullptr));

  const int cap = o.solver.max_num_iterations_step_1 + o.solver.max_num_iterations_step_2 + 4;
  std::vector<povar_iteration> its(cap);
  povar_solve_summary summary;
  rc = povar_bundle_adjust(h, &o.solver, its.data(), cap, &summary);
  if (rc != POVAR_OK) {
    std::fprintf(stderr, "bal: solve failed (%d): %s / %s\n", rc, summary.message, povar_last_error(h));
  } else if (rank == 0) {
    its.resize(summary.num_iterations);
    if (o.solver.verbosity_level >= 1) {
      std::printf("Final Cost: error: %.4e; %s\n", summary.final_cost, summary.message);
      std::printf("solve %.4f s (step 1 %.4f s, step 2 %.4f s), %d LM iterations, %lld power terms in %.4f s\n",
                  summary.total_time, summary.step1_time, summary.step2_time, summary.num_iterations,
                  static_cast<long long>(summary.power_terms), summary.power_series_time);
    }
    write_log(o, data, its, summary, load_time);
  }
  povar_destroy(h);
  povar_bal_free(&data);
  int status = rc == POVAR_OK ? 0 : 2;
  for (pid_t pid : children) {
    int st = 0;
    waitpid(pid, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) status = 2;
  }
  return status;
}
