// LinearizorB200: the reference's own `Linearizor<Scalar>` interface
// (/root/reference/src/rootba_povar/solver/linearizor.hpp:47-82) implemented on libpovar_b200.so.
//
// This is the reference-side binding of INTEGRATION.md 2, as real code: with this header and
// linearizor_factory_b200.cpp in place of solver/linearizor.cpp, the reference's UNMODIFIED driver
// (bundle_adjust_manual, solver/bal_bundle_adjustment.cpp:252-876) runs both LM loops on the GPU library.
// `make -C oracle plugin` builds that program (oracle/_ref/bal_ref_b200) from the reference's sources.
//
// The reference's caller owns BalProblem and mutates it between calls (backup_* / restore_*, the step-2
// normalisation, create_homogeneous_landmark), so the host copy is compared with what the device holds before
// every cost evaluation, linearisation and solve -- and uploaded only if the caller changed it -- and pulled back
// after every apply.  A driver that uses povar_backup / povar_restore / povar_normalize_joint instead
// (host/lm_driver.cpp) needs none of these copies.
//
// Phase times: every method copies the library's CUDA-event times (povar_get_timings) into the IterationSummary
// fields the reference's own linearizors fill through IF_SET(it_summary_)
// (solver/linearizor_power_varproj.cpp:61-306, solver/solver_summary.hpp:172-212).
//
// Several GPUs behind the reference's driver: run one process of the driver per rank with
//   POVAR_PLUGIN_WORLD=N POVAR_PLUGIN_RANK=r POVAR_PLUGIN_DEVICE=d POVAR_PLUGIN_ID_FILE=<path>
// Every process loads the whole BalProblem (the reference's loader), uploads and downloads only the landmarks
// of its shard (povar_partition_landmarks) and takes the same decisions as the others, because every number the
// driver sees is reduced over the shards inside the library.  Rank 0 writes the communicator id to the file
// (host rendezvous id if POVAR_PLUGIN_HOST_ID=1 -- ranks may then share a device -- else an NCCL id).
#pragma once

#include <povar_b200.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <glog/logging.h>

#include "rootba_povar/solver/linearizor.hpp"

namespace rootba_povar {

template <class Scalar_>   // instantiated for double only: the device path is FP64 like the reference
class LinearizorB200 : public Linearizor<Scalar_> {
 public:
  using Scalar = Scalar_;
  using VecX = Eigen::Matrix<Scalar, Eigen::Dynamic, 1>;

  LinearizorB200(BalProblem<Scalar>& problem, const SolverOptions& o, SolverSummary* summary, bool joint)
      : problem_(problem), summary_(summary), joint_(joint), alpha_(o.alpha) {
    static_assert(sizeof(Scalar) == sizeof(double), "libpovar_b200 computes in FP64");
    // the shard of this process: all landmarks, or the range povar_partition_landmarks gives this rank
    povar_comm_desc comm;
    const bool sharded = comm_from_environment(&comm);
    const int32_t L_all = static_cast<int32_t>(problem_.landmarks().size());
    lm_begin_ = 0;
    lm_end_ = L_all;
    if (sharded) {
      std::vector<int64_t> full_ptr{0};
      for (const auto& lm : problem_.landmarks()) full_ptr.push_back(full_ptr.back() + static_cast<int64_t>(lm.obs.size()));
      std::vector<int32_t> bounds(comm.world_size + 1);
      CHECK_EQ(povar_partition_landmarks(L_all, full_ptr.data(), comm.world_size, bounds.data()), POVAR_OK);
      lm_begin_ = bounds[comm.rank];
      lm_end_ = bounds[comm.rank + 1];
    }
    // canonical order = landmark index, then the std::map<cam, obs> order (bal_problem.hpp:226)
    std::vector<int64_t> lm_ptr{0};
    std::vector<int32_t> obs_cam;
    std::vector<double> obs_uv;
    for (int32_t l = lm_begin_; l < lm_end_; ++l) {
      const auto& lm = problem_.landmarks()[l];
      for (const auto& [cam, obs] : lm.obs) {
        obs_cam.push_back(static_cast<int32_t>(cam));
        obs_uv.push_back(obs.pos.x());
        obs_uv.push_back(obs.pos.y());          // already y-flipped by load_bal_eccv (bal_problem.cpp:240)
      }
      lm_ptr.push_back(static_cast<int64_t>(obs_cam.size()));
    }
    std::vector<double> cam_P;
    pack_cameras(cam_P);
    povar_options po;
    povar_options_default(&po);
    po.solver_type_step_1 = static_cast<int>(o.solver_type_step_1);   // same enumerators, same order
    po.solver_type_step_2 = static_cast<int>(o.solver_type_step_2);   // (solver_options.hpp:56-75)
    po.robust_norm = static_cast<int>(o.residual.robust_norm);
    po.huber_parameter = o.residual.huber_parameter;
    po.alpha = o.alpha;
    po.eta = o.eta;
    po.r_tolerance = o.r_tolerance;
    po.power_sc_iterations = o.power_sc_iterations;
    po.jacobi_scaling_epsilon = o.jacobi_scaling_epsilon;
    po.min_linear_solver_iterations = o.min_linear_solver_iterations;
    po.max_linear_solver_iterations = o.max_linear_solver_iterations;
    po.verbosity_level = 0;
    povar_problem_desc d;
    d.num_cams = static_cast<int32_t>(problem_.cameras().size());
    d.num_lms = lm_end_ - lm_begin_;
    d.num_obs = static_cast<int64_t>(obs_cam.size());
    d.lm_ptr = lm_ptr.data();
    d.obs_cam = obs_cam.data();
    d.obs_uv = obs_uv.data();
    d.cam_P = cam_P.data();
    CHECK_EQ(povar_create(&d, &po, sharded ? &comm : nullptr, &h_), POVAR_OK) << povar_last_error(nullptr);
  }
  ~LinearizorB200() override { povar_destroy(h_); }

  void start_iteration(IterationSummary* it_summary = nullptr) override {
    it_ = it_summary;
    povar_reset_timings(h_);
  }
  void finish_iteration() override { it_ = nullptr; }

  void initialize_varproj_lm_pOSE(Scalar alpha, bool initialization_varproj) override {
    if (!initialization_varproj) return;
    push_state(POVAR_STATE_POSE, /*landmarks=*/false);
    CHECK_EQ(povar_init_varproj(h_, alpha), POVAR_OK) << povar_last_error(h_);
    pull_state(POVAR_STATE_POSE);
  }
  void compute_error_pOSE(ResidualInfo& ri, bool /*initialization_varproj*/) override {
    push_state(POVAR_STATE_POSE, true);
    povar_residual_info r;
    CHECK_EQ(povar_cost_pose(h_, alpha_, &r), POVAR_OK) << povar_last_error(h_);
    fill(ri, r);
  }
  void compute_error_homogeneous(ResidualInfo& ri, bool /*initialization_varproj*/) override {
    push_state(POVAR_STATE_JOINT, true);
    povar_residual_info r;
    CHECK_EQ(povar_cost_homogeneous(h_, &r), POVAR_OK) << povar_last_error(h_);
    fill(ri, r);
  }
  void linearize_pOSE(Scalar alpha) override {
    push_state(POVAR_STATE_POSE, true);
    CHECK_EQ(povar_linearize_pose(h_, alpha), POVAR_OK) << "did not expect numerical failure during linearization";
    after_linearize();
  }
  void linearize_projective_space_homogeneous() override {
    push_state(POVAR_STATE_JOINT, true);
    CHECK_EQ(povar_linearize_homogeneous(h_), POVAR_OK) << "did not expect numerical failure during linearization";
    after_linearize();
  }
  VecX solve(const SolverOptions& /*solver_options*/, Scalar lambda, Scalar /*relative_error_change*/) override {
    return solve_impl(lambda, false, 12);
  }
  VecX solve_joint(Scalar lambda, Scalar /*relative_error_change*/) override { return solve_impl(lambda, true, 11); }
  Scalar apply(const SolverOptions& /*solver_options*/, Scalar alpha, VecX&& /*inc*/) override {
    double l_diff = 0;   // the increment stayed on the device
    CHECK_EQ(povar_apply_pose(h_, alpha, &l_diff), POVAR_OK) << povar_last_error(h_);
    after_apply();
    pull_state(POVAR_STATE_POSE);
    return l_diff;
  }
  Scalar apply_joint(VecX&& /*inc*/) override {
    double l_diff = 0;
    CHECK_EQ(povar_apply_joint(h_, &l_diff), POVAR_OK) << povar_last_error(h_);
    after_apply();
    pull_state(POVAR_STATE_JOINT);
    return l_diff;
  }

 private:
  VecX solve_impl(Scalar lambda, bool joint, int dim) {
    // After a rejected trial the caller restores BalProblem and solves again with a larger damping WITHOUT
    // linearising again (bal_bundle_adjustment.cpp:340-350, 508): the device still holds the trial state, and the
    // matrix-free products rebuild their Jacobian blocks from the state, so the restored host state goes up first.
    push_state(joint ? POVAR_STATE_JOINT : POVAR_STATE_POSE, true);
    VecX inc(dim * static_cast<int>(problem_.cameras().size()));
    int32_t its = 0;
    const int rc = joint ? povar_solve_joint(h_, lambda, inc.data(), &its) : povar_solve_pose(h_, lambda, inc.data(), &its);
    // rc == POVAR_NUM_NONFINITE_INC: inc has NaNs, the caller rejects the step (bal_bundle_adjustment.cpp:362-401)
    CHECK_GE(rc, 0) << povar_last_error(h_);
    povar_phase_times t;
    povar_get_timings(h_, &t);
    if (it_ != nullptr) {
      // the columns LinearizorPowerVarproj::solve / solve_joint fill (linearizor_power_varproj.cpp:139-166, 203-235)
      it_->stage2_time_in_seconds = t.prepare_time;
      it_->prepare_time_in_seconds = t.prepare_time;
      it_->solve_reduced_system_time_in_seconds = t.solve_reduced_system_time;
      it_->linear_solver_iterations = its;
      it_->linear_solver_type = "bal_power_sc";
    }
    if (summary_ != nullptr) summary_->num_linear_solves += 1;
    return inc;
  }
  void after_linearize() {
    povar_phase_times t;
    povar_get_timings(h_, &t);
    if (it_ != nullptr) {   // linearizor_power_varproj.cpp:61-72, 96-106
      it_->jacobian_evaluation_time_in_seconds = t.jacobian_evaluation_time;
      it_->stage1_time_in_seconds = t.jacobian_evaluation_time;
    }
    if (summary_ != nullptr) summary_->num_jacobian_evaluations += 1;
  }
  void after_apply() {
    povar_phase_times t;
    povar_get_timings(h_, &t);
    if (it_ != nullptr) it_->back_substitution_time_in_seconds = t.back_substitution_time;   // :257-306
  }
  // POVAR_PLUGIN_* (see the top of this file) -> communicator descriptor; false: single GPU
  static bool comm_from_environment(povar_comm_desc* comm) {
    const char* world = std::getenv("POVAR_PLUGIN_WORLD");
    if (world == nullptr || std::atoi(world) <= 1) return false;
    std::memset(comm, 0, sizeof(*comm));
    comm->world_size = std::atoi(world);
    const char* rank = std::getenv("POVAR_PLUGIN_RANK");
    const char* dev = std::getenv("POVAR_PLUGIN_DEVICE");
    const char* path = std::getenv("POVAR_PLUGIN_ID_FILE");
    CHECK(rank != nullptr && path != nullptr) << "POVAR_PLUGIN_RANK / POVAR_PLUGIN_ID_FILE missing";
    comm->rank = std::atoi(rank);
    comm->device = dev != nullptr ? std::atoi(dev) : comm->rank;
    // one id per process set: the step-1 and the step-2 linearizor share it (the library caches the communicator)
    static uint8_t id[128];
    static bool have_id = false;
    if (!have_id) {
      if (comm->rank == 0) {
        const char* host = std::getenv("POVAR_PLUGIN_HOST_ID");
        const int rc = (host != nullptr && std::atoi(host) != 0) ? povar_comm_host_id(id) : povar_comm_unique_id(id);
        CHECK_EQ(rc, POVAR_OK) << povar_last_error(nullptr);
        const std::string tmp = std::string(path) + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        CHECK(f != nullptr) << "cannot write " << tmp;
        CHECK_EQ(std::fwrite(id, 1, 128, f), 128u);
        std::fclose(f);
        CHECK_EQ(std::rename(tmp.c_str(), path), 0);
      } else {
        for (int tries = 0;; ++tries) {
          FILE* f = std::fopen(path, "rb");
          if (f != nullptr) {
            const size_t got = std::fread(id, 1, 128, f);
            std::fclose(f);
            if (got == 128) break;
          }
          CHECK_LT(tries, 6000) << "no communicator id in " << path;
          std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
      }
      have_id = true;
    }
    std::memcpy(comm->nccl_id, id, 128);
    return true;
  }
  static void fill(ResidualInfo& ri, const povar_residual_info& r) {
    ri.all.num_obs = static_cast<int>(r.num_obs_all);
    ri.all.error = r.error_all;
    ri.all.residual_sum = r.residual_sum_all;
    ri.valid.num_obs = static_cast<int>(r.num_obs_valid);
    ri.valid.error = r.error_valid;
    ri.valid.residual_sum = r.residual_sum_valid;
    ri.is_numerically_valid = r.is_numerically_valid != 0;
  }
  void pack_cameras(std::vector<double>& cam_P) const {
    const auto& cams = problem_.cameras();
    cam_P.resize(12 * cams.size());
    for (size_t c = 0; c < cams.size(); ++c) {
      for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 4; ++k) cam_P[12 * c + 4 * r + k] = cams[c].space_matrix(r, k);
      }
    }
  }
  // host BalProblem -> device, if the caller changed it since the device state was last seen
  void push_state(int which, bool landmarks) {
    std::vector<double> cam_P, X;
    pack_cameras(cam_P);
    const bool cams_changed = cam_P != dev_P_;
    bool lms_changed = false;
    if (landmarks) {
      const auto& lms = problem_.landmarks();
      const int w = which == POVAR_STATE_JOINT ? 4 : 3;
      X.resize(static_cast<size_t>(w) * (lm_end_ - lm_begin_));
      for (int32_t l = lm_begin_; l < lm_end_; ++l) {
        for (int k = 0; k < w; ++k) {
          X[static_cast<size_t>(w) * (l - lm_begin_) + k] =
              which == POVAR_STATE_JOINT ? lms[l].p_w_homogeneous(k) : lms[l].p_w(k);
        }
      }
      lms_changed = which != dev_which_ || X != dev_X_;
    }
    if (!cams_changed && !lms_changed) return;
    CHECK_EQ(povar_set_state(h_, which, cams_changed ? cam_P.data() : nullptr, lms_changed ? X.data() : nullptr), POVAR_OK)
        << povar_last_error(h_);
    if (cams_changed) dev_P_.swap(cam_P);
    if (lms_changed) {
      dev_X_.swap(X);
      dev_which_ = which;
    }
    ++uploads_;
  }
  // device -> host BalProblem
  void pull_state(int which) {
    auto& cams = problem_.cameras();
    auto& lms = problem_.landmarks();
    const int w = which == POVAR_STATE_JOINT ? 4 : 3;
    std::vector<double> cam_P(12 * cams.size()), X(static_cast<size_t>(w) * (lm_end_ - lm_begin_));
    CHECK_EQ(povar_get_state(h_, which, cam_P.data(), X.data()), POVAR_OK) << povar_last_error(h_);
    for (size_t c = 0; c < cams.size(); ++c) {
      for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 4; ++k) cams[c].space_matrix(r, k) = cam_P[12 * c + 4 * r + k];
      }
    }
    for (int32_t l = lm_begin_; l < lm_end_; ++l) {
      for (int k = 0; k < w; ++k) {
        const double v = X[static_cast<size_t>(w) * (l - lm_begin_) + k];
        if (which == POVAR_STATE_JOINT) lms[l].p_w_homogeneous(k) = v;
        else lms[l].p_w(k) = v;
      }
    }
    dev_P_.swap(cam_P);
    dev_X_.swap(X);
    dev_which_ = which;
  }

  BalProblem<Scalar>& problem_;
  SolverSummary* summary_ = nullptr;
  povar_handle* h_ = nullptr;
  IterationSummary* it_ = nullptr;
  int32_t lm_begin_ = 0, lm_end_ = 0;     // landmarks of this process's shard
  std::vector<double> dev_P_, dev_X_;     // what the device holds (last upload or download)
  int dev_which_ = -1;
  long long uploads_ = 0;
  bool joint_;
  double alpha_;
};

}  // namespace rootba_povar
