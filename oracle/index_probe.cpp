// TEST INFRASTRUCTURE (oracle build only) -- not product code.
//
// Dumps the indexing the reference itself holds for a data_custom file: the file is read by the reference's
// own loader (`load_normalized_bal_problem<double>` -> `BalProblem::load_bal_eccv`,
// /root/reference/src/rootba_povar/bal/bal_problem.cpp:182-303, 873-955), every landmark goes through
// `LandmarkBlockSC::allocate_landmark` (sc/landmark_block.hpp:101-133) and the resulting `pose_idx_`
// (:104-108, read through get_pose_idx(), :342) is printed next to the observation the reference stores
// for that (landmark, camera) pair (`Landmark::obs`, bal/bal_problem.hpp:226) and the camera matrices.
// tools/make_golden.py stores the output as tests/golden/index_<shape>.npz; the CPU and GPU tests compare
// povar_bal_read / povar_canonical_order and the device index against it bit for bit.
//
// usage: index_probe <data_custom file>
// stdout: "C L N", then per landmark "deg cam0 cam1 ...", then N lines "x y" (%.17g) in the same order,
//         then C lines of the 12 entries of space_matrix (row-major, %.17g)
#include <cstdio>
#include <string>
#include <vector>

#include "rootba_povar/bal/bal_dataset_options.hpp"
#include "rootba_povar/bal/bal_problem.hpp"
#include "rootba_povar/sc/landmark_block.hpp"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  using namespace rootba_povar;
  BalDatasetOptions opt;
  opt.input = argv[1];
  opt.quiet = true;
  BalProblem<double> problem = load_normalized_bal_problem<double>(opt, nullptr, nullptr);
  using Block = LandmarkBlockSC<double, 12>;
  Block::Options bopt;
  size_t nobs = 0;
  for (const auto& lm : problem.landmarks()) nobs += lm.obs.size();
  std::printf("%d %d %zu\n", problem.num_cameras(), problem.num_landmarks(), nobs);
  std::vector<Block> blocks(problem.landmarks().size());
  for (size_t l = 0; l < problem.landmarks().size(); ++l) {
    blocks[l].allocate_landmark(problem.landmarks()[l], bopt);
    const std::vector<size_t>& idx = blocks[l].get_pose_idx();
    std::printf("%zu", idx.size());
    for (size_t c : idx) std::printf(" %zu", c);
    std::printf("\n");
  }
  for (size_t l = 0; l < problem.landmarks().size(); ++l) {
    const auto& lm = problem.landmarks()[l];
    for (size_t c : blocks[l].get_pose_idx()) {
      const auto& o = lm.obs.at(static_cast<int>(c));
      std::printf("%.17g %.17g\n", o.pos[0], o.pos[1]);
    }
  }
  for (const auto& cam : problem.cameras()) {
    for (int r = 0; r < 3; ++r) {
      for (int k = 0; k < 4; ++k) std::printf("%.17g ", cam.space_matrix(r, k));
    }
    std::printf("\n");
  }
  return 0;
}
