// TEST INFRASTRUCTURE (oracle build only) -- not product code.
//
// Calls the reference's own `BalBundleAdjustmentHelper<double>::kernel_COD`
// (/root/reference/src/rootba_povar/bal/bal_bundle_adjustment_helper.cpp:201-216,
// linked from the object compiled out of the unmodified reference source) on
// row vectors read from stdin and prints the returned kernel matrices, so that
// tools/make_golden.py can store them as golden vectors (tests/golden/kernel_cod.npz).
//
// stdin : n count, then count*n doubles
// stdout: count blocks of n*(cols) doubles (row-major), preceded by "rows cols"
#include <cstdio>
#include <vector>

#include <Eigen/Dense>

#include "rootba_povar/bal/bal_bundle_adjustment_helper.hpp"

int main() {
  int n = 0;
  int count = 0;
  if (std::scanf("%d %d", &n, &count) != 2) return 1;
  for (int t = 0; t < count; ++t) {
    Eigen::MatrixXd m(1, n);
    for (int j = 0; j < n; ++j) {
      double v = 0;
      if (std::scanf("%lf", &v) != 1) return 1;
      m(0, j) = v;
    }
    Eigen::MatrixXd k = rootba_povar::BalBundleAdjustmentHelper<double>::kernel_COD(m);
    std::printf("%d %d\n", static_cast<int>(k.rows()), static_cast<int>(k.cols()));
    for (int r = 0; r < k.rows(); ++r) {
      for (int c = 0; c < k.cols(); ++c) std::printf("%.17g ", k(r, c));
      std::printf("\n");
    }
  }
  return 0;
}
