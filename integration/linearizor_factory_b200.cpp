// Replacement for the reference's factory TU (/root/reference/src/rootba_povar/solver/linearizor.cpp:47-79):
// every solver type of both steps is served by LinearizorB200 (the library reads the solver types from the
// options it is created with).  Compiled instead of solver/linearizor.cpp by `make -C oracle plugin`.
#include "linearizor_b200.hpp"

namespace rootba_povar {

template <typename Scalar_>
std::unique_ptr<Linearizor<Scalar_>> Linearizor<Scalar_>::create(BalProblem<Scalar>& bal_problem,
                                                                 const SolverOptions& options,
                                                                 SolverSummary* summary) {
  return std::make_unique<LinearizorB200<Scalar_>>(bal_problem, options, summary, /*joint=*/false);
}

template <typename Scalar_>
std::unique_ptr<Linearizor<Scalar_>> Linearizor<Scalar_>::create_homogeneous(BalProblem<Scalar>& bal_problem,
                                                                             const SolverOptions& options,
                                                                             SolverSummary* summary) {
  return std::make_unique<LinearizorB200<Scalar_>>(bal_problem, options, summary, /*joint=*/true);
}

template class Linearizor<double>;

}  // namespace rootba_povar
