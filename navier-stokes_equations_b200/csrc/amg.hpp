// Smoothed-aggregation multigrid hierarchy for the constant pressure Laplacian K_p,
// built once on the host and applied on the device as one V-cycle per preconditioner
// application.  Stands in for Trilinos ML (`PreconditionAMG`, elliptic, 1 V-cycle,
// Chebyshev degree-2 smoother, aggregation threshold 0.02 -- reference
// src/classes/NavierStokes.hpp:310-315, 338), which the reference rebuilds at every solve
// although K_p never changes.
#pragma once
#include <cstdint>
#include <vector>

namespace nsb {

struct HostCsr {
  int n = 0, m = 0;                       // rows, cols
  std::vector<int> ptr, col;
  std::vector<double> val;
  int64_t nnz() const { return (int64_t)col.size(); }
};

struct AmgLevel {
  HostCsr A, P, R;                        // operator, prolongation to this level from the next, R = P^T
  std::vector<double> dinv;               // 1/diag(A)
  double lmax = 0;                        // estimate of lambda_max(D^-1 A)
};

struct AmgHierarchy {
  std::vector<AmgLevel> levels;           // levels[0] is the fine operator; last level is solved densely
  std::vector<double> coarse_inv;         // dense inverse of the coarsest operator (row-major)
};

void amg_setup(const HostCsr& A, AmgHierarchy& H, double threshold = 0.02, int max_coarse = 1000, int max_levels = 12);

double power_lmax_jacobi(const HostCsr& A, const std::vector<double>& dinv, int iters = 30);
HostCsr transpose(const HostCsr& A);
HostCsr spgemm(const HostCsr& A, const HostCsr& B);

}  // namespace nsb
