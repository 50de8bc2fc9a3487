// Common device-side declarations for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "fe_tables.hpp"

namespace nsb {

#define NSB_FULL 0xffffffffu

// Per owned node, everything a row kernel needs, in one 32-byte record (two 128-bit loads).
struct __align__(16) NodeDesc {
  long long rowbase;      // offset of scalar row (node,0) in the value array
  long long prowbase;     // offset of the node's pressure row (vertex nodes), else 0
  int nbr0;               // first entry of the node's neighbour list in nbr_xoff
  int pnbr0;              // first entry in pnbr_xoff
  unsigned short nb, np;  // neighbour counts (velocity nodes / pressure DoFs)
  int pid;                // local pressure id of a vertex node, -1 otherwise
};

// Everything the kernels need to know about the (local) mesh + sparse structure.
struct DevMesh {
  const NodeDesc* nd;                   // [nn_own]
  int dim, nn_own, nn_tot, np_own, np_tot, nc;
  long long n_own, n_tot;               // local vector lengths (owned / owned+ghost)
  // cells
  const int* cell_xoff;                 // [nc][NN]  offset of (node,0) in a local vector
  const int* cell_poff;                 // [nc][NV]  offset of the vertex pressure DoF
  const double* cell_geom;              // [nc][16]
  const uint16_t* rank_uu;              // [nc][NN][NN]
  const uint16_t* rank_up;              // [nc][NN][NV]
  // owned nodes
  const long long* nbr_ptr;             // [nn_own+1]
  const int* nbr_xoff;                  // x-offset of neighbour node's first component
  const long long* pnbr_ptr;            // [nn_own+1]
  const int* pnbr_xoff;                 // x-offset of neighbour pressure DoF
  const int* selfrank;                  // [nn_own]
  const int* node_pid;                  // [nn_own] local pressure id or -1
  const int* pid_node;                  // [np_own]
  const int* pselfrank;                 // [np_own]
  const long long* rowbase;             // [nn_own]
  const long long* prowbase;            // [np_own]
  const long long* n2c_ptr;             // [nn_own+1]
  const uint32_t* n2c;
};

// reference NavierStokes.hpp:485-511 / cpp:660-676: everything assembly needs per call
struct AsmParams {
  double dt, theta, nu, rho, gamma;     // gamma = 0.1 grad-div weight (cpp:463,793), 0 when !use_supg
  double inv_dt;                        // 1/dt, computed once on the host (fp64 division is ~25 instructions on the device)
  int use_supg;
  int first_order_ustar;                // first_step || second_step || BackwardEuler (cpp:665)
};


// Context written per cell by the first pass and consumed by the node-row pass (Newton system: per-point data).
template <int DIM> struct Ctx {
  static constexpr int NV = DIM + 1;
  static constexpr int NQ = Fe<DIM>::NQ;
  static constexpr int GL = 0;                    // grad lambda [NV][DIM]
  static constexpr int ABSJ = NV * DIM;           // |det J|
  static constexpr int AVG = ABSJ + 1;            // average |local diagonal| (A.5 fallback)
  static constexpr int S = ABSJ + 2;              // s[q][k] = u*(q) . grad lambda_k
  static constexpr int TW = S + NQ * NV;          // tau(q) * JxW(q)   (0 without SUPG)
  static constexpr int H = TW + NQ;               // extra (Newton): grad u^k(q) [NQ][DIM][DIM]
  static constexpr int N_NEWTON = ((H + NQ * DIM * DIM + 7) / 8) * 8;
};

// Context of the LINEARISED system: the quadrature sums are finished per cell in pass 1, pass 2 only reads
//   S_ab = |J| Mhat_ab / dt + theta nu tr G_ab + sum_q [Pa cb + Qa (phi_b / dt + cb)]   (the delta_cd part of the block)
//   CA_a = sum_q tau JxW (u* . grad phi_a)                                             (SUPG pressure column)
template <int DIM> struct CtxL {
  static constexpr int NV = DIM + 1;
  static constexpr int NN = Fe<DIM>::NN;
  static constexpr int GL = 0;                          // grad lambda [NV][DIM]
  static constexpr int ABSJ = NV * DIM;
  static constexpr int AVG = ABSJ + 1;
  static constexpr int HDR = ((AVG + 1 + 7) / 8) * 8;   // 16 (3-D), 8 (2-D)
  static constexpr int S = HDR;                         // S[a][b]
  static constexpr int CA = S + NN * NN;
  static constexpr int N = ((CA + NN + 7) / 8) * 8;     // 128 (3-D), 56 (2-D)
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(NSB_FULL, v, src); }

__device__ __forceinline__ double warp_sum_fixed(double v) {
  // fixed butterfly order -> bit-reproducible
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NSB_FULL, v, o);
  return v;
}

}  // namespace nsb
