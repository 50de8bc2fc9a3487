// Element-wise ("unassembled") application of the velocity block F of the LINEARISED system, used
// inside the velocity polynomial of the block preconditioner (the stand-in for Ifpack ILU(1),
// reference NavierStokes.hpp:302-304, 325).
//
// For the linearised system every velocity-velocity cell block has the form
//     A_e[(a,c),(b,d)] = delta_cd S_ab + gamma |J| sum_{sa,sb} g_{i(a,sa),c} Khat[a][b][sa][sb] g_{i(b,sb),d}
// (reference cpp:744-793: mass, viscosity, convection and the SUPG velocity terms are all
// delta_cd; only grad-div couples components), so F x needs, per (owned node, cell) pair, the row
// S_e[a][0..NN) -- NN floats written by the assembly -- instead of the dim*dim*NN assembled
// values the row kernel streams: 2.4x fewer bytes per application on tets.  The grad-div part is
// rebuilt from the cell geometry (L2-resident) and the reference tensor Khat.
//
// One CTA per SpMV tile.  x of the tile's unique neighbour nodes is staged in shared memory
// (zeroed at Dirichlet DoFs = column elimination); one THREAD per (node, cell) pair computes the
// pair's dim outputs in fp64; the pairs of a node are consecutive, so a second phase sums them in
// list order (deterministic, no atomics) and applies the same fused epilogues as k_spmv_vel_f32.
// Dirichlet rows act as the identity after the block-Jacobi scaling (their node block is
// decoupled), exactly like the assembled operator.
#pragma once
#include "assemble.cuh"
#include "linalg.cuh"

namespace nsb {

#ifndef NSB_EBE_THREADS
#define NSB_EBE_THREADS 256
#endif
#ifndef NSB_EBE_MIN_CTAS
#define NSB_EBE_MIN_CTAS 3
#endif
constexpr int EBE_THREADS = NSB_EBE_THREADS;

struct EbeData {
  const float* s_rows;               // S_e[a][b] of pair p at ebe_index(p, b)  (blocked by 32 pairs, see below)
  const unsigned short* pair_loc;    // same blocking: position of cell node b in the tile's unique list
  const unsigned short* pair_ca;     // [NP] (position of the pair's cell in the tile's unique cell list) << 4 | local node a
  const int* tile_cell_ptr;          // [n_tiles+1] into tile_cells
  const int* tile_cells;             // unique local cells touched by the tile's pairs
  const unsigned char* cflag;        // Dirichlet flag per local DoF
  double gamma;                      // grad-div weight (0 without SUPG)
  int ypair_doubles;                 // size of the pair-result region of the dynamic shared memory
};

// index of (pair p, cell node b) in the blocked pair arrays: 32 pairs x 2 consecutive b per 256-byte (floats) /
// 128-byte (positions) warp row, so a lane fetches two b with one 8- / 4-byte load
template <int NN> __host__ __device__ inline size_t ebe_index(long long p, int b) {
  return ((size_t)(p >> 5) * (NN / 2) + (size_t)(b >> 1)) * 64 + (size_t)(p & 31) * 2 + (size_t)(b & 1);
}

template <int DIM, int MODE>
__global__ void __launch_bounds__(EBE_THREADS, NSB_EBE_MIN_CTAS)
k_apply_F_ebe(DevMesh M, SpmvTiles TL, EbeData E, const FeTables* __restrict__ gT, const double* __restrict__ x,
              double* __restrict__ y, const double* __restrict__ u, double* __restrict__ poly,
              const double* __restrict__ dinv, PolyCoef pc) {
  constexpr int NV = DIM + 1, NN = Fe<DIM>::NN, NH = NN / 2;
  constexpr int GEO_N = NV * DIM + 1;               // grad lambda [NV][DIM], |J|   (odd: spreads the banks)
  static_assert(NN % 2 == 0, "pair arrays hold two cell nodes per element");
  static_assert(TILE_MAX_NODES * 4 <= EBE_THREADS, "one thread per (node, component slot) in the row phase");
  extern __shared__ double ypair[];                 // [pairs of the tile][DIM], then the geometry of the tile's cells
  double* geo_s = ypair + E.ypair_doubles;
  __shared__ double xs[TILE_MAX_UNIQ * DIM];
  __shared__ int s_ia[NN][2];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int n0 = TL.node_ptr[t], n1 = TL.node_ptr[t + 1], nn = n1 - n0;
  const long long p0 = M.n2c_ptr[n0];
  const int npairs = (int)(M.n2c_ptr[n1] - p0);

  // ---- every load that does not depend on staged data is issued first, so that its latency overlaps the staging
  // (a) operands of the row phase: thread (s, c) = (tid / 4, tid % 4) owns row (n0 + s, c)
  const int rs = tid >> 2, rc = tid & 3;
  const bool has_row = rs < nn && rc < DIM;
  int q0 = 0, q1 = 0;
  double dv[DIM], uv = 0.0, pv = 0.0, xown = 0.0;
  bool crow = false;
#pragma unroll
  for (int e = 0; e < DIM; ++e) dv[e] = 0.0;
  if (has_row) {
    const int row = DIM * (n0 + rs) + rc;
    q0 = (int)(M.n2c_ptr[n0 + rs] - p0);
    q1 = (int)(M.n2c_ptr[n0 + rs + 1] - p0);
#pragma unroll
    for (int e = 0; e < DIM; ++e) dv[e] = __ldg(dinv + (size_t)(n0 + rs) * DIM * DIM + rc * DIM + e);
    crow = E.cflag[row] != 0;
    xown = x[row];
    if (MODE == 3) { uv = u[row]; pv = poly[row]; }
  }
  // (b) the first round of pair data
  float2 Sv[NH];
  unsigned lc[NH];
  unsigned ca = 0;
  auto load_pair = [&](int i) {
    const long long p = p0 + i;
    const size_t base = ebe_index<NN>(p, 0);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      Sv[h] = __ldcs(reinterpret_cast<const float2*>(E.s_rows + base + (size_t)h * 64));
      lc[h] = __ldcs(reinterpret_cast<const unsigned*>(E.pair_loc + base + (size_t)h * 64));
    }
    ca = __ldg(E.pair_ca + p);
  };
  if (tid < npairs) load_pair(tid);

  // ---- staging: x of the unique neighbour nodes (Dirichlet columns zeroed), Khat, the cells' geometry
  const int u0 = TL.uniq_ptr[t], nuq = TL.uniq_ptr[t + 1] - u0;
  for (int i = tid; i < nuq * DIM; i += EBE_THREADS) {
    const int xo = __ldg(TL.uniq_xoff + u0 + i / DIM) + i % DIM;
    xs[i] = E.cflag[xo] ? 0.0 : __ldg(x + xo);
  }
  if (tid < NN * 2) s_ia[tid / 2][tid % 2] = gT->idx[tid / 2][tid % 2];
  // per-thread loads of the 128-byte cell records would be 32 sectors per request (the L1 tag stage becomes
  // the limit), so the geometry of the tile's unique cells is staged once with coalesced loads
  const int c0 = E.tile_cell_ptr[t], nuc = E.tile_cell_ptr[t + 1] - c0;
  for (int i = tid; i < nuc * GEO_N; i += EBE_THREADS) {
    const int k = i % GEO_N;
    geo_s[i] = __ldg(M.cell_geom + (size_t)__ldg(E.tile_cells + c0 + i / GEO_N) * 16 + (k < NV * DIM ? k : 12));
  }
  __syncthreads();

  // ---- phase 1: one thread per (node, cell) pair
  for (int i = tid; i < npairs; i += EBE_THREADS) {
    if (i != tid) load_pair(i);
    const int a = (int)(ca & 15u);
    const double* geo = geo_s + (int)(ca >> 4) * GEO_N;
    double g[NV][DIM];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < DIM; ++k) g[v][k] = geo[v * DIM + k];
    const double ga = E.gamma * geo[NV * DIM];
    double yv[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) yv[c] = 0.0;
    // grad-div part without the Khat table: div x_h is P1 on the cell, D(lambda) = sum_k Dk lambda_k with
    //   Dk = 4 (z_k + sum over edges of the slot that differentiates w.r.t. lambda_k) - sum_vertices z_i,
    // z_b^s = g_{i(b,s)} . x_b, and  int dN_a/dlambda D  follows from  int lambda_i lambda_k = m2 (1 + delta_ik),
    // int lambda_k = m1  (products of linear functions: the quadrature rule of the assembly integrates them exactly)
    double Dk[NV], Z = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) Dk[k] = 0.0;
#pragma unroll
    for (int b = 0; b < NN; ++b) {
      const int loc = (b & 1) ? (int)(lc[b >> 1] >> 16) : (int)(lc[b >> 1] & 0xffffu);
      const double* xb = xs + loc * DIM;
      double xv[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) xv[d] = xb[d];
      const double s = (double)((b & 1) ? Sv[b >> 1].y : Sv[b >> 1].x);
#pragma unroll
      for (int d = 0; d < DIM; ++d) yv[d] += s * xv[d];
      const int ib0 = node_i<DIM>(b), ib1 = node_j<DIM>(b);
      double z0 = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) z0 += g[ib0][d] * xv[d];
      if (b < NV) {
        Dk[ib0] += z0;                                 // d/dlambda_i of a vertex function: 4 lambda_i - 1
        Z += z0;
      } else {
        double z1 = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) z1 += g[ib1][d] * xv[d];
        Dk[ib1] += z0;                                 // d/dlambda_i of 4 lambda_i lambda_j is 4 lambda_j
        Dk[ib0] += z1;
      }
    }
    constexpr double REFV = (DIM == 2) ? 0.5 : 1.0 / 6.0;
    constexpr double m2 = REFV / ((DIM + 1) * (DIM + 2)), m1 = REFV / (DIM + 1);
    double SD = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) { Dk[k] = 4.0 * Dk[k] - Z; SD += Dk[k]; }
    const int ia0 = s_ia[a][0], ia1 = s_ia[a][1];
    double Mi = 0.0, Mj = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double Mk = m2 * (SD + Dk[k]);
      if (k == ia0) Mi = Mk;
      if (k == ia1) Mj = Mk;
    }
    const bool avert = a < NV;
    const double w0 = ga * (avert ? 4.0 * Mi - m1 * SD : 4.0 * Mj);
    const double w1 = avert ? 0.0 : ga * 4.0 * Mi;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      yv[c] += geo[ia0 * DIM + c] * w0 + geo[ia1 * DIM + c] * w1;
      ypair[i * DIM + c] = yv[c];
    }
  }
  __syncthreads();

  // ---- phase 2: ordered sum over the pairs of each row, block-Jacobi scaling and the polynomial step (same
  // algebra as vel_epilogue).  The DIM rows of a node sit in adjacent lanes of one 4-lane group.
  double sum = 0.0;
  for (int q = q0; q < q1; ++q) sum += ypair[q * DIM + (rc < DIM ? rc : 0)];
  double tt = 0.0;
#pragma unroll
  for (int e = 0; e < DIM; ++e) tt += dv[e] * __shfl_sync(NSB_FULL, sum, (tid & 28) + e);
  if (has_row) {
    const int row = DIM * (n0 + rs) + rc;
    if (crow) tt = xown;                             // Dirichlet row: Dinv F acts as the identity
    if (MODE == 2) {
      y[row] = tt;
    } else {
      const double yn = pc.cu * uv + pc.ct * tt;
      y[row] = yn;
      poly[row] = pv + pc.cpu * uv + pc.cpy * yn;
    }
  }
}

}  // namespace nsb
