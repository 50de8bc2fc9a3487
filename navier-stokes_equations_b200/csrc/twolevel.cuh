// Kernels of the two-level velocity preconditioner (stands in, with the polynomial smoother, for the Ifpack ILU(1)
// application on F of the reference, src/classes/NavierStokes.hpp:302-304, 325).
//
// Coarse space = P1 vector functions on the mesh vertices.  Prolongation P: a vertex node of the P2 space takes the coarse
// value of its vertex, a line node the mean of its two end vertices; restriction = P^T.  The coarse operator is the Galerkin
// product F_c = P^T F P, formed cell by cell from what pass 1 of the assembly already wrote: every velocity cell block of
// the linearised system is delta_cd S_ab + gamma G^{cd}_ab (assemble.cuh), hence
//     (P_T^T F_T P_T)[(v,c),(w,d)] = delta_cd (P_s^T S P_s)_vw + gamma |T| d_c lambda_v d_d lambda_w
// with P_s the scalar 10x4 (6x3) cell prolongation -- the second term because P1 gradients are the prolonged P2 gradients.
// A fine DoF that carries a Dirichlet condition lies on an edge / at a vertex whose end vertices are all constrained, so
// the product restricted to the free coarse DoFs needs no masks; constrained coarse DoFs get identity rows and columns.
#pragma once
#include "assemble.cuh"
#include "velstream.cuh"

namespace nsb {

constexpr int CG_WARPS = 4;
constexpr int CG_MAX_NB = 64;       // coarse neighbours (P1 stencil) a row kernel can accumulate

// One warp per OWNED vertex: coarse block row of F_c in fp64, blocks in the order of the vertex's P1 neighbour list
// (cvals[(cnbr_ptr[P] + k) * dim*dim + c*dim + d]), and the inverse of its diagonal block.
template <int DIM>
__global__ void __launch_bounds__(CG_WARPS * 32)
k_coarse_rows(DevMesh M, const double* __restrict__ ctx, double gamma, double wsum, const long long* __restrict__ cnbr_ptr,
              const int* __restrict__ cnbr_vxoff, const unsigned char* __restrict__ cflag, double* __restrict__ cvals,
              double* __restrict__ cdinv) {
  constexpr int NV = DIM + 1, NN = Fe<DIM>::NN, PL = DIM * DIM;
  using CL = CtxL<DIM>;
  __shared__ double sacc[CG_WARPS][CG_MAX_NB * PL];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int P = blockIdx.x * CG_WARPS + wid;
  if (P >= M.np_own) return;
  double* acc = sacc[wid];
  const int A = M.pid_node[P];
  const long long q0 = cnbr_ptr[P];
  const int nbc = (int)(cnbr_ptr[P + 1] - q0);
  for (int k = lane; k < nbc * PL; k += 32) acc[k] = 0.0;
  __syncwarp();
  const long long kc0 = M.n2c_ptr[A];
  const int ncell = (int)(M.n2c_ptr[A + 1] - kc0);
  for (int ic = 0; ic < ncell; ++ic) {
    const uint32_t pk = __ldg(M.n2c + kc0 + ic);
    const size_t cell = pk >> 4;
    const int a = (int)(pk & 15u);                   // local index of the vertex in this cell (< NV)
    const double* cx = ctx + cell * CL::N;
    // t_b = S[a][b] + 1/2 sum over the line nodes e at a of S[e][b]        (row a of P_s^T S)
    double t = 0.0;
    if (lane < NN) {
      t = cx[CL::S + a * NN + lane];
#pragma unroll
      for (int e = NV; e < NN; ++e)
        if (node_i<DIM>(e) == a || node_j<DIM>(e) == a) t += 0.5 * cx[CL::S + e * NN + lane];
    }
    // (P_s^T S P_s)_aj = t_j + 1/2 sum over the line nodes f at j of t_f
    double s = 0.0;
#pragma unroll
    for (int f = 0; f < NN; ++f) {
      const double tf = __shfl_sync(NSB_FULL, t, f);
      if (lane < NV) {
        if (f == lane) s += tf;
        else if (f >= NV && (node_i<DIM>(f) == lane || node_j<DIM>(f) == lane)) s += 0.5 * tf;
      }
    }
    if (lane < NV) {
      const int j = lane;
      const int rp = M.rank_up[(cell * NN + a) * NV + j];
      const double vol = gamma * wsum * cx[CL::ABSJ];
      double ga[DIM], gj[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) { ga[k] = cx[CL::GL + a * DIM + k]; gj[k] = cx[CL::GL + j * DIM + k]; }
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int d = 0; d < DIM; ++d) acc[rp * PL + c * DIM + d] += ((c == d) ? s : 0.0) + vol * ga[c] * gj[d];
    }
    __syncwarp();
  }
  // Dirichlet rows / columns -> identity, store, invert the diagonal block
  const int self = M.pselfrank[P];
  const int xA = DIM * A;
  for (int k = lane; k < nbc * PL; k += 32) {
    const int Q = k / PL, c = (k % PL) / DIM, d = k % DIM;
    const bool crow = cflag[xA + c] != 0;
    const bool ccol = cflag[__ldg(cnbr_vxoff + q0 + Q) + d] != 0;
    double v = acc[k];
    if (crow || ccol) v = (Q == self && c == d && crow) ? 1.0 : 0.0;
    acc[k] = v;
    cvals[(q0 + Q) * PL + k % PL] = v;
  }
  __syncwarp();
  if (lane == 0) {
    double Dm[DIM][DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c)
#pragma unroll
      for (int e = 0; e < DIM; ++e) Dm[c][e] = acc[self * PL + c * DIM + e];
    double* o = cdinv + (size_t)P * PL;
    if (DIM == 2) {
      const double id = 1.0 / (Dm[0][0] * Dm[1][1] - Dm[0][1] * Dm[1][0]);
      o[0] = Dm[1][1] * id; o[1] = -Dm[0][1] * id; o[2] = -Dm[1][0] * id; o[3] = Dm[0][0] * id;
    } else {
      const double c00 = Dm[1][1] * Dm[2][2] - Dm[1][2] * Dm[2][1];
      const double c01 = Dm[1][2] * Dm[2][0] - Dm[1][0] * Dm[2][2];
      const double c02 = Dm[1][0] * Dm[2][1] - Dm[1][1] * Dm[2][0];
      const double id = 1.0 / (Dm[0][0] * c00 + Dm[0][1] * c01 + Dm[0][2] * c02);
      o[0] = c00 * id;
      o[1] = (Dm[0][2] * Dm[2][1] - Dm[0][1] * Dm[2][2]) * id;
      o[2] = (Dm[0][1] * Dm[1][2] - Dm[0][2] * Dm[1][1]) * id;
      o[3] = c01 * id;
      o[4] = (Dm[0][0] * Dm[2][2] - Dm[0][2] * Dm[2][0]) * id;
      o[5] = (Dm[0][2] * Dm[1][0] - Dm[0][0] * Dm[1][2]) * id;
      o[6] = c02 * id;
      o[7] = (Dm[0][1] * Dm[2][0] - Dm[0][0] * Dm[2][1]) * id;
      o[8] = (Dm[0][0] * Dm[1][1] - Dm[0][1] * Dm[1][0]) * id;
    }
  }
}

// Packs B_c = Dinv_c F_c into the coarse level's tile-planar layout (one CTA per tile), cf. k_vel_pack.
template <int DIM, typename VT>
__global__ void __launch_bounds__(256)
k_coarse_pack(const VsTile* __restrict__ tiles, const uint32_t* __restrict__ meta, const long long* __restrict__ cnbr_ptr,
              const double* __restrict__ cvals, const double* __restrict__ cdinv, VT* __restrict__ out) {
  constexpr int PL = DIM * DIM;
  const VsTile hd = tiles[blockIdx.x];
  VT* o = out + (long long)PL * 4 * hd.nq_off;
  const uint4* mt = reinterpret_cast<const uint4*>(meta) + hd.nq_off;
  for (int jb = threadIdx.x; jb < 4 * hd.NQ; jb += blockDim.x) {
    const int q = jb >> 2, e = jb & 3;
    double b[DIM][DIM];
#pragma unroll
    for (int r = 0; r < DIM; ++r)
#pragma unroll
      for (int c = 0; c < DIM; ++c) b[r][c] = 0.0;
    const uint4 m = __ldg(mt + q);
    if (m.z != 0xffffu && e < (int)(m.w >> 16)) {
      const int P = hd.n0 + (int)m.z;
      const long long blk = cnbr_ptr[P] + (int)(m.w & 0xffffu) + e;
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int c = 0; c < DIM; ++c)
#pragma unroll
          for (int ee = 0; ee < DIM; ++ee) b[r][c] += __ldg(cdinv + (size_t)P * PL + r * DIM + ee) * __ldg(cvals + blk * PL + ee * DIM + c);
    }
#pragma unroll
    for (int r = 0; r < DIM; ++r)
#pragma unroll
      for (int c = 0; c < DIM; ++c) o[((size_t)(r * DIM + c) * hd.NQ + q) * 4 + e] = vs_from_double<DIM, VT>(b[r][c]);
  }
}

// r_c = P^T r on the owned vertices (r needs valid ghost velocity entries); constrained coarse DoFs get 0
template <int DIM>
__global__ void k_restrict(int np_own, const int* __restrict__ pid_node, const long long* __restrict__ vedge_ptr,
                           const int* __restrict__ vedge_xoff, const unsigned char* __restrict__ cflag,
                           const double* __restrict__ r, double* __restrict__ rc) {
  const int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= np_own) return;
  const int xA = DIM * __ldg(pid_node + P);
  double s[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) s[c] = r[xA + c];
  for (long long k = vedge_ptr[P]; k < vedge_ptr[P + 1]; ++k) {
    const int xo = __ldg(vedge_xoff + k);
#pragma unroll
    for (int c = 0; c < DIM; ++c) s[c] += 0.5 * r[xo + c];
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) rc[DIM * P + c] = cflag[xA + c] ? 0.0 : s[c];
}

// y = P e_c on the owned fine nodes (e_c needs valid ghost vertices); constrained fine DoFs get 0
template <int DIM>
__global__ void k_prolong(int nn_own, const int* __restrict__ ends_xoff, const unsigned char* __restrict__ cflag,
                          const double* __restrict__ ec, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn_own * DIM) return;
  const int A = i / DIM, c = i % DIM;
  const int e0 = __ldg(ends_xoff + 2 * A), e1 = __ldg(ends_xoff + 2 * A + 1);
  y[i] = cflag[i] ? 0.0 : 0.5 * (ec[e0 + c] + ec[e1 + c]);
}

}  // namespace nsb
