// Bandwidth-bound kernels of the Krylov solve: node-block SpMV (and its fused
// Chebyshev / Schur variants), generic CSR kernels for the pressure multigrid, fused
// multi-dot / multi-axpy for Gram-Schmidt, and small BLAS-1.
//
// Replaces the Epetra/Trilinos primitives behind
//   SolverGMRES::solve(system_matrix, x, rhs, preconditioner)        reference NavierStokes.cpp:561, 853
//   TrilinosWrappers::BlockSparseMatrix::vmult, B->vmult, sadd, add  reference NavierStokes.hpp:334-343
// All reductions use fixed trees -> results are bit-reproducible run to run.
#pragma once
#include "device.cuh"

#ifndef NSB_STREAM_LOAD
#define NSB_STREAM_LOAD __ldcs
#endif

namespace nsb {

#ifndef NSB_SPMV_WARPS
#define NSB_SPMV_WARPS 8
#endif
constexpr int SPMV_WARPS = NSB_SPMV_WARPS;

// ------------------------------------------------------------------------------------
// Row kernels over the node-block structure.  One warp per owned P2 node handles the node's
// dim velocity rows (+ its pressure row when the node is a vertex): the rows share the
// column set, so x is gathered once and every matrix value is streamed exactly once,
// contiguously per row (coalesced warp loads, evict-first).  Columns are processed in
// chunks of 32*SPMV_UNROLL with all loads of a chunk issued before the first use, which is
// what keeps enough bytes in flight (the kernel is latency-, not instruction-bound).
// ------------------------------------------------------------------------------------
constexpr int SPMV_UNROLL = 4;

__device__ __forceinline__ NodeDesc load_desc(const NodeDesc* p) {
  const int4* q = reinterpret_cast<const int4*>(p);
  const int4 a = __ldg(q), b = __ldg(q + 1);
  NodeDesc d;
  d.rowbase = ((long long)(unsigned)a.y << 32) | (unsigned)a.x;
  d.prowbase = ((long long)(unsigned)a.w << 32) | (unsigned)a.z;
  d.nbr0 = b.x; d.pnbr0 = b.y;
  d.nb = (unsigned short)(b.z & 0xffff); d.np = (unsigned short)((unsigned)b.z >> 16);
  d.pid = b.w;
  return d;
}

// sum[r] += sum_k row_r[k] * x[col(k)] for k in [0, ncols); rows r < ROWS at rowptr[r].
// Columns k < nbd are velocity columns (DIM per neighbour node), the rest pressure columns.
template <int DIM, int ROWS, typename VT>
__device__ __forceinline__ void row_block_dot(const VT* const (&rowp)[ROWS], int ncols, int nbd, const int* __restrict__ nx,
                                              const int* __restrict__ px, const double* __restrict__ x, int lane,
                                              double (&sum)[ROWS]) {
  for (int k0 = 0; k0 < ncols; k0 += 32 * SPMV_UNROLL) {
    VT v[ROWS][SPMV_UNROLL];
    int xo[SPMV_UNROLL];
    double xv[SPMV_UNROLL];
#pragma unroll
    for (int u = 0; u < SPMV_UNROLL; ++u) {
      const int k = k0 + 32 * u + lane;
      const bool ok = k < ncols;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) v[r][u] = ok ? __ldcs(rowp[r] + k) : VT(0);
      xo[u] = -1;
      if (ok) xo[u] = (k < nbd) ? (nx[k / DIM] + k % DIM) : px[k - nbd];
    }
#pragma unroll
    for (int u = 0; u < SPMV_UNROLL; ++u) xv[u] = xo[u] >= 0 ? __ldg(x + xo[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < SPMV_UNROLL; ++u)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) sum[r] += (double)v[r][u] * xv[u];
  }
}

// A tile = a run of consecutive owned nodes.  The host precomputes, per tile, the UNIQUE
// neighbour nodes / pressure DoFs it touches and rewrites every neighbour reference as a 16-bit
// position in that list.  The CTA stages descriptors, those positions and the x-values of the
// unique neighbours in shared memory with coalesced loads (each x entry is read once per tile
// instead of once per row that references it), then its warps pull nodes off a shared counter
// and only stream matrix values; the x "gather" happens in shared memory.  This is what takes
// the scattered 8-byte gathers -- which saturated L1/TEX at ~12 wavefronts per request -- out
// of the inner loop.
#ifndef NSB_TILE_NODES
#define NSB_TILE_NODES 64
#endif
constexpr int TILE_MAX_NODES = NSB_TILE_NODES;
constexpr int TILE_MAX_IDX = 3072;       // staged 16-bit neighbour positions (velocity + pressure)
constexpr int TILE_MAX_UNIQ = 768;       // unique neighbour nodes per tile
constexpr int TILE_MAX_PUNIQ = 256;      // unique neighbour pressure DoFs per tile

struct SpmvTiles {
  const int* node_ptr;                   // [n_tiles+1] node ranges
  const int* uniq_ptr;                   // [n_tiles+1] into uniq_xoff
  const int* uniq_xoff;                  // x offset of (unique neighbour node, 0)
  const int* puniq_ptr;                  // [n_tiles+1] into puniq_xoff
  const int* puniq_xoff;                 // x offset of unique neighbour pressure DoFs
  const unsigned short* nbr_loc;         // parallel to nbr_xoff: position in the tile's unique list
  const unsigned short* pnbr_loc;        // parallel to pnbr_xoff
};

template <int DIM, typename XT = double> struct TileSmem {
  int4 desc[TILE_MAX_NODES * 2];
  XT xs[TILE_MAX_UNIQ * DIM + TILE_MAX_PUNIQ];
  unsigned short idx[TILE_MAX_IDX];
  int next, base_n, cnt_n, base_p, nuq;
};

template <int DIM, bool WITH_P, typename XT = double>
__device__ __forceinline__ void stage_tile(const DevMesh& M, const SpmvTiles& TL, int t, const double* __restrict__ x,
                                           TileSmem<DIM, XT>& T, int& n0, int& n1) {
  n0 = TL.node_ptr[t]; n1 = TL.node_ptr[t + 1];
  const int nn = n1 - n0;
  const int4* g = reinterpret_cast<const int4*>(M.nd + n0);
  for (int i = threadIdx.x; i < 2 * nn; i += blockDim.x) T.desc[i] = __ldg(g + i);
  // x values of the unique neighbours (independent of the descriptors)
  const int u0 = TL.uniq_ptr[t], nuq = TL.uniq_ptr[t + 1] - u0;
  for (int i = threadIdx.x; i < nuq * DIM; i += blockDim.x) T.xs[i] = (XT)__ldg(x + __ldg(TL.uniq_xoff + u0 + i / DIM) + i % DIM);
  if (WITH_P) {
    const int p0 = TL.puniq_ptr[t], npu = TL.puniq_ptr[t + 1] - p0;
    for (int i = threadIdx.x; i < npu; i += blockDim.x) T.xs[nuq * DIM + i] = (XT)__ldg(x + __ldg(TL.puniq_xoff + p0 + i));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int4 f = T.desc[1], l = T.desc[2 * nn - 1];
    T.base_n = f.x;
    T.cnt_n = l.x + (l.z & 0xffff) - f.x;
    T.base_p = f.y;
    T.nuq = nuq;
    T.next = 0;
  }
  __syncthreads();
  const int base_n = T.base_n, cnt_n = T.cnt_n;
  for (int i = threadIdx.x; i < cnt_n; i += blockDim.x) T.idx[i] = __ldg(TL.nbr_loc + base_n + i);
  if (WITH_P) {
    const int4 l = T.desc[2 * nn - 1];
    const int cnt_p = l.y + (int)((unsigned)l.z >> 16) - T.base_p;
    for (int i = threadIdx.x; i < cnt_p; i += blockDim.x) T.idx[cnt_n + i] = __ldg(TL.pnbr_loc + T.base_p + i);
  }
  __syncthreads();
}

template <int DIM, typename XT>
__device__ __forceinline__ NodeDesc desc_from_smem(const TileSmem<DIM, XT>& T, int slot) {
  const int4 a = T.desc[2 * slot], b = T.desc[2 * slot + 1];
  NodeDesc d;
  d.rowbase = ((long long)(unsigned)a.y << 32) | (unsigned)a.x;
  d.prowbase = ((long long)(unsigned)a.w << 32) | (unsigned)a.z;
  d.nbr0 = b.x; d.pnbr0 = b.y;
  d.nb = (unsigned short)(b.z & 0xffff); d.np = (unsigned short)((unsigned)b.z >> 16);
  d.pid = b.w;
  return d;
}

// sum[r] += sum_k row_r[k] * xs[pos(k)] with x staged in shared memory.
// Columns k < nbd are velocity columns (DIM per neighbour node), the rest pressure columns.
template <int DIM, int ROWS, typename VT>
__device__ __forceinline__ void tile_row_dot(const VT* const (&rowp)[ROWS], int ncols, int nbd, const unsigned short* nx,
                                             const unsigned short* px, const double* xs, int pbase, int lane,
                                             double (&sum)[ROWS]) {
  for (int k0 = 0; k0 < ncols; k0 += 32 * SPMV_UNROLL) {
    VT v[ROWS][SPMV_UNROLL];
    double xv[SPMV_UNROLL];
#pragma unroll
    for (int u = 0; u < SPMV_UNROLL; ++u) {
      const int k = k0 + 32 * u + lane;
      const bool ok = k < ncols;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) v[r][u] = ok ? __ldcs(rowp[r] + k) : VT(0);
      xv[u] = 0.0;
      if (ok) xv[u] = (k < nbd) ? xs[(int)nx[k / DIM] * DIM + k % DIM] : xs[pbase + px[k - nbd]];
    }
#pragma unroll
    for (int u = 0; u < SPMV_UNROLL; ++u)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) sum[r] += (double)v[r][u] * xv[u];
  }
}

// y = A x
template <int DIM, typename VT>
__global__ void __launch_bounds__(SPMV_WARPS * 32)
k_spmv_full(DevMesh M, SpmvTiles TL, const VT* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ TileSmem<DIM> T;
  const int lane = threadIdx.x & 31;
  int n0, n1;
  stage_tile<DIM, true>(M, TL, blockIdx.x, x, T, n0, n1);
  const int pbase = T.nuq * DIM;
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&T.next, 1);
    slot = __shfl_sync(NSB_FULL, slot, 0);
    const int A = n0 + slot;
    if (A >= n1) break;
    const NodeDesc d = desc_from_smem(T, slot);
    const int nbd = DIM * d.nb, len = nbd + d.np;
    const unsigned short* nx = T.idx + (d.nbr0 - T.base_n);
    const unsigned short* px = T.idx + T.cnt_n + (d.pnbr0 - T.base_p);
    if (d.pid >= 0) {
      const VT* rowp[DIM + 1];
#pragma unroll
      for (int c = 0; c < DIM; ++c) rowp[c] = vals + d.rowbase + (long long)c * len;
      rowp[DIM] = vals + d.prowbase;
      double sum[DIM + 1];
#pragma unroll
      for (int c = 0; c <= DIM; ++c) sum[c] = 0.0;
      tile_row_dot<DIM, DIM + 1, VT>(rowp, len, nbd, nx, px, T.xs, pbase, lane, sum);
#pragma unroll
      for (int c = 0; c <= DIM; ++c) sum[c] = warp_sum_fixed(sum[c]);
      if (lane <= DIM) {
        double v = sum[0];
#pragma unroll
        for (int c = 1; c <= DIM; ++c) if (lane == c) v = sum[c];
        if (lane < DIM) y[DIM * A + lane] = v;
        else y[DIM * M.nn_own + d.pid] = v;
      }
    } else {
      const VT* rowp[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) rowp[c] = vals + d.rowbase + (long long)c * len;
      double sum[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) sum[c] = 0.0;
      tile_row_dot<DIM, DIM, VT>(rowp, len, nbd, nx, px, T.xs, pbase, lane, sum);
#pragma unroll
      for (int c = 0; c < DIM; ++c) sum[c] = warp_sum_fixed(sum[c]);
      if (lane < DIM) {
        double v = sum[0];
#pragma unroll
        for (int c = 1; c < DIM; ++c) if (lane == c) v = sum[c];
        y[DIM * A + lane] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Velocity block F = A(0,0) only (columns < dim*nb of the velocity rows), with the fused
// epilogues of the polynomial preconditioner used in place of Ifpack ILU(1)
// (reference NavierStokes.hpp:302-304, 325).  Dinv = inverse node-diagonal blocks.
//   MODE 0:  y = F x
//   MODE 2:  y = Dinv (F x)                                     (Arnoldi on the scaled block)
//   MODE 4:  t = Dinv (F x) ; y = cu*u + ct*t
//   MODE 3:  t = Dinv (F x) ; y = cu*u + ct*t ; poly += cpu*u + cpy*y   (one root of the
//            GMRES polynomial in product form; u is read at the node's own entries)
// ------------------------------------------------------------------------------------
struct PolyCoef {
  double cu, ct, cpu, cpy;
};

// Operands of the fused epilogue, fetched at node start so that their latency overlaps the value stream.
template <int DIM> struct EpiOps {
  double dv[DIM];      // row `lane` of the inverse node-diagonal block
  double uv, pv;       // u[row], poly[row]
};

template <int DIM, int MODE>
__device__ __forceinline__ void vel_prefetch(int A, int lane, const double* __restrict__ u, const double* __restrict__ poly,
                                             const double* __restrict__ dinv, EpiOps<DIM>& e) {
  e.uv = 0.0; e.pv = 0.0;
#pragma unroll
  for (int c = 0; c < DIM; ++c) e.dv[c] = 0.0;
  if (MODE != 0 && lane < DIM) {
#pragma unroll
    for (int c = 0; c < DIM; ++c) e.dv[c] = __ldg(dinv + (size_t)A * DIM * DIM + lane * DIM + c);
    if (MODE >= 3) { e.uv = u[DIM * A + lane]; if (MODE == 3) e.pv = poly[DIM * A + lane]; }
  }
}

template <int DIM, int MODE>
__device__ __forceinline__ void vel_epilogue(int A, int lane, const double (&sum)[DIM], double* __restrict__ y,
                                             double* __restrict__ poly, const EpiOps<DIM>& e, const PolyCoef& pc) {
  if (lane < DIM) {
    const int row = DIM * A + lane;
    if (MODE == 0) {
      double v = sum[0];
#pragma unroll
      for (int c = 1; c < DIM; ++c) if (lane == c) v = sum[c];
      y[row] = v;
    } else {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c) t += e.dv[c] * sum[c];
      if (MODE == 2) {
        y[row] = t;
      } else {
        const double yv = pc.cu * e.uv + pc.ct * t;
        y[row] = yv;
        if (MODE == 3) poly[row] = e.pv + pc.cpu * e.uv + pc.cpy * yv;
      }
    }
  }
}

template <int DIM, int MODE, typename VT>
__global__ void __launch_bounds__(SPMV_WARPS * 32)
k_spmv_vel(DevMesh M, SpmvTiles TL, const VT* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
           const double* __restrict__ u, double* __restrict__ poly, const double* __restrict__ dinv, PolyCoef pc) {
  __shared__ TileSmem<DIM> T;
  const int lane = threadIdx.x & 31;
  int n0, n1;
  stage_tile<DIM, false>(M, TL, blockIdx.x, x, T, n0, n1);
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&T.next, 1);
    slot = __shfl_sync(NSB_FULL, slot, 0);
    const int A = n0 + slot;
    if (A >= n1) break;
    const NodeDesc d = desc_from_smem(T, slot);
    const int nbd = DIM * d.nb, len = nbd + d.np;
    EpiOps<DIM> eo;
    vel_prefetch<DIM, MODE>(A, lane, u, poly, dinv, eo);
    const VT* rowp[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) rowp[c] = vals + d.rowbase + (long long)c * len;
    double sum[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) sum[c] = 0.0;
    tile_row_dot<DIM, DIM, VT>(rowp, nbd, nbd, T.idx + (d.nbr0 - T.base_n), nullptr, T.xs, 0, lane, sum);
#pragma unroll
    for (int c = 0; c < DIM; ++c) sum[c] = warp_sum_fixed(sum[c]);
    vel_epilogue<DIM, MODE>(A, lane, sum, y, poly, eo, pc);
  }
}

// t = g - B y0 : pressure rows, velocity columns (reference NavierStokes.hpp:334-335)
template <int DIM, typename VT>
__global__ void __launch_bounds__(SPMV_WARPS * 32)
k_schur_rhs(DevMesh M, const VT* __restrict__ vals, const double* __restrict__ y0, const double* __restrict__ g,
            double* __restrict__ t) {
  const int lane = threadIdx.x & 31;
  const int Pid = blockIdx.x * SPMV_WARPS + (threadIdx.x >> 5);
  if (Pid >= M.np_own) return;
  const NodeDesc d = load_desc(M.nd + M.pid_node[Pid]);
  const int nbd = DIM * d.nb;
  const VT* rowp[1] = {vals + d.prowbase};
  double s[1] = {0.0};
  row_block_dot<DIM, 1, VT>(rowp, nbd, nbd, M.nbr_xoff + d.nbr0, nullptr, y0, lane, s);
  const double r = warp_sum_fixed(s[0]);
  if (lane == 0) t[Pid] = g[DIM * M.nn_own + Pid] - r;
}

// y = Dinv x (node-block Jacobi scaling of a velocity vector)
template <int DIM>
__global__ void k_block_scale(int nn, const double* __restrict__ dinv, const double* __restrict__ x, double* __restrict__ y) {
  const int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= nn) return;
  double r[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) r[c] = x[DIM * A + c];
#pragma unroll
  for (int e = 0; e < DIM; ++e) {
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < DIM; ++c) t += dinv[(size_t)A * DIM * DIM + e * DIM + c] * r[c];
    y[DIM * A + e] = t;
  }
}

// ------------------------------------------------------------------------------------
// generic CSR kernels (pressure mass matrix, pressure-Laplacian multigrid levels; these
// matrices are L2-resident, the kernels are latency-bound)
// ------------------------------------------------------------------------------------
struct DevCsr {
  int n = 0, m = 0;
  const int* ptr = nullptr;
  const int* col = nullptr;
  const double* val = nullptr;
};

// MODE 0: y = A x ; 1: y = b - A x ; 2: y += A x ;
// MODE 3: r = b - A x ; d = c1 d + c2 dinv r ; y = x + d     (scalar-Jacobi Chebyshev step)
template <int MODE>
__global__ void k_csr(DevCsr A, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
                      double* __restrict__ d, const double* __restrict__ dinv, double c1, double c2) {
  constexpr int LPR = 8;                            // lanes per row
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int sub = threadIdx.x % LPR;
  double s = 0.0;
  if (row < A.n)
    for (int k = A.ptr[row] + sub; k < A.ptr[row + 1]; k += LPR) s += A.val[k] * __ldg(x + A.col[k]);
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(NSB_FULL, s, o);
  if (row < A.n && sub == 0) {
    if (MODE == 0) y[row] = s;
    if (MODE == 1) y[row] = b[row] - s;
    if (MODE == 2) y[row] += s;
    if (MODE == 3) {
      const double dn = c1 * d[row] + c2 * dinv[row] * (b[row] - s);
      d[row] = dn;
      y[row] = x[row] + dn;
    }
  }
}

__global__ void k_cheb_first(int n, const double* __restrict__ dinv, const double* __restrict__ b,
                             double* __restrict__ d, double* __restrict__ x, double inv_theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double t = dinv[i] * b[i] * inv_theta; d[i] = t; x[i] = t; }
}

// y = Ainv b, dense n x n (coarsest multigrid level); one warp per row
__global__ void k_dense_mv(int n, const double* __restrict__ Ainv, const double* __restrict__ b, double* __restrict__ y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) s += Ainv[(size_t)row * n + k] * b[k];
  s = warp_sum_fixed(s);
  if (lane == 0) y[row] = s;
}

// ------------------------------------------------------------------------------------
// BLAS-1
// ------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_ELEMS = 8;                         // elements per thread per chunk
constexpr int RED_CHUNK = RED_THREADS * RED_ELEMS;   // rows per block

__device__ __forceinline__ double block_sum_fixed(double v, double* sbuf) {
  v = warp_sum_fixed(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sbuf[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (wid == 0) {
    t = (lane < (int)(blockDim.x >> 5)) ? sbuf[lane] : 0.0;
    t = warp_sum_fixed(t);
  }
  __syncthreads();
  return t;     // valid on warp 0
}

// partial[j*nblk + blk] = sum over the block's chunk of V_j[i] * w[i],  j = 0..nv-1.
// V_j = V + j*ld.  One pass over w, one pass over each V_j.
__global__ void __launch_bounds__(RED_THREADS)
k_multi_dot(int nv, const double* __restrict__ V, long long ld, const double* __restrict__ w, long long n,
            double* __restrict__ partial) {
  __shared__ double sbuf[RED_THREADS / 32];
  const long long base = (long long)blockIdx.x * RED_CHUNK;
  double wv[RED_ELEMS];
#pragma unroll
  for (int e = 0; e < RED_ELEMS; ++e) {
    const long long i = base + e * RED_THREADS + threadIdx.x;
    wv[e] = (i < n) ? w[i] : 0.0;
  }
  for (int j = 0; j < nv; ++j) {
    const double* vj = V + (long long)j * ld;
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < RED_ELEMS; ++e) {
      const long long i = base + e * RED_THREADS + threadIdx.x;
      if (i < n) s += __ldcs(vj + i) * wv[e];
    }
    s = block_sum_fixed(s, sbuf);
    if (threadIdx.x == 0) partial[(long long)j * gridDim.x + blockIdx.x] = s;
  }
}

// out[j] (+)= sum_blk partial[j*nblk + blk]   (one block per j, fixed order)
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_partials(int nblk, const double* __restrict__ partial, double* __restrict__ out, int accumulate) {
  __shared__ double sbuf[RED_THREADS / 32];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < nblk; k += RED_THREADS) s += partial[(long long)j * nblk + k];
  s = block_sum_fixed(s, sbuf);
  if (threadIdx.x == 0) out[j] = accumulate ? out[j] + s : s;
}

// w -= sum_j h[j] V_j  (sign = -1)   or   x += sum_j h[j] V_j  (sign = +1);
// optionally also writes the per-block partial of ||w||^2 afterwards.
__global__ void __launch_bounds__(RED_THREADS)
k_multi_axpy(int nv, const double* __restrict__ V, long long ld, const double* __restrict__ h, double sign,
             double* __restrict__ w, long long n, double* __restrict__ nrm_partial) {
  __shared__ double sbuf[RED_THREADS / 32];
  __shared__ double sh[160];
  for (int j = threadIdx.x; j < nv; j += RED_THREADS) sh[j] = sign * h[j];
  __syncthreads();
  const long long base = (long long)blockIdx.x * RED_CHUNK;
  double acc[RED_ELEMS];
#pragma unroll
  for (int e = 0; e < RED_ELEMS; ++e) {
    const long long i = base + e * RED_THREADS + threadIdx.x;
    acc[e] = (i < n) ? w[i] : 0.0;
  }
  for (int j = 0; j < nv; ++j) {
    const double* vj = V + (long long)j * ld;
    const double hj = sh[j];
#pragma unroll
    for (int e = 0; e < RED_ELEMS; ++e) {
      const long long i = base + e * RED_THREADS + threadIdx.x;
      if (i < n) acc[e] += hj * __ldg(vj + i);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < RED_ELEMS; ++e) {
    const long long i = base + e * RED_THREADS + threadIdx.x;
    if (i < n) { w[i] = acc[e]; s += acc[e] * acc[e]; }
  }
  if (nrm_partial) {
    s = block_sum_fixed(s, sbuf);
    if (threadIdx.x == 0) nrm_partial[blockIdx.x] = s;
  }
}

// partial ||x||^2
__global__ void __launch_bounds__(RED_THREADS)
k_norm2_partial(const double* __restrict__ x, long long n, double* __restrict__ partial) {
  __shared__ double sbuf[RED_THREADS / 32];
  const long long base = (long long)blockIdx.x * RED_CHUNK;
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < RED_ELEMS; ++e) {
    const long long i = base + e * RED_THREADS + threadIdx.x;
    if (i < n) { const double v = x[i]; s += v * v; }
  }
  s = block_sum_fixed(s, sbuf);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// y = alpha[0-dim device scalar inverse-sqrt] * x : v_{k+1} = w / ||w||, norm2 read from device memory
__global__ void k_scale_by_inv_norm(long long n, const double* __restrict__ x, const double* __restrict__ nrm2,
                                    double* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double s = 1.0 / sqrt(*nrm2);
  if (i < n) y[i] = x[i] * s;
}

__global__ void k_axpby(long long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * y[i];
}

// y = a*x + b*z
__global__ void k_lincomb(long long n, double a, const double* __restrict__ x, double b, const double* __restrict__ z,
                          double* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * z[i];
}

// out[i] = a * x[idx[i]] + b * z[idx[i]]   (owned pressure entries out of replicated global vectors)
__global__ void k_lincomb_gather(int n, const int* __restrict__ idx, double a, const double* __restrict__ x, double b,
                                 const double* __restrict__ z, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int g = idx[i]; out[i] = a * x[g] + b * z[g]; }
}

// scatter constraint values / flags;  x[dof] = val  (constraints.distribute, reference cpp:566, 862)
__global__ void k_scatter_vals(int n, const int* __restrict__ idx, const double* __restrict__ val, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[idx[i]] = val[i];
}
__global__ void k_scatter_flags(int n, const int* __restrict__ idx, unsigned char* __restrict__ f, unsigned char v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f[idx[i]] = v;
}
__global__ void k_gather(int n, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[idx[i]];
}
// packs dim consecutive values per node index (halo send buffers)
__global__ void k_gather_nodes(int n, int dim, const int* __restrict__ xoff, const double* __restrict__ x, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * dim) out[i] = x[xoff[i / dim] + i % dim];
}

}  // namespace nsb

// ------------------------------------------------------------------------------------
// Halo exchange by direct peer stores (NVLink): the sender writes its boundary values into the ghost tails of the peers'
// copies of the same vector and then raises a per-sender sequence flag in each peer; the receiver's next kernel in
// stream order is k_halo_wait.  Replaces the Epetra Import of ghosted vectors (reference NavierStokes.cpp:561, 853,
// 1053-1056, 1299-1300).
// ------------------------------------------------------------------------------------
namespace nsb {

// nseg segments; segment s holds entries [seg_ptr[s], seg_ptr[s+1]) of src (offsets into v) and lands at
// peer_vec[s % npeers] + land[s] in the peer's memory.
__global__ void __launch_bounds__(256)
k_halo_push(int nseg, int npeers, const int* __restrict__ seg_ptr, const int* __restrict__ src, const long long* __restrict__ land,
            double* const* __restrict__ peer_arena, long long vec_off, const double* __restrict__ v,
            unsigned long long* const* __restrict__ peer_flags, int my_rank, unsigned long long seq, unsigned int* counter) {
  const int total = seg_ptr[nseg];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int s = 0;
    while (e >= seg_ptr[s + 1]) ++s;
    double* dst = peer_arena[s % npeers] + vec_off + land[s] + (e - seg_ptr[s]);
    *dst = v[src[e]];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int prev;
  if (threadIdx.x == 0) prev = atomicAdd(counter, 1u);
  __syncthreads();
  if (prev == gridDim.x - 1) {
    // every block's stores are fenced: publish the sequence number to the peers
    __threadfence_system();
    if (threadIdx.x < npeers) {
      volatile unsigned long long* f = peer_flags[threadIdx.x] + my_rank;
      *f = seq;
    }
    if (threadIdx.x == 0) *counter = 0;
    __threadfence_system();
  }
}

// waits until every peer has written sequence number >= seq into flags[peer rank]
__global__ void k_halo_wait(int npeers, const int* __restrict__ peer_rank, const unsigned long long* flags, unsigned long long seq) {
  if (threadIdx.x < npeers) {
    const volatile unsigned long long* f = flags + peer_rank[threadIdx.x];
    while (*f < seq) { }
  }
  __threadfence_system();
}

// tells the peers that this rank has consumed every exchange up to `seq` (its kernels that read those ghosts precede this
// kernel in stream order): flags[nranks + my_rank] = seq in each peer
__global__ void k_halo_ack(int npeers, unsigned long long* const* __restrict__ peer_flags, int nranks, int my_rank, unsigned long long seq) {
  if (threadIdx.x < npeers) {
    volatile unsigned long long* f = peer_flags[threadIdx.x] + nranks + my_rank;
    *f = seq;
  }
  __threadfence_system();
}

}  // namespace nsb
