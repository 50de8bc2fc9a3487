// See structure.hpp.  Pure host code (C++17 + OpenMP), runs once per mesh.
#include "structure.hpp"

#include <algorithm>
#include <cstring>
#include <climits>
#include <cmath>
#include <numeric>

namespace nsb {

namespace {
struct Layout {
  int dim, NV, NN, DPC;
  // local dof index of (node a, component c); pressure of vertex v
  int udof(int a, int c) const { return a < NV ? a * (dim + 1) + c : NV * (dim + 1) + (a - NV) * dim + c; }
  int pdof(int v) const { return v * (dim + 1) + dim; }
};
}  // namespace

std::string build_structure(int dim, int64_t n_vertices, const double* coords, int64_t n_cells,
                            const uint32_t* cell_vertices, const uint32_t* cell_dofs, int64_t n_u,
                            int64_t n_p, const int32_t* cell_part, int rank, int nranks, Structure& S) {
  if (dim != 2 && dim != 3) return "dim must be 2 or 3";
  Layout L{dim, dim + 1, (dim == 2) ? 6 : 10, 0};
  L.DPC = dim * L.NN + L.NV;
  const int NV = L.NV, NN = L.NN, DPC = L.DPC;
  if (n_u % dim != 0) return "n_u is not a multiple of dim";
  const int64_t Nn = n_u / dim;
  if (Nn >= INT_MAX || n_p >= INT_MAX || n_cells >= (int64_t(1) << 28))
    return "mesh too large for 32-bit local indices";
  S = Structure();
  S.dim = dim; S.NV = NV; S.NN = NN; S.DPC = DPC;
  S.rank = rank; S.nranks = nranks;
  S.n_u_glob = n_u; S.n_p_glob = n_p; S.n_cells_glob = n_cells;

  // ---- 1. node / pressure ids per cell, validated against the node-block assumption
  std::vector<int> cn((size_t)n_cells * NN), cp((size_t)n_cells * NV);
  std::vector<int> node_pid_glob(Nn, -1);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t c = 0; c < n_cells; ++c) {
    const uint32_t* d = cell_dofs + (size_t)c * DPC;
    for (int a = 0; a < NN; ++a) {
      const uint32_t d0 = d[L.udof(a, 0)];
      if (d0 >= n_u || d0 % dim != 0) { ++bad; continue; }
      for (int k = 1; k < dim; ++k)
        if (d[L.udof(a, k)] != d0 + k) ++bad;
      cn[(size_t)c * NN + a] = (int)(d0 / dim);
    }
    for (int v = 0; v < NV; ++v) {
      const uint32_t pd = d[L.pdof(v)];
      if (pd < n_u || pd >= n_u + n_p) { ++bad; continue; }
      cp[(size_t)c * NV + v] = (int)(pd - n_u);
    }
  }
  if (bad) return "cell_dofs violate the node-block numbering (velocity DoFs of a node must be consecutive, pressure after velocity)";
  for (int64_t c = 0; c < n_cells; ++c)
    for (int v = 0; v < NV; ++v) {
      int& slot = node_pid_glob[cn[(size_t)c * NN + v]];
      const int pid = cp[(size_t)c * NV + v];
      if (slot == -1) slot = pid;
      else if (slot != pid) return "inconsistent vertex -> pressure DoF map";
    }

  // ---- 2. ownership: a node belongs to the lowest part touching it
  std::vector<int> owner(Nn, INT_MAX);
  for (int64_t c = 0; c < n_cells; ++c) {
    const int p = cell_part ? cell_part[c] : 0;
    if (p < 0 || p >= nranks) return "cell_part out of range";
    for (int a = 0; a < NN; ++a) {
      int& o = owner[cn[(size_t)c * NN + a]];
      o = std::min(o, p);
    }
  }
  for (int64_t n = 0; n < Nn; ++n)
    if (owner[n] == INT_MAX) return "velocity node without a cell";

  // ---- 3./4. local cells, owned and ghost nodes
  std::vector<int> g2l(Nn, -1);
  int nn_own = 0;
  for (int64_t n = 0; n < Nn; ++n)
    if (owner[n] == rank) g2l[n] = nn_own++;
  std::vector<int>& cell_gid = S.cell_gid;
  std::vector<int> ghosts;
  for (int64_t c = 0; c < n_cells; ++c) {
    bool touch = false;
    for (int a = 0; a < NN; ++a) touch |= (owner[cn[(size_t)c * NN + a]] == rank);
    if (!touch) continue;
    cell_gid.push_back((int)c);
    for (int a = 0; a < NN; ++a) {
      const int n = cn[(size_t)c * NN + a];
      if (owner[n] != rank && g2l[n] == -1) { g2l[n] = -2; ghosts.push_back(n); }
    }
  }
  std::sort(ghosts.begin(), ghosts.end(), [&](int x, int y) {
    return owner[x] != owner[y] ? owner[x] < owner[y] : x < y;
  });
  for (size_t k = 0; k < ghosts.size(); ++k) g2l[ghosts[k]] = nn_own + (int)k;
  const int nn_ghost = (int)ghosts.size(), nn_tot = nn_own + nn_ghost;
  const int nc = (int)cell_gid.size();
  S.nn_own = nn_own; S.nn_ghost = nn_ghost; S.nc = nc;
  S.node_gid.resize(nn_tot);
  for (int64_t n = 0; n < Nn; ++n)
    if (g2l[n] >= 0) S.node_gid[g2l[n]] = n;
  // pressure ids follow their vertex node
  std::vector<int> pg2l(n_p, -1);
  S.node_pid.assign(nn_tot, -1);
  int np_own = 0, np_tot = 0;
  for (int A = 0; A < nn_tot; ++A) {
    if (A == nn_own) np_own = np_tot;
    const int pid = node_pid_glob[S.node_gid[A]];
    if (pid >= 0) { pg2l[pid] = np_tot; S.node_pid[A] = np_tot++; }
  }
  if (nn_ghost == 0) np_own = np_tot;
  S.np_own = np_own; S.np_ghost = np_tot - np_own;
  S.pid_gid.resize(np_tot); S.pid_node.resize(np_tot);
  for (int A = 0; A < nn_tot; ++A)
    if (S.node_pid[A] >= 0) {
      S.pid_gid[S.node_pid[A]] = node_pid_glob[S.node_gid[A]];
      S.pid_node[S.node_pid[A]] = A;
    }

  // ---- 5. local connectivity
  S.cell_nodes.resize((size_t)nc * NN);
  S.cell_pids.resize((size_t)nc * NV);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    const int64_t g = cell_gid[c];
    for (int a = 0; a < NN; ++a) S.cell_nodes[(size_t)c * NN + a] = g2l[cn[(size_t)g * NN + a]];
    for (int v = 0; v < NV; ++v) S.cell_pids[(size_t)c * NV + v] = pg2l[cp[(size_t)g * NV + v]];
  }

  // ---- 6. node -> cells (ascending local cell id: the deterministic accumulation order)
  S.n2c_ptr.assign(nn_own + 1, 0);
  for (int c = 0; c < nc; ++c)
    for (int a = 0; a < NN; ++a) {
      const int A = S.cell_nodes[(size_t)c * NN + a];
      if (A < nn_own) S.n2c_ptr[A + 1]++;
    }
  for (int A = 0; A < nn_own; ++A) S.n2c_ptr[A + 1] += S.n2c_ptr[A];
  S.n2c.resize(S.n2c_ptr[nn_own]);
  {
    std::vector<int64_t> fill(S.n2c_ptr.begin(), S.n2c_ptr.end() - 1);
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < NN; ++a) {
        const int A = S.cell_nodes[(size_t)c * NN + a];
        if (A < nn_own) S.n2c[fill[A]++] = ((uint32_t)c << 4) | (uint32_t)a;
      }
  }

  // ---- 7. neighbour lists sorted by global id
  std::vector<int64_t> cnt_n(nn_own + 1, 0), cnt_p(nn_own + 1, 0);
  std::vector<std::vector<int>> tmp_n(nn_own), tmp_p(nn_own);
#pragma omp parallel
  {
    std::vector<std::pair<int64_t, int>> buf;
#pragma omp for schedule(dynamic, 1024)
    for (int A = 0; A < nn_own; ++A) {
      buf.clear();
      for (int64_t k = S.n2c_ptr[A]; k < S.n2c_ptr[A + 1]; ++k) {
        const int c = (int)(S.n2c[k] >> 4);
        for (int b = 0; b < NN; ++b) {
          const int B = S.cell_nodes[(size_t)c * NN + b];
          buf.emplace_back(S.node_gid[B], B);
        }
      }
      std::sort(buf.begin(), buf.end());
      buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
      tmp_n[A].resize(buf.size());
      for (size_t k = 0; k < buf.size(); ++k) tmp_n[A][k] = buf[k].second;
      buf.clear();
      for (int64_t k = S.n2c_ptr[A]; k < S.n2c_ptr[A + 1]; ++k) {
        const int c = (int)(S.n2c[k] >> 4);
        for (int v = 0; v < NV; ++v) {
          const int P = S.cell_pids[(size_t)c * NV + v];
          buf.emplace_back(S.pid_gid[P], P);
        }
      }
      std::sort(buf.begin(), buf.end());
      buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
      tmp_p[A].resize(buf.size());
      for (size_t k = 0; k < buf.size(); ++k) tmp_p[A][k] = buf[k].second;
      cnt_n[A + 1] = (int64_t)tmp_n[A].size();
      cnt_p[A + 1] = (int64_t)tmp_p[A].size();
    }
  }
  for (int A = 0; A < nn_own; ++A) { cnt_n[A + 1] += cnt_n[A]; cnt_p[A + 1] += cnt_p[A]; }
  S.nbr_ptr = cnt_n; S.pnbr_ptr = cnt_p;
  S.nbr.resize(cnt_n[nn_own]); S.pnbr.resize(cnt_p[nn_own]);
  S.selfrank.assign(nn_own, -1);
#pragma omp parallel for schedule(static)
  for (int A = 0; A < nn_own; ++A) {
    std::copy(tmp_n[A].begin(), tmp_n[A].end(), S.nbr.begin() + cnt_n[A]);
    std::copy(tmp_p[A].begin(), tmp_p[A].end(), S.pnbr.begin() + cnt_p[A]);
    for (size_t k = 0; k < tmp_n[A].size(); ++k)
      if (tmp_n[A][k] == A) S.selfrank[A] = (int)k;
    std::vector<int>().swap(tmp_n[A]);
    std::vector<int>().swap(tmp_p[A]);
  }
  S.pselfrank.assign(np_own, -1);
  for (int P = 0; P < np_own; ++P) {
    const int A = S.pid_node[P];
    for (int64_t k = S.pnbr_ptr[A]; k < S.pnbr_ptr[A + 1]; ++k)
      if (S.pnbr[k] == P) S.pselfrank[P] = (int)(k - S.pnbr_ptr[A]);
  }

  // ---- 8. row offsets in the scalar-CSR value array
  S.rowbase.resize(nn_own); S.prowbase.resize(np_own);
  int64_t off = 0;
  int maxlen = 0, maxsm = 0;
  for (int A = 0; A < nn_own; ++A) {
    const int len = S.row_len(A);
    if (len >= 65536) return "row too long for 16-bit neighbour ranks";
    S.rowbase[A] = off;
    off += (int64_t)dim * len;
    maxlen = std::max(maxlen, len);
    maxsm = std::max(maxsm, (dim + (S.node_pid[A] >= 0 ? 1 : 0)) * len);
  }
  for (int P = 0; P < np_own; ++P) {
    S.prowbase[P] = off;
    off += S.row_len(S.pid_node[P]);
  }
  S.nnz_local = off;
  S.max_row_len = maxlen;
  S.max_node_smem_doubles = maxsm;

  // ---- 9. per-cell rank tables (rows of ghost nodes are never assembled here -> 0)
  S.rank_uu.assign((size_t)nc * NN * NN, 0);
  S.rank_up.assign((size_t)nc * NN * NV, 0);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int c = 0; c < nc; ++c) {
    for (int a = 0; a < NN; ++a) {
      const int A = S.cell_nodes[(size_t)c * NN + a];
      if (A >= nn_own) continue;
      const int* nb = S.nbr.data() + S.nbr_ptr[A];
      const int nnb = (int)(S.nbr_ptr[A + 1] - S.nbr_ptr[A]);
      for (int b = 0; b < NN; ++b) {
        const int64_t g = S.node_gid[S.cell_nodes[(size_t)c * NN + b]];
        int lo = 0, hi = nnb;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (S.node_gid[nb[mid]] < g) lo = mid + 1; else hi = mid;
        }
        S.rank_uu[((size_t)c * NN + a) * NN + b] = (uint16_t)lo;
      }
      const int* pb = S.pnbr.data() + S.pnbr_ptr[A];
      const int npb = (int)(S.pnbr_ptr[A + 1] - S.pnbr_ptr[A]);
      for (int v = 0; v < NV; ++v) {
        const int64_t g = S.pid_gid[S.cell_pids[(size_t)c * NV + v]];
        int lo = 0, hi = npb;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (S.pid_gid[pb[mid]] < g) lo = mid + 1; else hi = mid;
        }
        S.rank_up[((size_t)c * NN + a) * NV + v] = (uint16_t)lo;
      }
    }
  }

  // ---- 10. geometry of the affine map (MappingFE(FE_SimplexP(1)))
  S.cell_geom.assign((size_t)nc * 16, 0.0);
  int negative = 0;
#pragma omp parallel for schedule(static) reduction(+ : negative)
  for (int c = 0; c < nc; ++c) {
    const uint32_t* cv = cell_vertices + (size_t)cell_gid[c] * NV;
    double X[4][3] = {{0}};
    for (int v = 0; v < NV; ++v)
      for (int k = 0; k < dim; ++k) X[v][k] = coords[(size_t)cv[v] * dim + k];
    double* g = &S.cell_geom[(size_t)c * 16];
    double det;
    double gl[4][3] = {{0}};
    if (dim == 2) {
      const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0];     // J = [[a b],[c d]] columns x_k - x_0
      const double cc = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
      det = a * d - b * cc;
      const double id = 1.0 / det;
      gl[1][0] = d * id;  gl[1][1] = -b * id;          // rows of J^{-1}
      gl[2][0] = -cc * id; gl[2][1] = a * id;
      for (int k = 0; k < 2; ++k) gl[0][k] = -(gl[1][k] + gl[2][k]);
    } else {
      double J[3][3];
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) J[r][k] = X[k + 1][r] - X[0][r];
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double id = 1.0 / det;
      // J^{-1} = adj(J)/det ; row k of J^{-1} is grad lambda_{k+1}
      gl[1][0] = c00 * id;
      gl[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
      gl[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      gl[2][0] = c01 * id;
      gl[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
      gl[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      gl[3][0] = c02 * id;
      gl[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
      gl[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      for (int k = 0; k < 3; ++k) gl[0][k] = -((gl[1][k] + gl[2][k]) + gl[3][k]);
    }
    if (!(det > 0)) ++negative;
    double h = 0;
    for (int v = 0; v < NV; ++v)
      for (int w = v + 1; w < NV; ++w) {
        double s = 0;
        for (int k = 0; k < dim; ++k) s += (X[v][k] - X[w][k]) * (X[v][k] - X[w][k]);
        h = std::max(h, std::sqrt(s));
      }
    for (int v = 0; v < NV; ++v)
      for (int k = 0; k < dim; ++k) g[v * dim + k] = gl[v][k];
    g[12] = det;
    g[13] = h;
    g[14] = 1.0 / h;
  }
  if (negative) return "mesh has cells with non-positive measure";
  (void)n_vertices;

  // ---- 11. halo plan.  Every rank derives all send lists from the replicated mesh:
  // for a cell whose nodes have owners O, each node n is a ghost on every s in O \ {owner(n)}.
  if (nranks > 1) {
    std::vector<std::vector<int64_t>> send_to(nranks);   // global node ids this rank sends to s
    std::vector<char> seen(nranks);
    for (int64_t c = 0; c < n_cells; ++c) {
      int os[10], no = 0;
      for (int a = 0; a < NN; ++a) {
        const int o = owner[cn[(size_t)c * NN + a]];
        bool dup = false;
        for (int k = 0; k < no; ++k) dup |= (os[k] == o);
        if (!dup) os[no++] = o;
      }
      if (no == 1) continue;
      for (int a = 0; a < NN; ++a) {
        const int n = cn[(size_t)c * NN + a];
        if (owner[n] != rank) continue;
        for (int k = 0; k < no; ++k)
          if (os[k] != rank) send_to[os[k]].push_back(n);
      }
    }
    std::vector<int64_t> recv_cnt(nranks, 0), recv_pcnt(nranks, 0);
    for (int k = 0; k < nn_ghost; ++k) {
      recv_cnt[owner[ghosts[k]]]++;
      if (node_pid_glob[ghosts[k]] >= 0) recv_pcnt[owner[ghosts[k]]]++;
    }
    for (int s = 0; s < nranks; ++s) {
      auto& v = send_to[s];
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      if (v.empty() && recv_cnt[s] == 0) continue;
      S.peer.push_back(s);
      std::vector<int> sn, sp;
      for (int64_t n : v) {
        sn.push_back(g2l[n]);
        if (node_pid_glob[n] >= 0) sp.push_back(pg2l[node_pid_glob[n]]);
      }
      S.send_nodes.push_back(sn);
      S.send_pids.push_back(sp);
      S.recv_node_count.push_back((int)recv_cnt[s]);
      S.recv_pid_count.push_back((int)recv_pcnt[s]);
    }
  }
  return "";
}

void export_pattern(const Structure& S, std::vector<int64_t>& rowptr, std::vector<uint32_t>& col) {
  const int dim = S.dim;
  const int64_t nrows = S.n_own_dofs();
  rowptr.assign(nrows + 1, 0);
  col.resize(S.nnz_local);
  auto fill_row = [&](int A, uint32_t* out) {
    int64_t k = 0;
    for (int64_t j = S.nbr_ptr[A]; j < S.nbr_ptr[A + 1]; ++j)
      for (int d = 0; d < dim; ++d) out[k++] = (uint32_t)(S.node_gid[S.nbr[j]] * dim + d);
    for (int64_t j = S.pnbr_ptr[A]; j < S.pnbr_ptr[A + 1]; ++j)
      out[k++] = (uint32_t)(S.n_u_glob + S.pid_gid[S.pnbr[j]]);
  };
  for (int A = 0; A < S.nn_own; ++A) {
    const int len = S.row_len(A);
    for (int c = 0; c < dim; ++c) {
      rowptr[(int64_t)dim * A + c] = S.rowbase[A] + (int64_t)c * len;
      fill_row(A, col.data() + S.rowbase[A] + (int64_t)c * len);
    }
  }
  for (int P = 0; P < S.np_own; ++P) {
    rowptr[(int64_t)dim * S.nn_own + P] = S.prowbase[P];
    fill_row(S.pid_node[P], col.data() + S.prowbase[P]);
  }
  rowptr[nrows] = S.nnz_local;
}

void export_row_gids(const Structure& S, std::vector<int64_t>& gid) {
  gid.resize(S.n_own_dofs());
  for (int A = 0; A < S.nn_own; ++A)
    for (int c = 0; c < S.dim; ++c) gid[(int64_t)S.dim * A + c] = S.node_gid[A] * S.dim + c;
  for (int P = 0; P < S.np_own; ++P) gid[(int64_t)S.dim * S.nn_own + P] = S.n_u_glob + S.pid_gid[P];
}

std::string build_tile_plan(const Structure& S, const TileLimits& L, TilePlan& P) {
  P = TilePlan();
  const int ntot = S.nn_own + S.nn_ghost, ptot = S.np_own + S.np_ghost;
  P.nbr_loc.assign(S.nbr.size(), 0);
  P.pnbr_loc.assign(S.pnbr.size(), 0);
  std::vector<int> stamp(ntot, -1), pstamp(ptot, -1), posn(ntot, 0), posp(ptot, 0);
  P.node_ptr.push_back(0); P.uniq_ptr.push_back(0); P.puniq_ptr.push_back(0);
  int A = 0, tile = 0;
  std::vector<int> U, PU;
  while (A < S.nn_own) {
    U.clear(); PU.clear();
    int idx = 0, cn = 0, vblk = 0;
    const int start = A;
    while (A < S.nn_own && cn < L.max_nodes) {
      const int nb = (int)(S.nbr_ptr[A + 1] - S.nbr_ptr[A]), np = (int)(S.pnbr_ptr[A + 1] - S.pnbr_ptr[A]);
      if (nb + np > L.max_idx || nb > L.max_uniq || np > L.max_puniq || (L.max_vel_blocks && (nb + 3) / 4 * 4 > L.max_vel_blocks))
        return "a node has more neighbours than one SpMV tile can stage";
      if (idx + nb + np > L.max_idx) break;
      if (L.max_vel_blocks && vblk + (nb + 3) / 4 * 4 > L.max_vel_blocks) break;      // blocks padded to quads
      vblk += (nb + 3) / 4 * 4;
      int newu = 0, newp = 0;
      for (int64_t k = S.nbr_ptr[A]; k < S.nbr_ptr[A + 1]; ++k) newu += stamp[S.nbr[k]] != tile;
      for (int64_t k = S.pnbr_ptr[A]; k < S.pnbr_ptr[A + 1]; ++k) newp += pstamp[S.pnbr[k]] != tile;
      if ((int)U.size() + newu > L.max_uniq || (int)PU.size() + newp > L.max_puniq) break;
      for (int64_t k = S.nbr_ptr[A]; k < S.nbr_ptr[A + 1]; ++k)
        if (stamp[S.nbr[k]] != tile) { stamp[S.nbr[k]] = tile; U.push_back(S.nbr[k]); }
      for (int64_t k = S.pnbr_ptr[A]; k < S.pnbr_ptr[A + 1]; ++k)
        if (pstamp[S.pnbr[k]] != tile) { pstamp[S.pnbr[k]] = tile; PU.push_back(S.pnbr[k]); }
      idx += nb + np; ++cn; ++A;
    }
    // memory order, then positions
    std::sort(U.begin(), U.end(), [&](int a, int b) { return S.node_xoff(a) < S.node_xoff(b); });
    std::sort(PU.begin(), PU.end(), [&](int a, int b) { return S.pid_xoff(a) < S.pid_xoff(b); });
    for (size_t i = 0; i < U.size(); ++i) { posn[U[i]] = (int)i; P.uniq_xoff.push_back((int)S.node_xoff(U[i])); }
    for (size_t i = 0; i < PU.size(); ++i) { posp[PU[i]] = (int)i; P.puniq_xoff.push_back((int)S.pid_xoff(PU[i])); }
    for (int B = start; B < A; ++B) {
      for (int64_t k = S.nbr_ptr[B]; k < S.nbr_ptr[B + 1]; ++k) P.nbr_loc[k] = (unsigned short)posn[S.nbr[k]];
      for (int64_t k = S.pnbr_ptr[B]; k < S.pnbr_ptr[B + 1]; ++k) P.pnbr_loc[k] = (unsigned short)posp[S.pnbr[k]];
    }
    bool reads_ghost = false;
    for (int n : U) reads_ghost |= (n >= S.nn_own);
    (reads_ghost ? P.tiles_bnd : P.tiles_int).push_back(tile);
    P.node_ptr.push_back(A); P.uniq_ptr.push_back((int)P.uniq_xoff.size()); P.puniq_ptr.push_back((int)P.puniq_xoff.size());
    ++tile;
  }
  return "";
}

std::string build_vel_stream(const Structure& S, const TilePlan& P, VsPlan& V) {
  V = VsPlan();
  const int nt = P.n_tiles();
  V.tiles.resize(nt);
  auto quads = [&](int A) { return (int)((S.nbr_ptr[A + 1] - S.nbr_ptr[A] + 3) / 4); };
  int64_t off = 0;
  std::vector<int> qstart;
  for (int t = 0; t < nt; ++t) {
    VsTile& T = V.tiles[t];
    std::memset(&T, 0, sizeof(T));
    T.n0 = P.node_ptr[t]; T.nn = P.node_ptr[t + 1] - T.n0;
    qstart.assign(T.nn + 1, 0);
    for (int a = 0; a < T.nn; ++a) qstart[a + 1] = qstart[a] + quads(T.n0 + a);
    if (qstart[T.nn] > VS_MAX_QUADS) return "an SpMV tile holds more velocity blocks than the streamed operator can stage";
    T.nquad = qstart[T.nn];
    T.NQ = (T.nquad + 31) / 32 * 32;
    if ((off + T.NQ) * 4 >= (int64_t)INT32_MAX) return "too many velocity blocks for 32-bit tile offsets";
    T.nq_off = (int)off;
    T.u0 = P.uniq_ptr[t]; T.nuq = P.uniq_ptr[t + 1] - T.u0;
    // node-aligned split of the quads among the consumer warps, balanced by quad count
    int a = 0;
    for (int w = 0; w <= VS_CONSUMERS; ++w) {
      const int target = (int)((int64_t)T.nquad * w / VS_CONSUMERS);
      while (a < T.nn && qstart[a] < target) ++a;
      T.split[w] = (unsigned short)qstart[a];
    }
    T.split[VS_CONSUMERS] = (unsigned short)T.nquad;
    off += T.NQ;
  }
  V.total_nq = off;
  V.meta.assign((size_t)off * 4, 0u);
#pragma omp parallel for schedule(static)
  for (int t = 0; t < nt; ++t) {
    const VsTile& T = V.tiles[t];
    uint32_t* m = V.meta.data() + (size_t)T.nq_off * 4;
    int q = 0;
    for (int a = 0; a < T.nn; ++a) {
      const int64_t k0 = S.nbr_ptr[T.n0 + a], k1 = S.nbr_ptr[T.n0 + a + 1];
      for (int64_t k = k0; k < k1; k += 4, ++q) {
        uint32_t loc[4];
        const int cnt = (int)std::min<int64_t>(4, k1 - k);
        for (int e = 0; e < 4; ++e) loc[e] = P.nbr_loc[k + (e < cnt ? e : 0)];      // padding blocks: a valid position, zero values
        m[4 * q + 0] = loc[0] | (loc[1] << 16);
        m[4 * q + 1] = loc[2] | (loc[3] << 16);
        m[4 * q + 2] = (uint32_t)a;
        m[4 * q + 3] = (uint32_t)(k - k0) | ((uint32_t)cnt << 16);
      }
    }
    for (; q < T.NQ; ++q) { m[4 * q + 0] = 0; m[4 * q + 1] = 0; m[4 * q + 2] = 0xffffu; m[4 * q + 3] = 0; }
  }
  return "";
}

int64_t verify_vel_stream(const Structure& S, const TilePlan& P, const VsPlan& V) {
  int64_t bad = 0;
  const int nt = P.n_tiles();
  if ((int)V.tiles.size() != nt) return 1;
  int64_t off = 0;
  int next_node = 0;
  for (int t = 0; t < nt; ++t) {
    const VsTile& T = V.tiles[t];
    if (T.n0 != next_node || T.nn <= 0 || T.nn > 64) ++bad;
    next_node = T.n0 + T.nn;
    if (T.nq_off != off || T.NQ % 32 || T.NQ < T.nquad || T.NQ > VS_MAX_QUADS) ++bad;
    if (T.u0 != P.uniq_ptr[t] || T.nuq != P.uniq_ptr[t + 1] - P.uniq_ptr[t]) ++bad;
    const uint32_t* m = V.meta.data() + (size_t)off * 4;
    int q = 0;
    std::vector<int> qstart(1, 0);
    for (int a = 0; a < T.nn; ++a) {
      const int64_t k0 = S.nbr_ptr[T.n0 + a], k1 = S.nbr_ptr[T.n0 + a + 1];
      for (int64_t k = k0; k < k1; k += 4, ++q) {
        const int cnt = (int)(m[4 * q + 3] >> 16), first = (int)(m[4 * q + 3] & 0xffffu);
        if ((int)m[4 * q + 2] != a || first != (int)(k - k0) || cnt != (int)std::min<int64_t>(4, k1 - k)) ++bad;
        for (int e = 0; e < 4; ++e) {
          const int loc = (int)((m[4 * q + e / 2] >> (16 * (e & 1))) & 0xffffu);
          if (loc >= T.nuq) { ++bad; continue; }
          if (e < cnt && P.uniq_xoff[T.u0 + loc] != (int)S.node_xoff(S.nbr[k + e])) ++bad;
        }
      }
      qstart.push_back(q);
    }
    if (q != T.nquad) ++bad;
    for (; q < T.NQ; ++q) if (m[4 * q + 2] != 0xffffu) ++bad;
    if (T.split[0] != 0 || T.split[VS_CONSUMERS] != T.nquad) ++bad;
    for (int w = 0; w < VS_CONSUMERS; ++w) {
      if (T.split[w] > T.split[w + 1]) ++bad;
      if (!std::binary_search(qstart.begin(), qstart.end(), (int)T.split[w])) ++bad;      // node-aligned
    }
    off += T.NQ;
  }
  if (next_node != S.nn_own || off != V.total_nq) ++bad;
  return bad;
}

std::string build_coarse(const Structure& S, CoarseLevel& C) {
  C = CoarseLevel();
  static const int lines2[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  static const int lines3[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  const int dim = S.dim, NV = S.NV, NN = S.NN, NL = NN - NV;
  Structure& Sc = C.Sc;
  Sc.dim = dim; Sc.NV = NV; Sc.NN = NN; Sc.DPC = S.DPC;
  Sc.rank = S.rank; Sc.nranks = S.nranks;
  Sc.nn_own = S.np_own; Sc.nn_ghost = S.np_ghost; Sc.np_own = 0; Sc.np_ghost = 0; Sc.nc = 0;
  Sc.node_gid = S.pid_gid;
  Sc.nbr_ptr.assign(S.np_own + 1, 0);
  for (int P = 0; P < S.np_own; ++P) {
    const int A = S.pid_node[P];
    Sc.nbr_ptr[P + 1] = Sc.nbr_ptr[P] + (S.pnbr_ptr[A + 1] - S.pnbr_ptr[A]);
  }
  Sc.nbr.resize((size_t)Sc.nbr_ptr[S.np_own]);
  Sc.selfrank.assign(S.np_own, 0);
  for (int P = 0; P < S.np_own; ++P) {
    const int A = S.pid_node[P];
    std::copy(S.pnbr.begin() + S.pnbr_ptr[A], S.pnbr.begin() + S.pnbr_ptr[A + 1], Sc.nbr.begin() + Sc.nbr_ptr[P]);
    Sc.selfrank[P] = S.pselfrank[P];
  }
  Sc.pnbr_ptr.assign(S.np_own + 1, 0);
  Sc.node_pid.assign(S.np_own + S.np_ghost, -1);
  // halo plan of the coarse vectors = the pressure halo
  Sc.peer = S.peer;
  Sc.send_nodes = S.send_pids;
  Sc.send_pids.assign(S.peer.size(), std::vector<int>());
  Sc.recv_node_count = S.recv_pid_count;
  Sc.recv_pid_count.assign(S.peer.size(), 0);
  // line nodes at every owned vertex, end vertices of every owned node
  std::vector<std::pair<int, int>> ve;
  C.node_ends.assign((size_t)S.nn_own * 2, -1);
  for (int c = 0; c < S.nc; ++c) {
    const int* cn = S.cell_nodes.data() + (size_t)c * NN;
    const int* cp = S.cell_pids.data() + (size_t)c * NV;
    for (int v = 0; v < NV; ++v)
      if (cn[v] < S.nn_own) { C.node_ends[2 * (size_t)cn[v]] = cp[v]; C.node_ends[2 * (size_t)cn[v] + 1] = cp[v]; }
    for (int l = 0; l < NL; ++l) {
      const int i = dim == 2 ? lines2[l][0] : lines3[l][0], j = dim == 2 ? lines2[l][1] : lines3[l][1];
      const int e = cn[NV + l];
      if (e < S.nn_own) { C.node_ends[2 * (size_t)e] = cp[i]; C.node_ends[2 * (size_t)e + 1] = cp[j]; }
      if (cp[i] < S.np_own) ve.emplace_back(cp[i], e);
      if (cp[j] < S.np_own) ve.emplace_back(cp[j], e);
    }
  }
  for (int A = 0; A < S.nn_own; ++A)
    if (C.node_ends[2 * (size_t)A] < 0) return "an owned node belongs to no local cell";
  std::sort(ve.begin(), ve.end());
  ve.erase(std::unique(ve.begin(), ve.end()), ve.end());
  C.vedge_ptr.assign(S.np_own + 1, 0);
  for (auto& pr : ve) C.vedge_ptr[pr.first + 1]++;
  for (int P = 0; P < S.np_own; ++P) C.vedge_ptr[P + 1] += C.vedge_ptr[P];
  C.vedge.resize(ve.size());
  for (size_t k = 0; k < ve.size(); ++k) C.vedge[k] = ve[k].second;
  return "";
}

int64_t verify_tile_plan(const Structure& S, const TileLimits& L, const TilePlan& P) {
  int64_t bad = 0;
  const int nt = P.n_tiles();
  if (nt < 0 || P.node_ptr.empty() || P.node_ptr.front() != 0 || P.node_ptr.back() != S.nn_own) return 1;
  if ((int)P.uniq_ptr.size() != nt + 1 || (int)P.puniq_ptr.size() != nt + 1) return 1;
  std::vector<char> seen(nt, 0);
  for (int t : P.tiles_int) { if (t < 0 || t >= nt || seen[t]) ++bad; else seen[t] = 1; }
  for (int t : P.tiles_bnd) { if (t < 0 || t >= nt || seen[t]) ++bad; else seen[t] = 2; }
  for (int t = 0; t < nt; ++t) if (!seen[t]) ++bad;
  if (bad) return bad;
  for (int t = 0; t < nt; ++t) {
    const int n0 = P.node_ptr[t], n1 = P.node_ptr[t + 1];
    const int u0 = P.uniq_ptr[t], nu = P.uniq_ptr[t + 1] - u0, q0 = P.puniq_ptr[t], nq = P.puniq_ptr[t + 1] - q0;
    if (n1 <= n0 || n1 - n0 > L.max_nodes || nu > L.max_uniq || nq > L.max_puniq) ++bad;
    int64_t idx = 0;
    bool ghost = false;
    for (int i = 1; i < nu; ++i) if (P.uniq_xoff[u0 + i] <= P.uniq_xoff[u0 + i - 1]) ++bad;      // memory order, unique
    for (int i = 1; i < nq; ++i) if (P.puniq_xoff[q0 + i] <= P.puniq_xoff[q0 + i - 1]) ++bad;
    for (int i = 0; i < nu; ++i) ghost |= P.uniq_xoff[u0 + i] >= S.n_own_dofs();
    if ((seen[t] == 2) != ghost) ++bad;
    for (int A = n0; A < n1; ++A) {
      idx += (S.nbr_ptr[A + 1] - S.nbr_ptr[A]) + (S.pnbr_ptr[A + 1] - S.pnbr_ptr[A]);
      for (int64_t k = S.nbr_ptr[A]; k < S.nbr_ptr[A + 1]; ++k) {
        const int loc = P.nbr_loc[k];
        if (loc >= nu || P.uniq_xoff[u0 + loc] != (int)S.node_xoff(S.nbr[k])) ++bad;
      }
      for (int64_t k = S.pnbr_ptr[A]; k < S.pnbr_ptr[A + 1]; ++k) {
        const int loc = P.pnbr_loc[k];
        if (loc >= nq || P.puniq_xoff[q0 + loc] != (int)S.pid_xoff(S.pnbr[k])) ++bad;
      }
    }
    if (idx > L.max_idx) ++bad;
  }
  return bad;
}

}  // namespace nsb
