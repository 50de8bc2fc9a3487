#include "dof_handler.hpp"

#include <algorithm>
#include <stdexcept>

namespace nsb_host {

namespace {
const int kLines2[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int kLines3[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
}  // namespace

void DofHandler::distribute(const Mesh& m) {
  dim = m.dim;
  const int NV = dim + 1, NL = dim == 2 ? 3 : 6;
  dofs_per_cell = dim * (NV + NL) + NV;
  const int64_t V = m.n_vertices(), C = m.n_cells();
  const int(*lines)[2] = dim == 2 ? kLines2 : kLines3;
  // edge table: key -> id via sort
  std::vector<uint64_t> keys((size_t)C * NL);
  for (int64_t c = 0; c < C; ++c)
    for (int l = 0; l < NL; ++l) {
      uint64_t a = m.cells[(size_t)c * NV + lines[l][0]], b = m.cells[(size_t)c * NV + lines[l][1]];
      if (a > b) std::swap(a, b);
      keys[(size_t)c * NL + l] = a * (uint64_t)V + b;
    }
  std::vector<uint64_t> uk(keys);
  std::sort(uk.begin(), uk.end());
  uk.erase(std::unique(uk.begin(), uk.end()), uk.end());
  const int64_t E = (int64_t)uk.size();
  // first-touch enumeration: per cell its vertices (dim+1 DoFs each), then its lines (dim DoFs each)
  const uint32_t INV = 0xffffffffu;
  std::vector<uint32_t> vraw(V, INV), eraw(E, INV);
  std::vector<int> cell_edge((size_t)C * NL);
  uint32_t next = 0;
  for (int64_t c = 0; c < C; ++c) {
    for (int v = 0; v < NV; ++v) {
      uint32_t& r = vraw[m.cells[(size_t)c * NV + v]];
      if (r == INV) { r = next; next += dim + 1; }
    }
    for (int l = 0; l < NL; ++l) {
      const int e = (int)(std::lower_bound(uk.begin(), uk.end(), keys[(size_t)c * NL + l]) - uk.begin());
      cell_edge[(size_t)c * NL + l] = e;
      if (eraw[e] == INV) { eraw[e] = next; next += dim; }
    }
  }
  for (int64_t v = 0; v < V; ++v)
    if (vraw[v] == INV) throw std::runtime_error("mesh has a vertex that belongs to no cell");
  const int64_t n_raw = next;
  // stable component-wise renumbering: velocity DoFs keep their order in [0,n_u), pressure in [n_u, n)
  std::vector<unsigned char> is_p(n_raw, 0);
  for (int64_t v = 0; v < V; ++v) is_p[vraw[v] + dim] = 1;
  std::vector<uint32_t> renum(n_raw);
  n_p = V;
  n_u = n_raw - V;
  uint32_t iu = 0, ip = (uint32_t)n_u;
  for (int64_t i = 0; i < n_raw; ++i) renum[i] = is_p[i] ? ip++ : iu++;
  // tables
  vertex_dof0.resize(V); vertex_pdof.resize(V);
  for (int64_t v = 0; v < V; ++v) { vertex_dof0[v] = renum[vraw[v]]; vertex_pdof[v] = renum[vraw[v] + dim]; }
  edge_dof0.resize(E);
  for (int64_t e = 0; e < E; ++e) edge_dof0[e] = {uk[e], renum[eraw[e]]};
  cell_dofs.resize((size_t)C * dofs_per_cell);
  support_points.assign((size_t)(n_u + n_p) * dim, 0.0);
  component.assign(n_u + n_p, 0);
  for (int64_t c = 0; c < C; ++c) {
    uint32_t* d = &cell_dofs[(size_t)c * dofs_per_cell];
    int k = 0;
    for (int v = 0; v < NV; ++v) {
      const uint32_t vid = m.cells[(size_t)c * NV + v];
      for (int comp = 0; comp <= dim; ++comp) {
        const uint32_t g = comp < dim ? vertex_dof0[vid] + comp : vertex_pdof[vid];
        d[k++] = g;
        component[g] = (unsigned char)comp;
        for (int x = 0; x < dim; ++x) support_points[(size_t)g * dim + x] = m.points[(size_t)vid * dim + x];
      }
    }
    for (int l = 0; l < NL; ++l) {
      const uint32_t a = m.cells[(size_t)c * NV + lines[l][0]], b = m.cells[(size_t)c * NV + lines[l][1]];
      const uint32_t g0 = edge_dof0[cell_edge[(size_t)c * NL + l]].second;
      for (int comp = 0; comp < dim; ++comp) {
        d[k++] = g0 + comp;
        component[g0 + comp] = (unsigned char)comp;
        for (int x = 0; x < dim; ++x)
          support_points[(size_t)(g0 + comp) * dim + x] = 0.5 * (m.points[(size_t)a * dim + x] + m.points[(size_t)b * dim + x]);
      }
    }
  }
}

uint32_t DofHandler::edge_first_dof(uint32_t a, uint32_t b, int64_t V) const {
  if (a > b) std::swap(a, b);
  const uint64_t key = (uint64_t)a * (uint64_t)V + b;
  auto it = std::lower_bound(edge_dof0.begin(), edge_dof0.end(), std::make_pair(key, (uint32_t)0));
  if (it == edge_dof0.end() || it->first != key) throw std::runtime_error("boundary face edge not found in the mesh");
  return it->second;
}

namespace {
template <typename Sink>
void visit_boundary_dofs(const Mesh& m, const DofHandler& dh, const std::vector<BoundaryFace>& bf, int id, bool velocity,
                         bool pressure, Sink sink) {
  const int dim = m.dim;
  const int64_t V = m.n_vertices();
  for (const auto& b : bf) {
    if (b.id != id) continue;
    for (int k = 0; k < dim; ++k) {
      if (velocity)
        for (int c = 0; c < dim; ++c) sink(dh.vertex_dof0[b.v[k]] + c);
      if (pressure) sink(dh.vertex_pdof[b.v[k]]);
    }
    if (velocity)
      for (int i = 0; i < dim; ++i)
        for (int j = i + 1; j < dim; ++j) {
          const uint32_t g0 = dh.edge_first_dof(b.v[i], b.v[j], V);
          for (int c = 0; c < dim; ++c) sink(g0 + c);
        }
  }
}
}  // namespace

void interpolate_boundary_values(const Mesh& m, const DofHandler& dh, const std::vector<BoundaryFace>& bf, int id,
                                 const std::function<double(const double*, int)>& value, bool velocity, bool pressure,
                                 Constraints& c) {
  // one call = one boundary_values map, then lines are added where none exists yet
  std::map<uint32_t, double> bv;
  visit_boundary_dofs(m, dh, bf, id, velocity, pressure, [&](uint32_t g) {
    bv[g] = value(&dh.support_points[(size_t)g * dh.dim], dh.component[g]);
  });
  for (auto& kv : bv) c.add_if_new(kv.first, kv.second);
}

void interpolate_boundary_values_map(const Mesh& m, const DofHandler& dh, const std::vector<BoundaryFace>& bf, int id,
                                     const std::function<double(const double*, int)>& value,
                                     std::map<uint32_t, double>& out) {
  visit_boundary_dofs(m, dh, bf, id, true, false, [&](uint32_t g) {
    out[g] = value(&dh.support_points[(size_t)g * dh.dim], dh.component[g]);
  });
}

void make_sparsity_pattern(const DofHandler& dh, std::vector<int64_t>& rowptr, std::vector<uint32_t>& col) {
  const int64_t N = dh.n_dofs(), C = (int64_t)dh.cell_dofs.size() / dh.dofs_per_cell;
  const int K = dh.dofs_per_cell;
  // dof -> cells
  std::vector<int64_t> ptr(N + 1, 0);
  for (size_t i = 0; i < dh.cell_dofs.size(); ++i) ptr[dh.cell_dofs[i] + 1]++;
  for (int64_t i = 0; i < N; ++i) ptr[i + 1] += ptr[i];
  std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
  std::vector<uint32_t> d2c(dh.cell_dofs.size());
  for (int64_t c = 0; c < C; ++c)
    for (int k = 0; k < K; ++k) d2c[fill[dh.cell_dofs[(size_t)c * K + k]]++] = (uint32_t)c;
  rowptr.assign(N + 1, 0);
  std::vector<std::vector<uint32_t>> rows(N);
#pragma omp parallel
  {
    std::vector<uint32_t> buf;
#pragma omp for schedule(dynamic, 2048)
    for (int64_t i = 0; i < N; ++i) {
      buf.clear();
      for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) {
        const uint32_t* d = &dh.cell_dofs[(size_t)d2c[p] * K];
        buf.insert(buf.end(), d, d + K);
      }
      std::sort(buf.begin(), buf.end());
      buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
      rows[i] = buf;
    }
  }
  for (int64_t i = 0; i < N; ++i) rowptr[i + 1] = rowptr[i] + (int64_t)rows[i].size();
  col.resize(rowptr[N]);
  for (int64_t i = 0; i < N; ++i) std::copy(rows[i].begin(), rows[i].end(), col.begin() + rowptr[i]);
}

}  // namespace nsb_host
