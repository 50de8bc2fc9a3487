#include "mesh.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace nsb_host {

Mesh read_msh(const std::string& path, int dim) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::runtime_error("Could not open mesh file: " + path);
  // pre-pass of the reference: strip '\r', turn $ParametricNodes into $Nodes keeping `id x y z`
  std::stringstream ss;
  std::string line;
  bool in_param = false, first = false;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line == "$ParametricNodes") { ss << "$Nodes\n"; in_param = true; first = true; }
    else if (line == "$EndParametricNodes") { ss << "$EndNodes\n"; in_param = false; }
    else if (in_param) {
      if (first) { ss << line << "\n"; first = false; }
      else {
        std::istringstream is(line);
        long id; double x, y, z;
        is >> id >> x >> y >> z;
        ss.precision(17);
        ss << id << " " << x << " " << y << " " << z << "\n";
      }
    } else ss << line << "\n";
  }
  Mesh m;
  m.dim = dim;
  std::unordered_map<long, uint32_t> id2idx;
  const int ctype = dim == 3 ? 4 : 2, ftype = dim == 3 ? 2 : 1;
  while (std::getline(ss, line)) {
    if (line == "$MeshFormat") {
      double ver; int ft, ds;
      ss >> ver >> ft >> ds;
      if (ver < 2.0 || ver >= 3.0 || ft != 0) throw std::runtime_error("only Gmsh MSH 2.x ASCII is supported: " + path);
    } else if (line == "$Nodes") {
      long n; ss >> n;
      m.points.resize((size_t)n * dim);
      for (long k = 0; k < n; ++k) {
        long id; double x[3];
        ss >> id >> x[0] >> x[1] >> x[2];
        id2idx[id] = (uint32_t)k;
        for (int d = 0; d < dim; ++d) m.points[(size_t)k * dim + d] = x[d];
      }
    } else if (line == "$Elements") {
      long n; ss >> n;
      for (long k = 0; k < n; ++k) {
        long id; int type, ntags;
        ss >> id >> type >> ntags;
        int tag = 0;
        for (int t = 0; t < ntags; ++t) { int v; ss >> v; if (t == 0) tag = v; }
        int nn = type == 1 ? 2 : type == 2 ? 3 : type == 4 ? 4 : type == 15 ? 1 : -1;
        if (nn < 0) throw std::runtime_error("unsupported gmsh element type in " + path);
        long v[4];
        for (int i = 0; i < nn; ++i) ss >> v[i];
        if (type == ctype) {
          for (int i = 0; i < nn; ++i) m.cells.push_back(id2idx.at(v[i]));
          m.cell_tag.push_back(tag);
        } else if (type == ftype) {
          for (int i = 0; i < nn; ++i) m.faces.push_back(id2idx.at(v[i]));
          m.face_tag.push_back(tag);
        }
      }
    }
  }
  if (m.points.empty() || m.cells.empty()) throw std::runtime_error("mesh file has no nodes or no cells: " + path);
  return m;
}

Mesh read_bin(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw std::runtime_error("Could not open mesh file: " + path);
  char magic[8];
  in.read(magic, 8);
  if (std::memcmp(magic, "NSBMESH1", 8) != 0) throw std::runtime_error("not an NSBMESH1 file: " + path);
  int32_t hdr[4];
  in.read((char*)hdr, sizeof(hdr));
  Mesh m;
  m.dim = hdr[0];
  const int64_t V = hdr[1], C = hdr[2], F = hdr[3];
  m.points.resize((size_t)V * m.dim);
  in.read((char*)m.points.data(), m.points.size() * sizeof(double));
  std::vector<int32_t> tmp((size_t)C * (m.dim + 1));
  in.read((char*)tmp.data(), tmp.size() * 4);
  m.cells.assign(tmp.begin(), tmp.end());
  m.cell_tag.resize(C);
  in.read((char*)m.cell_tag.data(), (size_t)C * 4);
  tmp.resize((size_t)F * m.dim);
  in.read((char*)tmp.data(), tmp.size() * 4);
  m.faces.assign(tmp.begin(), tmp.end());
  m.face_tag.resize(F);
  in.read((char*)m.face_tag.data(), (size_t)F * 4);
  if (!in) throw std::runtime_error("truncated mesh file: " + path);
  return m;
}

Mesh read_mesh(const std::string& path, int dim) {
  if (path.size() > 4 && path.substr(path.size() - 4) == ".bin") {
    Mesh m = read_bin(path);
    if (m.dim != dim) throw std::runtime_error("mesh dimension mismatch: " + path);
    return m;
  }
  return read_msh(path, dim);
}

namespace {
struct FaceKey {
  uint32_t v[3];
  bool operator<(const FaceKey& o) const { return std::lexicographical_compare(v, v + 3, o.v, o.v + 3); }
};
FaceKey make_key(int dim, const uint32_t* fv) {
  FaceKey k{{fv[0], fv[1], dim == 3 ? fv[2] : 0xffffffffu}};
  if (k.v[0] > k.v[1]) std::swap(k.v[0], k.v[1]);
  if (dim == 3) {
    if (k.v[1] > k.v[2]) std::swap(k.v[1], k.v[2]);
    if (k.v[0] > k.v[1]) std::swap(k.v[0], k.v[1]);
  }
  return k;
}
}  // namespace

extern "C" {
// METIS 5 API of the toolkit's static library (no header shipped): idx_t = int64_t, real_t = float
int METIS_SetDefaultOptions(int64_t* options);
int METIS_PartGraphRecursive(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize, int64_t* adjwgt,
                             int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval, int64_t* part);
int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize, int64_t* adjwgt,
                        int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval, int64_t* part);
}

std::vector<int32_t> partition_cells(const Mesh& m, int nparts, int method) {
  const int64_t C = m.n_cells();
  std::vector<int32_t> part((size_t)C, 0);
  if (nparts <= 1) return part;
  if (method == 0) {
    for (int64_t c = 0; c < C; ++c) part[c] = (int32_t)((c * (int64_t)nparts) / C);
    return part;
  }
  // face-dual graph: two cells are adjacent when they share a face
  const int dim = m.dim, NV = dim + 1;
  std::vector<std::pair<FaceKey, int64_t>> all;
  all.reserve((size_t)C * NV);
  for (int64_t c = 0; c < C; ++c)
    for (int f = 0; f < NV; ++f) {
      uint32_t fv[3] = {0, 0, 0};
      const int* lv = face_vertices(dim, f);
      for (int i = 0; i < dim; ++i) fv[i] = m.cells[(size_t)c * NV + lv[i]];
      all.push_back({make_key(dim, fv), c});
    }
  std::sort(all.begin(), all.end());
  std::vector<int64_t> deg((size_t)C + 1, 0);
  for (size_t i = 0; i + 1 < all.size(); ++i)
    if (!(all[i].first < all[i + 1].first) && !(all[i + 1].first < all[i].first)) { deg[all[i].second + 1]++; deg[all[i + 1].second + 1]++; }
  for (int64_t c = 0; c < C; ++c) deg[c + 1] += deg[c];
  std::vector<int64_t> adj((size_t)deg[C]), fill(deg.begin(), deg.end() - 1);
  for (size_t i = 0; i + 1 < all.size(); ++i)
    if (!(all[i].first < all[i + 1].first) && !(all[i + 1].first < all[i].first)) {
      adj[fill[all[i].second]++] = all[i + 1].second;
      adj[fill[all[i + 1].second]++] = all[i].second;
    }
  for (int64_t c = 0; c < C; ++c) std::sort(adj.begin() + deg[c], adj.begin() + deg[c + 1]);
  int64_t nv = C, ncon = 1, np = nparts, objval = 0;
  int64_t options[64];
  METIS_SetDefaultOptions(options);
  std::vector<int64_t> p64((size_t)C, 0);
  const int rc = (nparts <= 8 ? METIS_PartGraphRecursive : METIS_PartGraphKway)(&nv, &ncon, deg.data(), adj.data(), nullptr, nullptr, nullptr, &np,
                                                                                nullptr, nullptr, options, &objval, p64.data());
  if (rc != 1) throw std::runtime_error("METIS partitioning failed");
  for (int64_t c = 0; c < C; ++c) part[c] = (int32_t)p64[c];
  return part;
}

std::vector<BoundaryFace> boundary_faces(const Mesh& m) {
  const int dim = m.dim, NV = dim + 1;
  // count face occurrences
  std::vector<std::pair<FaceKey, std::pair<int64_t, int>>> all;
  all.reserve((size_t)m.n_cells() * NV);
  for (int64_t c = 0; c < m.n_cells(); ++c)
    for (int f = 0; f < NV; ++f) {
      uint32_t fv[3] = {0, 0, 0};
      const int* lv = face_vertices(dim, f);
      for (int i = 0; i < dim; ++i) fv[i] = m.cells[(size_t)c * NV + lv[i]];
      all.push_back({make_key(dim, fv), {c, f}});
    }
  std::sort(all.begin(), all.end(), [](const auto& a, const auto& b) {
    if (a.first < b.first) return true;
    if (b.first < a.first) return false;
    return a.second < b.second;
  });
  std::map<FaceKey, int> tagged;
  for (int64_t f = 0; f < m.n_faces(); ++f) tagged[make_key(dim, &m.faces[(size_t)f * dim])] = m.face_tag[f];
  std::vector<BoundaryFace> out;
  for (size_t i = 0; i < all.size();) {
    size_t j = i + 1;
    while (j < all.size() && !(all[i].first < all[j].first) && !(all[j].first < all[i].first)) ++j;
    if (j - i == 1) {
      BoundaryFace b;
      b.cell = all[i].second.first;
      b.face = all[i].second.second;
      auto it = tagged.find(all[i].first);
      b.id = it == tagged.end() ? 0 : it->second;
      const int* lv = face_vertices(dim, b.face);
      b.v = {0, 0, 0};
      for (int k = 0; k < dim; ++k) b.v[k] = m.cells[(size_t)b.cell * NV + lv[k]];
      out.push_back(b);
    }
    i = j;
  }
  // cell order, then face order: the traversal order of the reference's cell loops
  std::sort(out.begin(), out.end(), [](const BoundaryFace& a, const BoundaryFace& b) {
    return a.cell != b.cell ? a.cell < b.cell : a.face < b.face;
  });
  return out;
}

bool assign_boundary_ids_geometrically(const Mesh& m, std::vector<BoundaryFace>& bf, int inlet, int outlet, int wall,
                                       int cylinder) {
  bool hi = false, ho = false, hw = false, hc = false;
  for (auto& b : bf) { hi |= b.id == inlet; ho |= b.id == outlet; hw |= b.id == wall; hc |= b.id == cylinder; }
  if (hi && ho && hw && hc) return false;
  const int dim = m.dim;
  const double tol = 1e-6, cx = 0.2, cy = 0.2, cz = 0.45, L = 2.2, r_cyl = 0.05;
  for (auto& b : bf) {
    double ctr[3] = {0, 0, 0};
    for (int k = 0; k < dim; ++k)
      for (int d = 0; d < dim; ++d) ctr[d] += m.points[(size_t)b.v[k] * dim + d] / dim;
    if (dim == 2) {
      const double dist = std::sqrt((ctr[0] - cx) * (ctr[0] - cx) + (ctr[1] - cy) * (ctr[1] - cy));
      if (dist < r_cyl + 0.02) b.id = cylinder;
      else if (std::fabs(ctr[0]) < tol) b.id = inlet;
      else if (std::fabs(ctr[0] - L) < tol) b.id = outlet;
      else b.id = wall;
    } else {
      const double dist = std::sqrt((ctr[1] - cy) * (ctr[1] - cy) + (ctr[2] - cz) * (ctr[2] - cz));
      if (dist < r_cyl + 0.02) b.id = cylinder;
      else if (std::fabs(ctr[2]) < tol) b.id = inlet;
      else if (std::fabs(ctr[2] - L) < tol) b.id = outlet;     // NB the reference tests z against L = 2.2 (cpp:141,181)
      else b.id = wall;
    }
  }
  return true;
}

}  // namespace nsb_host
