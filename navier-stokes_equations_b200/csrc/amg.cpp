// See amg.hpp.  Host code, runs once per mesh (K_p is constant in time).
#include "amg.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace nsb {

HostCsr transpose(const HostCsr& A) {
  HostCsr T;
  T.n = A.m; T.m = A.n;
  T.ptr.assign(A.m + 1, 0);
  for (int c : A.col) T.ptr[c + 1]++;
  for (int i = 0; i < A.m; ++i) T.ptr[i + 1] += T.ptr[i];
  T.col.resize(A.col.size()); T.val.resize(A.col.size());
  std::vector<int> fill(T.ptr.begin(), T.ptr.end() - 1);
  for (int i = 0; i < A.n; ++i)
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int p = fill[A.col[k]]++;
      T.col[p] = i; T.val[p] = A.val[k];
    }
  return T;
}

HostCsr spgemm(const HostCsr& A, const HostCsr& B) {
  HostCsr C;
  C.n = A.n; C.m = B.m;
  std::vector<std::vector<std::pair<int, double>>> rows(A.n);
#pragma omp parallel
  {
    std::vector<std::pair<int, double>> buf;
#pragma omp for schedule(dynamic, 512)
    for (int i = 0; i < A.n; ++i) {
      buf.clear();
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int j = A.col[k];
        const double a = A.val[k];
        if (a == 0.0) continue;
        for (int l = B.ptr[j]; l < B.ptr[j + 1]; ++l) buf.emplace_back(B.col[l], a * B.val[l]);
      }
      std::sort(buf.begin(), buf.end(), [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
      auto& out = rows[i];
      for (size_t k = 0; k < buf.size();) {
        size_t e = k;
        double s = 0;
        while (e < buf.size() && buf[e].first == buf[k].first) s += buf[e++].second;
        out.emplace_back(buf[k].first, s);
        k = e;
      }
    }
  }
  C.ptr.assign(A.n + 1, 0);
  for (int i = 0; i < A.n; ++i) C.ptr[i + 1] = C.ptr[i] + (int)rows[i].size();
  C.col.resize(C.ptr[A.n]); C.val.resize(C.ptr[A.n]);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A.n; ++i) {
    int p = C.ptr[i];
    for (auto& e : rows[i]) { C.col[p] = e.first; C.val[p++] = e.second; }
  }
  return C;
}

double power_lmax_jacobi(const HostCsr& A, const std::vector<double>& dinv, int iters) {
  const int n = A.n;
  std::vector<double> v(n), w(n);
  // deterministic pseudo-random start
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    v[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5;
  }
  double lam = 1.0;
  for (int it = 0; it < iters; ++it) {
    double nv = 0;
    for (int i = 0; i < n; ++i) nv += v[i] * v[i];
    nv = std::sqrt(nv);
    if (nv == 0) return 1.0;
    for (int i = 0; i < n; ++i) v[i] /= nv;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      double t = 0;
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) t += A.val[k] * v[A.col[k]];
      w[i] = dinv[i] * t;
    }
    double nw = 0;
    for (int i = 0; i < n; ++i) nw += w[i] * w[i];
    lam = std::sqrt(nw);
    v.swap(w);
  }
  return lam;
}

static void dense_inverse(int n, std::vector<double>& a) {
  // Gauss-Jordan with partial pivoting, in place (row-major)
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
    if (piv != c)
      for (int k = 0; k < n; ++k) {
        std::swap(a[(size_t)c * n + k], a[(size_t)piv * n + k]);
        std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]);
      }
    const double d = 1.0 / a[(size_t)c * n + c];
    for (int k = 0; k < n; ++k) { a[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) {
        a[(size_t)r * n + k] -= f * a[(size_t)c * n + k];
        inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k];
      }
    }
  }
  a.swap(inv);
}

void amg_setup(const HostCsr& A0, AmgHierarchy& H, double threshold, int max_coarse, int max_levels) {
  H.levels.clear();
  H.levels.emplace_back();
  H.levels[0].A = A0;
  for (int lev = 0;; ++lev) {
    AmgLevel& L = H.levels[lev];
    const HostCsr& A = L.A;
    const int n = A.n;
    L.dinv.assign(n, 1.0);
    std::vector<double> diag(n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        if (A.col[k] == i) diag[i] = A.val[k];
    for (int i = 0; i < n; ++i) L.dinv[i] = diag[i] != 0.0 ? 1.0 / diag[i] : 1.0;
    L.lmax = power_lmax_jacobi(A, L.dinv);
    if (n <= max_coarse || lev + 1 >= max_levels) break;

    // ---- strength of connection  |a_ij| >= threshold * sqrt(a_ii a_jj)
    std::vector<int> sptr(n + 1, 0), scol;
    scol.reserve(A.col.size());
    for (int i = 0; i < n; ++i) {
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int j = A.col[k];
        if (j == i || A.val[k] == 0.0) continue;
        if (std::fabs(A.val[k]) >= threshold * std::sqrt(std::fabs(diag[i] * diag[j]))) scol.push_back(j);
      }
      sptr[i + 1] = (int)scol.size();
    }
    // ---- greedy aggregation (deterministic: natural order)
    std::vector<int> agg(n, -1);
    int nagg = 0;
    for (int i = 0; i < n; ++i) {                       // pass 1: roots with a fully free neighbourhood
      if (agg[i] != -1 || sptr[i + 1] == sptr[i]) continue;
      bool free_nb = true;
      for (int k = sptr[i]; k < sptr[i + 1]; ++k) free_nb &= (agg[scol[k]] == -1);
      if (!free_nb) continue;
      agg[i] = nagg;
      for (int k = sptr[i]; k < sptr[i + 1]; ++k) agg[scol[k]] = nagg;
      ++nagg;
    }
    {
      std::vector<int> agg1(agg);                       // pass 2: attach to the strongest neighbouring aggregate
      for (int i = 0; i < n; ++i) {
        if (agg[i] != -1 || sptr[i + 1] == sptr[i]) continue;
        double best = -1;
        int who = -1;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
          const int j = A.col[k];
          if (j == i || agg1[j] == -1) continue;
          const double s = std::fabs(A.val[k]) / std::sqrt(std::fabs(diag[i] * diag[j]));
          if (s >= threshold && s > best) { best = s; who = agg1[j]; }
        }
        if (who >= 0) agg[i] = who;
      }
    }
    for (int i = 0; i < n; ++i) {                       // pass 3: leftovers
      if (agg[i] != -1 || sptr[i + 1] == sptr[i]) continue;
      agg[i] = nagg;
      for (int k = sptr[i]; k < sptr[i + 1]; ++k)
        if (agg[scol[k]] == -1) agg[scol[k]] = nagg;
      ++nagg;
    }
    if (nagg == 0 || nagg >= n * 0.8) break;            // coarsening stalled: solve this level directly
    // ---- tentative prolongator (piecewise constant) and its Jacobi smoothing
    HostCsr Pt;
    Pt.n = n; Pt.m = nagg;
    Pt.ptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) Pt.ptr[i + 1] = Pt.ptr[i] + (agg[i] >= 0 ? 1 : 0);
    Pt.col.resize(Pt.ptr[n]); Pt.val.assign(Pt.ptr[n], 1.0);
    for (int i = 0; i < n; ++i)
      if (agg[i] >= 0) Pt.col[Pt.ptr[i]] = agg[i];
    const double omega = (4.0 / 3.0) / L.lmax;
    HostCsr DA = A;                                     // I - omega D^-1 A
    for (int i = 0; i < n; ++i)
      for (int k = DA.ptr[i]; k < DA.ptr[i + 1]; ++k)
        DA.val[k] = ((DA.col[k] == i) ? 1.0 : 0.0) - omega * L.dinv[i] * A.val[k];
    HostCsr P = spgemm(DA, Pt);
    // rows without an aggregate (isolated / Dirichlet rows) must not interpolate
    for (int i = 0; i < n; ++i)
      if (agg[i] < 0)
        for (int k = P.ptr[i]; k < P.ptr[i + 1]; ++k) P.val[k] = 0.0;
    HostCsr R = transpose(P);
    HostCsr AP = spgemm(A, P);
    HostCsr Ac = spgemm(R, AP);
    L.P = std::move(P);
    L.R = std::move(R);
    H.levels.emplace_back();
    H.levels.back().A = std::move(Ac);
  }
  // ---- dense inverse of the coarsest operator
  const HostCsr& Ac = H.levels.back().A;
  const int nc = Ac.n;
  H.coarse_inv.assign((size_t)nc * nc, 0.0);
  for (int i = 0; i < nc; ++i)
    for (int k = Ac.ptr[i]; k < Ac.ptr[i + 1]; ++k) H.coarse_inv[(size_t)i * nc + Ac.col[k]] += Ac.val[k];
  for (int i = 0; i < nc; ++i)
    if (H.coarse_inv[(size_t)i * nc + i] == 0.0) H.coarse_inv[(size_t)i * nc + i] = 1.0;
  dense_inverse(nc, H.coarse_inv);
}

}  // namespace nsb
