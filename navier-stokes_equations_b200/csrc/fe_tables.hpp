// Reference-element tables of the Taylor-Hood P2/P1 pair on simplices, in deal.II's
// conventions:
//   FESystem(FE_SimplexP(2)^dim, FE_SimplexP(1))     reference src/classes/NavierStokes.hpp:429-432
//   QGaussSimplex<dim>(3)                            reference src/classes/NavierStokes.hpp:433
// All device kernels read these from __constant__ memory; the host fills them once.
#pragma once
#include <cstring>

namespace nsb {

template <int DIM> struct Fe {
  static constexpr int NV = DIM + 1;                      // vertices / P1 functions
  static constexpr int NL = (DIM == 2) ? 3 : 6;           // lines
  static constexpr int NN = NV + NL;                      // P2 nodes per cell
  static constexpr int DPC = DIM * NN + NV;               // dofs per cell (15 / 34)
  static constexpr int NQ = (DIM == 2) ? 7 : 10;          // QGaussSimplex<dim>(3), deal.II 9.3/9.4
};

// One table block per dimension, padded to the 3-D sizes so a single struct serves both.
struct FeTables {
  int dim, nv, nn, nq, dpc;
  int idx[10][2];          // P2 node a -> barycentric indices (i,j); vertex nodes have i==j
  double lam[16][4];       // lam[q][k]   barycentric coordinates of quadrature point q
  double w[16];            // reference weights (sum 1/2 resp. 1/6)
  double phi[16][10];      // phi[q][a]   P2 values
  double dco[10][2][16];   // dco[a][s][q] = d phi_a / d lambda_{idx[a][s]} at q  (slot 1 of a vertex node is 0)
  double Mhat[10][10];     // sum_q w phi_a phi_b
  double Khat[10][10][2][2];   // sum_q w dco[a][sa] dco[b][sb]
  double Bhat[10][4][2];   // sum_q w dco[a][sa] lam_j
  double MPhat[4][4];      // sum_q w lam_i lam_j      (P1 mass)
};

inline void fill_tables(int dim, FeTables& T) {
  std::memset(&T, 0, sizeof(T));
  T.dim = dim;
  T.nv = dim + 1;
  static const int lines2[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  static const int lines3[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  const int nl = (dim == 2) ? 3 : 6;
  T.nn = T.nv + nl;
  T.dpc = dim * T.nn + T.nv;
  for (int v = 0; v < T.nv; ++v) T.idx[v][0] = T.idx[v][1] = v;
  for (int l = 0; l < nl; ++l) {
    T.idx[T.nv + l][0] = (dim == 2) ? lines2[l][0] : lines3[l][0];
    T.idx[T.nv + l][1] = (dim == 2) ? lines2[l][1] : lines3[l][1];
  }
  if (dim == 2) {
    // 7-point rule with deal.II 9.3's literal constants
    static const double p[7][2] = {{0.3333333333330, 0.3333333333330}, {0.7974269853530, 0.1012865073230},
                                   {0.1012865073230, 0.7974269853530}, {0.1012865073230, 0.1012865073230},
                                   {0.0597158717898, 0.4701420641050}, {0.4701420641050, 0.0597158717898},
                                   {0.4701420641050, 0.4701420641050}};
    static const double w[7] = {0.225, 0.125939180545, 0.125939180545, 0.125939180545,
                                0.132394152789, 0.132394152789, 0.132394152789};
    T.nq = 7;
    for (int q = 0; q < 7; ++q) {
      T.lam[q][0] = 1.0 - (p[q][0] + p[q][1]);
      T.lam[q][1] = p[q][0];
      T.lam[q][2] = p[q][1];
      T.w[q] = 0.5 * w[q];
    }
  } else {
    // 10-point degree-3 Keast rule
    const double a = 0.5684305841968444, b = 0.1438564719343852;
    const double p[10][3] = {{a, b, b}, {b, b, b}, {b, b, a}, {b, a, b}, {0.0, 0.5, 0.5},
                             {0.5, 0.0, 0.5}, {0.5, 0.5, 0.0}, {0.5, 0.0, 0.0}, {0.0, 0.5, 0.0}, {0.0, 0.0, 0.5}};
    T.nq = 10;
    for (int q = 0; q < 10; ++q) {
      T.lam[q][0] = 1.0 - ((p[q][0] + p[q][1]) + p[q][2]);
      T.lam[q][1] = p[q][0];
      T.lam[q][2] = p[q][1];
      T.lam[q][3] = p[q][2];
      T.w[q] = ((q < 4) ? 0.2177650698804054 : 0.0214899534130631) / 6.0;
    }
  }
  for (int q = 0; q < T.nq; ++q)
    for (int a = 0; a < T.nn; ++a) {
      const int i = T.idx[a][0], j = T.idx[a][1];
      if (a < T.nv) {
        T.phi[q][a] = T.lam[q][i] * (2.0 * T.lam[q][i] - 1.0);
        T.dco[a][0][q] = 4.0 * T.lam[q][i] - 1.0;
        T.dco[a][1][q] = 0.0;
      } else {
        T.phi[q][a] = 4.0 * T.lam[q][i] * T.lam[q][j];
        T.dco[a][0][q] = 4.0 * T.lam[q][j];
        T.dco[a][1][q] = 4.0 * T.lam[q][i];
      }
    }
  for (int a = 0; a < T.nn; ++a)
    for (int b = 0; b < T.nn; ++b) {
      double m = 0;
      for (int q = 0; q < T.nq; ++q) m += T.w[q] * T.phi[q][a] * T.phi[q][b];
      T.Mhat[a][b] = m;
      for (int sa = 0; sa < 2; ++sa)
        for (int sb = 0; sb < 2; ++sb) {
          double k = 0;
          for (int q = 0; q < T.nq; ++q) k += T.w[q] * T.dco[a][sa][q] * T.dco[b][sb][q];
          T.Khat[a][b][sa][sb] = k;
        }
    }
  for (int a = 0; a < T.nn; ++a)
    for (int j = 0; j < T.nv; ++j)
      for (int sa = 0; sa < 2; ++sa) {
        double s = 0;
        for (int q = 0; q < T.nq; ++q) s += T.w[q] * T.dco[a][sa][q] * T.lam[q][j];
        T.Bhat[a][j][sa] = s;
      }
  for (int i = 0; i < T.nv; ++i)
    for (int j = 0; j < T.nv; ++j) {
      double s = 0;
      for (int q = 0; q < T.nq; ++q) s += T.w[q] * T.lam[q][i] * T.lam[q][j];
      T.MPhat[i][j] = s;
    }
}

}  // namespace nsb
