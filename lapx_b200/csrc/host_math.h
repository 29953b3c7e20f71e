// host_math.h — small host-side linear algebra and table builders shared by solver.cu and the CPU
// emulation harness in tests/emu (pure C++, no CUDA).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>

#include "evp_core.h"

namespace evp {
namespace host {

static const int kI[6] = {0, 1, 2, 1, 0, 0};
static const int kJ[6] = {0, 1, 2, 2, 2, 1};
static const double kW[6] = {1.0, 1.0, 1.0, 1.41421356237309504880, 1.41421356237309504880, 1.41421356237309504880};

// ---- small host linear algebra -------------------------------------------------------------
inline bool solve_n(int n, double *A, double *b) {  // Gauss with partial pivoting
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[n * r + c]) > std::fabs(A[n * p + c])) p = r;
    if (!(std::fabs(A[n * p + c]) > 0)) return false;
    if (p != c) {
      for (int k = 0; k < n; ++k) std::swap(A[n * c + k], A[n * p + k]);
      std::swap(b[c], b[p]);
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = A[n * r + c] / A[n * c + c];
      for (int k = c; k < n; ++k) A[n * r + k] -= f * A[n * c + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= A[n * r + k] * b[k];
    b[r] = s / A[n * r + r];
  }
  return true;
}
inline bool inv6(const double *A, double *Ai) {
  for (int c = 0; c < 6; ++c) {
    double M[36], r[6] = {0, 0, 0, 0, 0, 0};
    std::memcpy(M, A, sizeof(M));
    r[c] = 1.0;
    if (!solve_n(6, M, r)) return false;
    for (int k = 0; k < 6; ++k) Ai[6 * k + c] = r[k];
  }
  return true;
}
inline void voigt_to_mandel(const double *cv, double *cm) {
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) cm[6 * a + b] = kW[a] * kW[b] * cv[6 * a + b];
}
// Mandel 6x6 -> b-basis 6x6 (M_b = T M T^T), packed upper triangle
inline void mandel_to_bpacked(const double *Mm, double *packed) {
  double T[36] = {0};
  const double r2 = 1.0 / std::sqrt(2.0), r6 = 1.0 / std::sqrt(6.0), r3 = 1.0 / std::sqrt(3.0);
  T[0] = -r2; T[1] = r2;
  T[6] = -r6; T[7] = -r6; T[8] = 2 * r6;
  T[12 + 3] = 1; T[18 + 4] = 1; T[24 + 5] = 1;
  T[30] = r3; T[31] = r3; T[32] = r3;
  double A[36], B[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += T[6 * i + k] * Mm[6 * k + j];
      A[6 * i + j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += A[6 * i + k] * T[6 * j + k];
      B[6 * i + j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) packed[sidx(i, j)] = 0.5 * (B[6 * i + j] + B[6 * j + i]);
}
inline void mandel_vec_to_b(const double *m, double *b) {
  double c[6];
  for (int k = 0; k < 6; ++k) c[k] = m[k] / kW[k];
  cart_to_b(c, b);
}
// Mandel rotation matrix: mandel(R A R^T) = Q mandel(A)
inline void mandel_rotation(const double *R, double *Q) {
  for (int mu = 0; mu < 6; ++mu) {
    double B[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const double v = (mu < 3) ? 1.0 : 1.0 / std::sqrt(2.0);
    B[kI[mu]][kJ[mu]] = v;
    B[kJ[mu]][kI[mu]] = v;
    double RB[3][3], T[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += R[3 * i + k] * B[k][j];
        RB[i][j] = s;
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += RB[i][k] * R[3 * j + k];
        T[i][j] = s;
      }
    for (int la = 0; la < 6; ++la) Q[6 * la + mu] = kW[la] * T[kI[la]][kJ[la]];
  }
}

inline void build_phase_dev(const evp_phase &in, PhaseDev &pd) {
  std::memset(&pd, 0, sizeof(pd));
  pd.nsys = in.nsys;
  pd.nmodes = in.nmodes;
  double Cm[36], Sm[36];
  voigt_to_mandel(in.c_voigt, Cm);
  inv6(Cm, Sm);
  mandel_to_bpacked(Sm, pd.Sc);
  for (int s = 0; s < in.nsys; ++s) {
    double b[3], n[3], bl = 0, nl = 0;
    for (int k = 0; k < 3; ++k) { bl += in.b[s][k] * in.b[s][k]; nl += in.n[s][k] * in.n[s][k]; }
    bl = std::sqrt(bl); nl = std::sqrt(nl);
    for (int k = 0; k < 3; ++k) { b[k] = in.b[s][k] / bl; n[k] = in.n[s][k] / nl; }
    double mc[6], mb[6];
    for (int c = 0; c < 6; ++c) mc[c] = 0.5 * (b[kI[c]] * n[kJ[c]] + b[kJ[c]] * n[kI[c]]);
    cart_to_b(mc, mb);
    for (int c = 0; c < 5; ++c) pd.m[s][c] = mb[c];
    for (int i = 0; i < 5; ++i)
      for (int j = i; j < 5; ++j) pd.mm[s][s5idx(i, j)] = mb[i] * mb[j];
    const int m = in.mode[s];
    pd.mode[s] = m;
    pd.g0[s] = in.gamma0[m];
    pd.nrate[s] = in.nrate[m];
    const double nm1 = in.nrate[m] - 1.0;
    const int ni = (int)nm1;
    pd.npow[s] = ((double)ni == nm1 && ni >= 0 && ni <= 63) ? ni : -1;
    pd.twin[s] = in.twin[m];
    pd.alpha[s][0] = 0.5 * (b[2] * n[1] - n[2] * b[1]);
    pd.alpha[s][1] = 0.5 * (b[0] * n[2] - n[0] * b[2]);
    pd.alpha[s][2] = 0.5 * (b[1] * n[0] - n[1] * b[0]);
    for (int k = 0; k < 3; ++k) pd.nrm[s][k] = n[k];
    pd.itshear[s] = (in.twin[m] && in.twin_shear[m] > 0) ? 1.0 / in.twin_shear[m] : 0.0;
  }
  pd.twin_thr1 = in.twin_thr1;
  pd.twin_thr2 = in.twin_thr2;
  for (int m = 0; m < EVP_MAX_MODES; ++m) {
    pd.tau0[m] = in.tau0[m]; pd.tau1[m] = in.tau1[m]; pd.theta0[m] = in.theta0[m]; pd.theta1[m] = in.theta1[m];
    for (int m2 = 0; m2 < EVP_MAX_MODES; ++m2) pd.hlat[m][m2] = in.hlat[m][m2];
  }
}


// packed b-basis reference compliance + isotropy flag (diag(a,a,a,a,a,b), zero off-diagonal)
inline void build_s0b(const double *S0m, double *S0b, int *iso) {
  mandel_to_bpacked(S0m, S0b);
  double mx = 0, off = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      const double v = std::fabs(S0b[sidx(i, j)]);
      mx = std::max(mx, v);
      if (i != j) off = std::max(off, v);
    }
  double dd = 0;
  for (int i = 1; i < 5; ++i) dd = std::max(dd, std::fabs(S0b[sidx(i, i)] - S0b[sidx(0, 0)]));
  *iso = (off <= 1e-13 * mx && dd <= 1e-13 * mx) ? 1 : 0;
}

// Green-operator constants from the Mandel reference stiffness / compliance
inline void build_green_const(const double *C0m, const double *S0m, GreenConst &G) {
  const int vm[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
  auto C4 = [&](int i, int j, int k, int l) {
    const int a = vm[i][j], b = vm[k][l];
    return C0m[6 * a + b] / (kW[a] * kW[b]);
  };
  for (int q = 0; q < 6; ++q) {
    const int i = kI[q], k = kJ[q];
    for (int m = 0; m < 3; ++m) G.KA[q][m] = C4(i, m, k, m);
    G.KA[q][3] = C4(i, 1, k, 2) + C4(i, 2, k, 1);
    G.KA[q][4] = C4(i, 0, k, 2) + C4(i, 2, k, 0);
    G.KA[q][5] = C4(i, 0, k, 1) + C4(i, 1, k, 0);
  }
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) G.SC[6 * a + b] = S0m[6 * a + b] * kW[b] / kW[a];
}

}  // namespace host
}  // namespace evp
