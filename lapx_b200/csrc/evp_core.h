// evp_core.h — thread-level math of the CUDA kernels, written as __host__ __device__ inline
// functions so that tests/emu can run exactly the same arithmetic on the CPU (no GPU in the
// build container).  Nothing in here is a CPU fallback: the product library only calls these
// from __global__ kernels (kernels.cu).
//
// Units implemented:  a1/a3 Stockham radix passes, a2 Green-operator point
// function, a4 crystal-frame Newton solve, a5 multiplier identity, a6 norm contributions.
#pragma once

#include <math.h>
#include <stdint.h>

#include <utility>

#include "../../include/evpfft.h"

#if defined(__CUDACC__)
#define EVP_HD __host__ __device__ __forceinline__
#else
#define EVP_HD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace evp {

constexpr double kRSQ2 = 0.70710678118654752440;   // 1/sqrt(2)
constexpr double kSQ2 = 1.41421356237309504880;
constexpr double kRSQ3 = 0.57735026918962576451;
constexpr double kRSQ6 = 0.40824829046386301637;

// ---------------------------------------------------------------------------------------------
// symmetric-tensor bases.  Cartesian order 11,22,33,23,13,12.  b-basis (Lebensohn):
//   b0=(22-11)/sqrt2  b1=(2*33-11-22)/sqrt6  b2=sqrt2*23  b3=sqrt2*13  b4=sqrt2*12  b5=tr/sqrt3
// ---------------------------------------------------------------------------------------------
EVP_HD void cart_to_b(const double c[6], double b[6]) {
  b[0] = (c[1] - c[0]) * kRSQ2;
  b[1] = (2.0 * c[2] - c[0] - c[1]) * kRSQ6;
  b[2] = kSQ2 * c[3];
  b[3] = kSQ2 * c[4];
  b[4] = kSQ2 * c[5];
  b[5] = (c[0] + c[1] + c[2]) * kRSQ3;
}
EVP_HD void b_to_cart(const double b[6], double c[6]) {
  const double h = b[5] * kRSQ3, d1 = b[1] * kRSQ6, d0 = b[0] * kRSQ2;
  c[0] = h - d1 - d0;
  c[1] = h - d1 + d0;
  c[2] = h + 2.0 * d1;
  c[3] = b[2] * kRSQ2;
  c[4] = b[3] * kRSQ2;
  c[5] = b[4] * kRSQ2;
}

// packed upper triangle of a symmetric 6x6: row i starts at i*6 - i(i-1)/2
EVP_HD constexpr int sidx(int i, int j) { return (i <= j) ? (i * 6 - (i * (i - 1)) / 2 + (j - i)) : (j * 6 - (j * (j - 1)) / 2 + (i - j)); }

// 5x5 rotation of deviatoric b-vectors: a_sample = M a_crystal, a_crystal = M^T a_sample,
// M[al*5+be] = b^al : (R b^be R^T), R = crystal->sample, row major.  The 6th (hydrostatic)
// component is invariant.
EVP_HD void rot_b5(const double R[9], double M[25]) {
  // symmetric outer products of the columns r_k of R:  O^{kl}_c, c in Cartesian 6-order
  double O[6][6];  // [pair kl in order 00,11,22,12,02,01][cart comp]
  const int ck[6] = {0, 1, 2, 1, 0, 0}, cl[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    const int k = ck[p], l = cl[p];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const int i = ck[c], j = cl[c];
      O[p][c] = 0.5 * (R[3 * i + k] * R[3 * j + l] + R[3 * i + l] * R[3 * j + k]);
    }
  }
  double t[6], tb[6];
  // b0 -> (O11 - O00)/sqrt2
#pragma unroll
  for (int c = 0; c < 6; ++c) t[c] = (O[1][c] - O[0][c]) * kRSQ2;
  cart_to_b(t, tb);
#pragma unroll
  for (int a = 0; a < 5; ++a) M[a * 5 + 0] = tb[a];
#pragma unroll
  for (int c = 0; c < 6; ++c) t[c] = (2.0 * O[2][c] - O[0][c] - O[1][c]) * kRSQ6;
  cart_to_b(t, tb);
#pragma unroll
  for (int a = 0; a < 5; ++a) M[a * 5 + 1] = tb[a];
#pragma unroll
  for (int q = 0; q < 3; ++q) {  // b2 <- sqrt2*O12, b3 <- sqrt2*O02, b4 <- sqrt2*O01
#pragma unroll
    for (int c = 0; c < 6; ++c) t[c] = kSQ2 * O[3 + q][c];
    cart_to_b(t, tb);
#pragma unroll
    for (int a = 0; a < 5; ++a) M[a * 5 + 2 + q] = tb[a];
  }
}

// ---------------------------------------------------------------------------------------------
// per-phase constant tables used by the constitutive kernel (crystal frame, b-basis)
// ---------------------------------------------------------------------------------------------
struct PhaseDev {
  int32_t nsys;
  int32_t nmodes;
  double Sc[21];                  // crystal compliance, packed
  double m[EVP_MAX_SYS][5];       // Schmid tensors (deviatoric b components)
  double mm[EVP_MAX_SYS][15];     // m (x) m, packed upper triangle of the 5x5 block
  double g0[EVP_MAX_SYS];
  double nrate[EVP_MAX_SYS];
  int32_t npow[EVP_MAX_SYS];      // n-1 when n is an integer in [1,64], else -1 -> pow()
  int32_t twin[EVP_MAX_SYS];
  int32_t mode[EVP_MAX_SYS];
  double alpha[EVP_MAX_SYS][3];   // skew part of b(x)n, axial (32,13,21), crystal frame
  // hardening (commit kernel)
  double tau0[EVP_MAX_MODES], tau1[EVP_MAX_MODES], theta0[EVP_MAX_MODES], theta1[EVP_MAX_MODES];
  double hlat[EVP_MAX_MODES][EVP_MAX_MODES];
  double nrm[EVP_MAX_SYS][3];     // unit plane normals (crystal frame): twin reorientation 2 n n^T - I
  double itshear[EVP_MAX_SYS];    // 1 / characteristic twin shear of the system's mode (0 for slip)
  double twin_thr1, twin_thr2;
};

EVP_HD constexpr int s5idx(int i, int j) { return (i <= j) ? (i * 5 - (i * (i - 1)) / 2 + (j - i)) : (j * 5 - (j * (j - 1)) / 2 + (i - j)); }

// x^(n-1) for the power law; integer exponents by binary powering (same tree as the oracle's
// x^(n-1); the oracle multiplies the same sequence).
EVP_HD double pow_nm1(double x, int npow, double nrate) {
  if (npow >= 0) {
    double r = 1.0, b = x;
    int k = npow;
    while (k) {
      if (k & 1) r *= b;
      b *= b;
      k >>= 1;
    }
    return r;
  }
  return pow(x, nrate - 1.0);
}

// shear rate and tangent of one system at resolved shear stress tau, itc = 1/tau_c
EVP_HD void slip_rate(const PhaseDev &P, int s, double tau, double itc, double &gd, double &dgd) {
  const double x = fabs(tau) * itc;
  const double xn1 = pow_nm1(x, P.npow[s], P.nrate[s]);
  const bool off = (P.twin[s] != 0) && (tau <= 0.0);
  const double g = off ? 0.0 : P.g0[s] * xn1;
  gd = g * x * (tau >= 0.0 ? 1.0 : -1.0);
  dgd = g * P.nrate[s] * itc;
}

// in-place LDL^T of a packed symmetric 6x6 and solve; returns false on a non-positive pivot
EVP_HD bool ldl6_solve(double a[21], double b[6]) {
  double invd[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double w[6];
    double d = a[sidx(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) {
      w[k] = a[sidx(k, j)] * a[sidx(k, k)];
      d -= a[sidx(k, j)] * w[k];
    }
    a[sidx(j, j)] = d;
    ok = ok && (d > 0.0);
    const double inv = 1.0 / d;
    invd[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = a[sidx(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= a[sidx(k, i)] * w[k];
      a[sidx(j, i)] = t * inv;
    }
  }
#pragma unroll
  for (int i = 1; i < 6; ++i) {
#pragma unroll
    for (int k = 0; k < i; ++k) b[i] -= a[sidx(k, i)] * b[k];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] *= invd[i];
#pragma unroll
  for (int i = 4; i >= 0; --i) {
#pragma unroll
    for (int k = i + 1; k < 6; ++k) b[i] -= a[sidx(i, k)] * b[k];
  }
  return ok;
}

// x^K with a compile-time exponent (square-and-multiply, K = n-1 of the power law)
template <int K>
EVP_HD double pow_ct(double x) {
  if constexpr (K == 0) return 1.0;
  else if constexpr (K == 1) return x;
  else {
    const double h = pow_ct<K / 2>(x);
    return (K & 1) ? h * h * x : h * h;
  }
}

// NPOW_T >= 0: every system uses the compile-time exponent n-1 = NPOW_T;  NPOW_T < 0: per-system tables
template <int NPOW_T>
EVP_HD void slip_rate_t(const PhaseDev &P, int s, double tau, double itc, double &gd, double &dgd) {
  const double x = fabs(tau) * itc;
  const double xn1 = (NPOW_T >= 0) ? pow_ct<(NPOW_T >= 0 ? NPOW_T : 0)>(x) : pow_nm1(x, P.npow[s], P.nrate[s]);
  const bool off = (P.twin[s] != 0) && (tau <= 0.0);
  const double g = off ? 0.0 : P.g0[s] * xn1;
  gd = g * x * (tau >= 0.0 ? 1.0 : -1.0);
  dgd = g * P.nrate[s] * itc;
}

// Row a4: Newton solve of  Jb*s + dt*edp(s) = g  in the crystal frame (b-basis).  NS_T > 0: the loop over systems is fully unrolled (tables become
// constant-bank operands);  NS_T == 0: runtime P.nsys.  JB(k)/GV(i): accessors of the packed
// Jb = S0_c + S_c and of g (kept in shared memory by the kernel, in arrays by the emulation).
template <int NS_T, int NPOW_T, class JB, class GV, class ITC>
EVP_HD int newton_crystal_t(const PhaseDev &P, JB Jb, GV g, double s[6], double dt, double tol, int itmax, ITC itc, int *bad) {
  int it = 0;
  bool conv = false;
  while (it < itmax) {
    double A[15], F[6];
#pragma unroll
    for (int k = 0; k < 15; ++k) A[k] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) F[i] = g(i);
    const int ns = (NS_T > 0) ? NS_T : P.nsys;
#pragma unroll
    for (int q = 0; q < ns; ++q) {
      double tau = 0.0;
#pragma unroll
      for (int c = 0; c < 5; ++c) tau += P.m[q][c] * s[c];
      double gd, dgd;
      slip_rate_t<NPOW_T>(P, q, tau, itc(q), gd, dgd);
      const double a = dt * gd, bcoef = dt * dgd;
#pragma unroll
      for (int c = 0; c < 5; ++c) F[c] -= a * P.m[q][c];
#pragma unroll
      for (int k = 0; k < 15; ++k) A[k] += bcoef * P.mm[q][k];
    }
    double J[21];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) {
        const double jb = Jb(sidx(i, j));
        J[sidx(i, j)] = (i < 5 && j < 5) ? jb + A[s5idx(i, j)] : jb;
        F[i] -= jb * s[j];
        if (j != i) F[j] -= jb * s[i];
      }
    const bool ok = ldl6_solve(J, F);
    double dn = 0.0, sn = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      s[i] += F[i];
      dn += F[i] * F[i];
      sn += s[i] * s[i];
    }
    ++it;
    if (!ok || !(dn == dn) || !(sn == sn) || dn > 1e300 || sn > 1e300) { *bad = 1; conv = true; break; }
    if (dn <= tol * tol * sn) { conv = true; break; }
  }
  if (!conv) *bad |= 2;   // itmax exhausted without meeting tol: the multiplier identity (a5) does not hold for this voxel
  return it;
}

// kernel-constant parameters of the constitutive kernel
struct ConstParams {
  double S0b[21];           // reference compliance, b-basis, packed (sample frame)
  double dt, tol_newton;
  int newton_itmax, iso_c0;
  // uniform-exponent fast path (phase 0): dt * gamma0_s * n per system, and 1/n
  double dtg0n[EVP_MAX_SYS];
  double inv_n;
};

EVP_HD void fill_uniform_rate(const PhaseDev &P0, ConstParams &cp) {
  for (int q = 0; q < EVP_MAX_SYS; ++q) cp.dtg0n[q] = (q < P0.nsys) ? cp.dt * P0.g0[q] * P0.nrate[q] : 0.0;
  cp.inv_n = (P0.nsys > 0 && P0.nrate[0] > 0.0) ? 1.0 / P0.nrate[0] : 1.0;
}

// kn = dt*gamma0*n / tau_c^n for the uniform-exponent fast path (npow = n-1 integer)
EVP_HD double rate_factor(double dtg0n, double crss, int npow) {
  const double itc = 1.0 / crss;
  return dtg0n * pow_nm1(itc, npow, 0.0) * itc;
}

// reciprocal of a pivot: hardware seed (>= 20 bits) + two Newton steps on the device (no slow-path call),
// plain division on the host.  Pivots are compliances of order 1/modulus: never subnormal.
EVP_HD double rcp_pivot(double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / d;
#endif
}

// LDL^T solve as ldl6_solve, with rcp_pivot for the six pivot reciprocals
EVP_HD bool ldl6_solve_fast(double a[21], double b[6]) {
  double invd[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double w[6];
    double d = a[sidx(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) {
      w[k] = a[sidx(k, j)] * a[sidx(k, k)];
      d -= a[sidx(k, j)] * w[k];
    }
    a[sidx(j, j)] = d;
    ok = ok && (d > 0.0);
    const double inv = rcp_pivot(d);
    invd[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = a[sidx(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= a[sidx(k, i)] * w[k];
      a[sidx(j, i)] = t * inv;
    }
  }
#pragma unroll
  for (int i = 1; i < 6; ++i) {
#pragma unroll
    for (int k = 0; k < i; ++k) b[i] -= a[sidx(k, i)] * b[k];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] *= invd[i];
#pragma unroll
  for (int i = 4; i >= 0; --i) {
#pragma unroll
    for (int k = i + 1; k < 6; ++k) b[i] -= a[sidx(i, k)] * b[k];
  }
  return ok;
}

// LDL^T solve of a packed symmetric 5x5 (s5idx), pivot reciprocals by rcp_pivot
EVP_HD bool ldl5_solve_fast(double a[15], double b[5]) {
  double invd[5];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    double w[5];
    double d = a[s5idx(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) {
      w[k] = a[s5idx(k, j)] * a[s5idx(k, k)];
      d -= a[s5idx(k, j)] * w[k];
    }
    a[s5idx(j, j)] = d;
    ok = ok && (d > 0.0);
    const double inv = rcp_pivot(d);
    invd[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 5; ++i) {
      double t = a[s5idx(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= a[s5idx(k, i)] * w[k];
      a[s5idx(j, i)] = t * inv;
    }
  }
#pragma unroll
  for (int i = 1; i < 5; ++i) {
#pragma unroll
    for (int k = 0; k < i; ++k) b[i] -= a[s5idx(k, i)] * b[k];
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) b[i] *= invd[i];
#pragma unroll
  for (int i = 3; i >= 0; --i) {
#pragma unroll
    for (int k = i + 1; k < 5; ++k) b[i] -= a[s5idx(i, k)] * b[k];
  }
  return ok;
}

// Fast path: the plastic tangent lives in the deviatoric 5x5 block only, so the hydrostatic unknown s6 is eliminated once per
// orientation class (block elimination is exact: the Newton iterates are those of the 6x6 solve):
//   Jb = [K b; b^T d]   ->   table [K' = K - b b^T / d | b | d],     s6 = (g6 - b.s5)/d,     K' s5 + (1/n) A s5 = g5 - b g6 / d.
// In place on the packed 6x6: entries (i,j<5) become K', (i,5) keep b, (5,5) keeps d.
EVP_HD void jb_eliminate_hydrostatic(double Jb[21]) {
  const double invd = 1.0 / Jb[sidx(5, 5)];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = i; j < 5; ++j) Jb[sidx(i, j)] -= Jb[sidx(i, 5)] * Jb[sidx(j, 5)] * invd;
}

// |x|^K for NQ values at once, level by level (independent chains side by side: instruction-level parallelism)
template <int K, int NQ>
EVP_HD void pow_ct_arr(const double (&x)[NQ], double (&r)[NQ]) {
  if constexpr (K == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) r[q] = 1.0;
  } else if constexpr (K == 1) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) r[q] = fabs(x[q]);
  } else {
    pow_ct_arr<K / 2, NQ>(x, r);
#pragma unroll
    for (int q = 0; q < NQ; ++q) r[q] *= r[q];
    if constexpr (K & 1) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) r[q] *= fabs(x[q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FCC {111}<110> specialisation.  The twelve Schmid tensors in the b-basis are universal constants (no lattice
// parameter enters): entries 0, +-1/(2 sqrt3), +-1/2, +-1/sqrt3, 16 of the 60 components and 76 of the 180 packed
// products vanish.  With the table known at compile time the zero terms are never issued and the others are literal
// operands.  Order and signs are those of evp_phase_fcc (host_tables.cpp); the solver only selects this path after
// comparing the uploaded phase table with fcc_m() entry by entry (fcc_table_matches).
// ---------------------------------------------------------------------------------------------
EVP_HD constexpr double fcc_m(int q, int c) {
  constexpr double a = 0.28867513459481288225, h = 0.5, t = 0.57735026918962576451;
  constexpr double T[12][5] = {{a, -h, 0, -a, a},  {-a, -h, -a, 0, a}, {-t, 0, -a, a, 0}, {a, -h, 0, a, -a},
                               {a, h, a, 0, a},    {t, 0, a, a, 0},    {-a, h, 0, a, a},  {-a, -h, a, 0, -a},
                               {-t, 0, a, a, 0},   {a, -h, 0, a, a},   {-a, -h, a, 0, a}, {-t, 0, a, -a, 0}};
  return T[q][c];
}
EVP_HD constexpr int s5row(int k) { return k < 5 ? 0 : (k < 9 ? 1 : (k < 12 ? 2 : (k < 14 ? 3 : 4))); }
EVP_HD constexpr int s5col(int k) { return k - s5idx(s5row(k), s5row(k)) + s5row(k); }
EVP_HD constexpr double fcc_mm(int q, int k) { return fcc_m(q, s5row(k)) * fcc_m(q, s5col(k)); }
inline bool fcc_table_matches(const PhaseDev &P) {
  if (P.nsys != 12) return false;
  for (int q = 0; q < 12; ++q)
    for (int c = 0; c < 5; ++c)
      if (fabs(P.m[q][c] - fcc_m(q, c)) > 1e-14) return false;
  return true;
}
// HCP, 24 systems in the order of evp_phase_hcp (3 prismatic<a>, 3 basal<a>, 12 pyramidal<c+a>, 6 tensile twins): the
// VALUES depend on c/a, the ZERO PATTERN of the b-basis Schmid tensors does not (prismatic and basal slip touch two or
// three of the five deviatoric components).  28 of the 120 components and 114 of the 360 packed products vanish; the
// pattern is compiled in, the values stay run-time constants.  Selected after hcp24_pattern_matches().
EVP_HD constexpr bool hcp24_nz(int q, int c) {
  constexpr int M[24] = {0x11, 0x10, 0x11, 0x08, 0x0c, 0x0c, 0x1e, 0x1f, 0x1f, 0x1e, 0x1f, 0x1f,
                         0x1f, 0x1f, 0x1e, 0x1f, 0x1f, 0x1e, 0x1f, 0x07, 0x1f, 0x1f, 0x07, 0x1f};   // bit c: component c may be non-zero
  return ((M[q] >> c) & 1) != 0;
}
inline bool hcp24_pattern_matches(const PhaseDev &P) {
  if (P.nsys != 24) return false;
  for (int q = 0; q < 24; ++q)
    for (int c = 0; c < 5; ++c)
      if (!hcp24_nz(q, c) && fabs(P.m[q][c]) > 1e-14) return false;
  return true;
}

// Structured Schmid tables.  TAB = 0: run-time tables, every term issued;  1: FCC literals;  2: HCP-24 zero pattern with
// run-time values.
template <int TAB>
EVP_HD constexpr bool tab_nz(int q, int c) { return TAB == 1 ? fcc_m(q, c) != 0.0 : (TAB == 2 ? hcp24_nz(q, c) : true); }
template <int TAB, int Q, int C>
EVP_HD double tab_m(const PhaseDev &P) {
  if constexpr (TAB == 1) { constexpr double v = fcc_m(Q, C); return v; }
  else return P.m[Q][C];
}
template <int TAB, int Q, int K>
EVP_HD double tab_mm(const PhaseDev &P) {
  if constexpr (TAB == 1) { constexpr double v = fcc_mm(Q, K); return v; }
  else return P.mm[Q][K];
}
// tau_q = m_q . s without the zero components
template <int TAB, int Q, int C, bool STARTED>
EVP_HD double tab_tau(const PhaseDev &P, const double *s, double acc) {
  if constexpr (C == 5) {
    return acc;
  } else {
    if constexpr (!tab_nz<TAB>(Q, C)) return tab_tau<TAB, Q, C + 1, STARTED>(P, s, acc);
    else if constexpr (!STARTED) return tab_tau<TAB, Q, C + 1, true>(P, s, tab_m<TAB, Q, C>(P) * s[C]);
    else return tab_tau<TAB, Q, C + 1, true>(P, s, acc + tab_m<TAB, Q, C>(P) * s[C]);
  }
}
// A += w * (m_q (x) m_q) without the zero products
template <int TAB, int Q, int K>
EVP_HD void tab_tangent(const PhaseDev &P, double *A, double w) {
  if constexpr (K < 15) {
    if constexpr (tab_nz<TAB>(Q, s5row(K)) && tab_nz<TAB>(Q, s5col(K))) A[K] += w * tab_mm<TAB, Q, K>(P);
    tab_tangent<TAB, Q, K + 1>(P, A, w);
  }
}
template <int TAB, int Q0, int... Qs>
EVP_HD void tab_all_tau(const PhaseDev &P, const double *s, double *tau, std::integer_sequence<int, Qs...>) {
  ((tau[Qs] = tab_tau<TAB, Q0 + Qs, 0, false>(P, s, 0.0)), ...);
}
template <int TAB, int Q0, int... Qs>
EVP_HD void tab_all_tangent(const PhaseDev &P, double *A, const double *w, std::integer_sequence<int, Qs...>) {
  (tab_tangent<TAB, Q0 + Qs, 0>(P, A, w[Qs]), ...);
}
// systems Q0 .. NS_T-1 of a structured table, G at a time: projection, power, tangent coefficient, tangent
template <int TAB, int NS_T, int NPOW_T, bool TWIN, int G, int Q0, class KN>
EVP_HD void tab_groups(const PhaseDev &P, const double *s, KN kn, double *A);

// Row a4, uniform-exponent fast path (every system of the phase has the same integer n = NPOW_T + 1, NS_T systems,
// NS_T a multiple of G).  Same Newton iteration and stop rule as newton_crystal_t; what differs is the arithmetic
// organisation:
//  * systems are processed G at a time, phase by phase (projection, power, tangent), so G independent dependency
//    chains are in flight (the DFMA pipe is half rate with a long dependent-issue latency);
//  * the plastic strain-rate term of the residual is taken from the tangent: edp is homogeneous of degree n in s, so
//    dt*edp(s) = (1/n) A s with A = dt * d(edp)/ds (Euler; holds with the one-sided twin cut-off too) — 25 FMAs instead
//    of 5*NS_T + NS_T;
//  * the per-voxel, per-system factor kn = dt*gamma0*n / tau_c^n is prepared once per increment (rate_factor above,
//    k_prep_itc), so the tangent coefficient is one multiply after the power: no tau/tau_c ratio in the loop.
//    Range: |tau|^(n-1) and tau_c^-n are formed separately; with n <= 20 and stresses below 1e12 in any unit system
//    both stay far inside the fp64 range.
template <int TAB, int NS_T, int NPOW_T, bool TWIN, int G, int Q0, class KN>
EVP_HD void tab_groups(const PhaseDev &P, const double *s, KN kn, double *A) {
  if constexpr (Q0 < NS_T) {
    double tau[G], w[G];
    tab_all_tau<TAB, Q0>(P, s, tau, std::make_integer_sequence<int, G>{});
    pow_ct_arr<NPOW_T, G>(tau, w);   // |tau|^(n-1)
#pragma unroll
    for (int q = 0; q < G; ++q) {
      double t = w[q] * kn(Q0 + q);
      if (TWIN) t = (P.twin[Q0 + q] != 0 && tau[q] <= 0.0) ? 0.0 : t;
      w[q] = t;
    }
    tab_all_tangent<TAB, Q0>(P, A, w, std::make_integer_sequence<int, G>{});
    tab_groups<TAB, NS_T, NPOW_T, TWIN, G, Q0 + G>(P, s, kn, A);
  }
}

template <int NS_T, int NPOW_T, bool TWIN, int G, int TAB, class JB, class GV, class KN>
EVP_HD int newton_crystal_p(const PhaseDev &P, const ConstParams &cp, JB Jb, GV g, double s[6], KN kn, int *bad) {
  static_assert(NS_T > 0 && NS_T % G == 0 && NPOW_T >= 0, "uniform fast path");
  static_assert(TAB != 1 || (NS_T == 12 && !TWIN), "FCC table: 12 systems, no twins");
  static_assert(TAB != 2 || NS_T == 24, "HCP pattern: 24 systems");
  const double tol = cp.tol_newton;
  const int itmax = cp.newton_itmax;
  // Jb is the eliminated table [K' | b | d]: g5 <- g5 - b g6/d once per voxel
  const double invd = rcp_pivot(Jb(sidx(5, 5)));   // d = S0_c66 + S_c66 > 0: a compliance, never subnormal
  const double g6d = g(5) * invd;
#pragma unroll
  for (int i = 0; i < 5; ++i) g(i, g(i) - Jb(sidx(i, 5)) * g6d);
  int it = 0;
  bool conv = false;
  while (it < itmax) {
    double A[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) A[k] = 0.0;
    if constexpr (TAB != 0) {
      tab_groups<TAB, NS_T, NPOW_T, TWIN, G, 0>(P, s, kn, A);
    } else {
#pragma unroll
      for (int q0 = 0; q0 < NS_T; q0 += G) {
        double tau[G], w[G];
#pragma unroll
        for (int q = 0; q < G; ++q) tau[q] = P.m[q0 + q][0] * s[0];
#pragma unroll
        for (int c = 1; c < 5; ++c)
#pragma unroll
          for (int q = 0; q < G; ++q) tau[q] += P.m[q0 + q][c] * s[c];
        pow_ct_arr<NPOW_T, G>(tau, w);   // |tau|^(n-1)
#pragma unroll
        for (int q = 0; q < G; ++q) {
          double t = w[q] * kn(q0 + q);   // dt * d(gamma_dot)/d(tau) = [dt gamma0 n / tau_c^n] |tau|^(n-1)
          if (TWIN) t = (P.twin[q0 + q] != 0 && tau[q] <= 0.0) ? 0.0 : t;
          w[q] = t;
        }
#pragma unroll
        for (int q = 0; q < G; ++q)
#pragma unroll
          for (int k = 0; k < 15; ++k) A[k] += w[q] * P.mm[q0 + q][k];
      }
    }
    // 5x5 system of the deviatoric unknowns (hydrostatic one eliminated, jb_eliminate_hydrostatic):
    //   F = g' - K' s5 - (1/n) A s5 ;  J = K' + A ;  g' was written over g(0..4) before the loop
    double J[15], F[5], As[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { F[i] = g(i); As[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = i; j < 5; ++j) {
        const double kp = Jb(sidx(i, j));
        const double a = A[s5idx(i, j)];
        J[s5idx(i, j)] = kp + a;
        As[i] += a * s[j];
        F[i] -= kp * s[j];
        if (j != i) { As[j] += a * s[i]; F[j] -= kp * s[i]; }
      }
#pragma unroll
    for (int i = 0; i < 5; ++i) F[i] -= cp.inv_n * As[i];
    const bool ok = ldl5_solve_fast(J, F);
    double dn = 0.0, sn = 0.0, bs = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      s[i] += F[i];
      dn += F[i] * F[i];
      sn += s[i] * s[i];
      bs += Jb(sidx(i, 5)) * s[i];
    }
    {
      const double s6 = g6d - bs * invd;     // the linear hydrostatic equation holds exactly after every update
      const double d6 = s6 - s[5];
      s[5] = s6;
      dn += d6 * d6;
      sn += s6 * s6;
    }
    ++it;
    if (!ok || !(dn == dn) || !(sn == sn) || dn > 1e300 || sn > 1e300) { *bad = 1; conv = true; break; }
    if (dn <= tol * tol * sn) { conv = true; break; }
  }
  if (!conv) *bad |= 2;   // itmax exhausted without meeting tol: the multiplier identity (a5) does not hold for this voxel
  return it;
}

// rotate the packed reference compliance into the crystal frame: S0c = Q^T S0b Q, Q = diag(M, 1)
EVP_HD void rotate_s0(const double *S0b, const double M[25], double out[21]) {
#pragma unroll
  for (int be = 0; be < 6; ++be) {
    double T[6];  // T[ga] = sum_de S0b[ga][de] Q[de][be]
#pragma unroll
    for (int ga = 0; ga < 6; ++ga) {
      if (be == 5) {
        T[ga] = S0b[sidx(ga, 5)];
      } else {
        double acc = 0.0;
#pragma unroll
        for (int de = 0; de < 5; ++de) acc += S0b[sidx(ga, de)] * M[de * 5 + be];
        T[ga] = acc;
      }
    }
#pragma unroll
    for (int al = 0; al <= be; ++al) {
      if (al == 5) {
        out[sidx(5, 5)] = T[5];
      } else {
        double acc = 0.0;
#pragma unroll
        for (int ga = 0; ga < 5; ++ga) acc += M[ga * 5 + al] * T[ga];
        out[sidx(al, be)] = acc;
      }
    }
  }
}

// Rows a4+a5+a6 for one voxel, split in three stages.
//  increment_invariants : per voxel, once per increment — the deviatoric rotation M (25) and the
//                         packed Jb = S0_c + S_c (21); both depend only on the lattice orientation,
//                         which is constant within an increment.
//  constitutive_prep    : g = M^T (S0:sig_old + e - eps_p), initial guess, s_old (stored through functors)
//  constitutive_finish  : norms + rotation of the converged crystal-frame stress back to the sample frame
EVP_HD void increment_invariants(const PhaseDev &P, const ConstParams &cp, const double R[9], double M[25], double Jb[21]) {
  rot_b5(R, M);
  if (cp.iso_c0) {
#pragma unroll
    for (int k = 0; k < 21; ++k) Jb[k] = cp.S0b[k] + P.Sc[k];
  } else {
    rotate_s0(cp.S0b, M, Jb);
#pragma unroll
    for (int k = 0; k < 21; ++k) Jb[k] += P.Sc[k];
  }
}

template <class MV, class STG, class STS>
EVP_HD void constitutive_prep(const ConstParams &cp, MV M, const double sig[6], const double em[6], STG stG, STS stS, double sc[6]) {
  double so[6], eb[6];
  cart_to_b(sig, so);
  cart_to_b(em, eb);
  double g[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double acc = eb[i];
#pragma unroll
    for (int j = 0; j < 6; ++j) acc += cp.S0b[sidx(i, j)] * so[j];
    g[i] = acc;
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    double x = 0.0, y = 0.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const double m = M(b * 5 + a);
      x += m * g[b];
      y += m * so[b];
    }
    stG(a, x);
    sc[a] = y;
  }
  stG(5, g[5]);
  sc[5] = so[5];
#pragma unroll
  for (int a = 0; a < 6; ++a) stS(a, sc[a]);
}

template <class MV, class JB, class SV>
EVP_HD void constitutive_finish(const PhaseDev &P, MV M, const double sc[6], JB Jb, SV sold, double sig[6], double *ds, double *de) {
  double d[6], ds2 = 0.0, de2 = 0.0, acc[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    d[a] = sc[a] - sold(a);
    ds2 += d[a] * d[a];
    acc[a] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) {
      const double w = Jb(sidx(i, j)) - P.Sc[sidx(i, j)];
      acc[i] += w * d[j];
      if (j != i) acc[j] += w * d[i];
    }
#pragma unroll
  for (int a = 0; a < 6; ++a) de2 += acc[a] * acc[a];
  *ds = sqrt(ds2);
  *de = sqrt(de2);
  double sb[6];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    double x = 0.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) x += M(a * 5 + b) * sc[b];
    sb[a] = x;
  }
  sb[5] = sc[5];
  b_to_cart(sb, sig);
}

// constitutive_finish for the fast path: Jb is the eliminated table [K' | b | d] (jb_eliminate_hydrostatic), so the reference
// compliance in the crystal frame is  S0_c = [K' + b b^T/d - Sc55 | b - Sc_b ; . | d - Sc66].
template <class MV, class JB, class SV>
EVP_HD void constitutive_finish_p(const PhaseDev &P, MV M, const double sc[6], JB Jb, SV sold, double sig[6], double *ds, double *de) {
  double d[6], ds2 = 0.0, de2 = 0.0, acc[6], bd = 0.0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    d[a] = sc[a] - sold(a);
    ds2 += d[a] * d[a];
    acc[a] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) bd += Jb(sidx(i, 5)) * d[i];
  const double dd = Jb(sidx(5, 5));
  const double t = bd * rcp_pivot(dd) + d[5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = i; j < 5; ++j) {
      const double w = Jb(sidx(i, j)) - P.Sc[sidx(i, j)];
      acc[i] += w * d[j];
      if (j != i) acc[j] += w * d[i];
    }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double b = Jb(sidx(i, 5));
    acc[i] += b * t - P.Sc[sidx(i, 5)] * d[5];
    acc[5] += (b - P.Sc[sidx(i, 5)]) * d[i];
  }
  acc[5] += (dd - P.Sc[sidx(5, 5)]) * d[5];
#pragma unroll
  for (int a = 0; a < 6; ++a) de2 += acc[a] * acc[a];
  *ds = sqrt(ds2);
  *de = sqrt(de2);
  double sb[6];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    double x = 0.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) x += M(a * 5 + b) * sc[b];
    sb[a] = x;
  }
  sb[5] = sc[5];
  b_to_cart(sb, sig);
}

// ---------------------------------------------------------------------------------------------
// Row a2: Green operator at one frequency (vector form):
//   A_ik = C0_ijkl xi_j xi_l, G = A^-1, t = lam.xi, u = G t, de_ij = (u_i xi_j + u_j xi_i)/2
// KA[ik(6)][6]: host-precomputed coefficients so that A_ik = sum_m KA[ik][m] * p_m with
// p = (xx, yy, zz, yz, xz, xy) products of xi.  SC[36]: S0 acting on Cartesian 6-vectors
// (Nyquist planes).  `scale` = 1/(nx ny nz) folds the inverse-FFT normalisation in.
// ---------------------------------------------------------------------------------------------
struct GreenConst {
  double KA[6][6];
  double SC[36];
};

// scaled inverse acoustic tensor g = scale * (C0 : xi xi)^-1, symmetric, order 00,01,02,11,12,22
EVP_HD void green_G(const GreenConst &G0, double x, double y, double z, double scale, double g[6]) {
  const double p[6] = {x * x, y * y, z * z, y * z, x * z, x * y};
  double A[6];  // 00,11,22,12,02,01
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 6; ++m) acc += G0.KA[q][m] * p[m];
    A[q] = acc;
  }
  // inverse of the symmetric 3x3 [A0 A5 A4; A5 A1 A3; A4 A3 A2]
  const double c00 = A[1] * A[2] - A[3] * A[3];
  const double c01 = A[4] * A[3] - A[5] * A[2];
  const double c02 = A[5] * A[3] - A[4] * A[1];
  const double c11 = A[0] * A[2] - A[4] * A[4];
  const double c12 = A[5] * A[4] - A[0] * A[3];
  const double c22 = A[0] * A[1] - A[5] * A[5];
  const double det = A[0] * c00 + A[5] * c01 + A[4] * c02;
  const double id = scale / det;
  g[0] = c00 * id; g[1] = c01 * id; g[2] = c02 * id; g[3] = c11 * id; g[4] = c12 * id; g[5] = c22 * id;
}
// de = sym(u (x) xi), u = g (lam . xi): real-linear, applied to the real and imaginary parts in turn.
// lam / out in Cartesian order 11,22,33,23,13,12.
EVP_HD void green_apply(const double g[6], double x, double y, double z, const double lam[6], double out[6]) {
  const double t0 = lam[0] * x + lam[5] * y + lam[4] * z;
  const double t1 = lam[5] * x + lam[1] * y + lam[3] * z;
  const double t2 = lam[4] * x + lam[3] * y + lam[2] * z;
  const double u0 = g[0] * t0 + g[1] * t1 + g[2] * t2;
  const double u1 = g[1] * t0 + g[3] * t1 + g[4] * t2;
  const double u2 = g[2] * t0 + g[4] * t1 + g[5] * t2;
  out[0] = u0 * x;
  out[1] = u1 * y;
  out[2] = u2 * z;
  out[3] = 0.5 * (u1 * z + u2 * y);
  out[4] = 0.5 * (u0 * z + u2 * x);
  out[5] = 0.5 * (u0 * y + u1 * x);
}
// Nyquist planes: de = scale * S0 : lam  (SC acts on Cartesian 6-vectors)
EVP_HD void green_nyquist(const GreenConst &G0, double scale, const double lam[6], double out[6]) {
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) acc += G0.SC[6 * a + b] * lam[b];
    out[a] = acc * scale;
  }
}

EVP_HD void green_point(const GreenConst &G0, double x, double y, double z, bool zero, bool nyq, double scale,
                        const double2 lam[6], double2 out[6]) {
  if (zero) {
#pragma unroll
    for (int c = 0; c < 6; ++c) out[c] = make_double2(0.0, 0.0);
    return;
  }
  double lr[6], li[6], orr[6], oi[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { lr[c] = lam[c].x; li[c] = lam[c].y; }
  if (nyq) {
    green_nyquist(G0, scale, lr, orr);
    green_nyquist(G0, scale, li, oi);
  } else {
    double g[6];
    green_G(G0, x, y, z, scale, g);
    green_apply(g, x, y, z, lr, orr);
    green_apply(g, x, y, z, li, oi);
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) out[c] = make_double2(orr[c], oi[c]);
}


// Local rotation spectrum of a compatible strain spectrum (commit step, §8(f).1): axial (w32,w13,w21) of
// skew(u (x) xi) = (e^.xi (x) xi - xi (x) e^.xi)/|xi|^2, applied to the real or imaginary part. e in order 11,22,33,23,13,12.
EVP_HD void rot_apply(double x, double y, double z, double scale, const double e[6], double w[3]) {
  const double t0 = e[0] * x + e[5] * y + e[4] * z;
  const double t1 = e[5] * x + e[1] * y + e[3] * z;
  const double t2 = e[4] * x + e[3] * y + e[2] * z;
  const double s = scale / (x * x + y * y + z * z);
  w[0] = (t2 * y - t1 * z) * s;
  w[1] = (t0 * z - t2 * x) * s;
  w[2] = (t1 * x - t0 * y) * s;
}
// R <- exp([w]x) R, w axial (w32,w13,w21), R row major
EVP_HD void rotate_lattice(double R[9], const double w[3]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;
  if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
  else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double Q[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double k2 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) k2 += K[3 * i + k] * K[3 * k + j];
      Q[3 * i + j] = ((i == j) ? 1.0 : 0.0) + a * K[3 * i + j] + b * k2;
    }
  double Rn[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rn[3 * i + j] = Q[3 * i] * R[j] + Q[3 * i + 1] * R[3 + j] + Q[3 * i + 2] * R[6 + j];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = Rn[k];
}

// ---------------------------------------------------------------------------------------------
// Rows a1/a3: Stockham autosort radix passes.  A block owns L lines of length N in shared memory;
// thread (l, q), q in [0, N/8), owns the 8 points q + m*N/8 of line l in every pass.
// ---------------------------------------------------------------------------------------------
EVP_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
EVP_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
EVP_HD double2 cmul(double2 a, double2 w) { return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
EVP_HD double2 cmulc(double2 a, double2 w) { return make_double2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }  // a*conj(w)
template <bool INV>
EVP_HD double2 mul_mi(double2 a) {  // forward: a * (-i);  inverse: a * (+i)
  return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

template <bool INV>
EVP_HD void bfly2(double2 *v) {
  const double2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <bool INV>
EVP_HD void bfly4(double2 *v) {
  const double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
  const double2 t2 = cadd(v[1], v[3]), t3 = mul_mi<INV>(csub(v[1], v[3]));
  v[0] = cadd(t0, t2);
  v[2] = csub(t0, t2);
  v[1] = cadd(t1, t3);
  v[3] = csub(t1, t3);
}
template <bool INV>
EVP_HD void bfly8(double2 *v) {
  // radix-2 DIF stage, then two radix-4
  double2 a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = cadd(v[i], v[i + 4]);
    b[i] = csub(v[i], v[i + 4]);
  }
  // b[i] *= W8^i  (forward W8 = exp(-i pi/4))
  {
    const double2 t = b[1];
    b[1] = INV ? make_double2((t.x - t.y) * kRSQ2, (t.x + t.y) * kRSQ2) : make_double2((t.x + t.y) * kRSQ2, (t.y - t.x) * kRSQ2);
    b[2] = mul_mi<INV>(b[2]);
    const double2 u = b[3];
    b[3] = INV ? make_double2((-u.x - u.y) * kRSQ2, (u.x - u.y) * kRSQ2) : make_double2((u.y - u.x) * kRSQ2, (-u.x - u.y) * kRSQ2);
  }
  bfly4<INV>(a);
  bfly4<INV>(b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = a[i];
    v[2 * i + 1] = b[i];
  }
}
template <int R, bool INV>
EVP_HD void bfly(double2 *v) {
  if (R == 2) bfly2<INV>(v);
  else if (R == 4) bfly4<INV>(v);
  else bfly8<INV>(v);
}

// first-pass radix for length N = 2^k >= 8 so that the remaining passes are all radix 8
EVP_HD constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }
EVP_HD constexpr int first_radix(int n) { return (ilog2(n) % 3 == 0) ? 8 : (ilog2(n) % 3 == 1 ? 2 : 4); }

// load the thread's 8 points for a pass of radix R:  butterfly b handles j = q + b*N/8,
// inputs i = j + r*N/R.  OFF(i) maps a line index to the shared-memory element offset.
template <int N, int R, class OFF>
EVP_HD void pass_load(const double2 *s, int q, double2 v[8], OFF off) {
#pragma unroll
  for (int b = 0; b < 8 / R; ++b) {
    const int j = q + b * (N / 8);
#pragma unroll
    for (int r = 0; r < R; ++r) v[b * R + r] = s[off(j + r * (N / R))];
  }
}
// twiddle + butterfly + autosort store.  tw[k] = exp(-2 pi i k / N).
template <int N, int R, int NS, bool INV, class OFF, class TW>
EVP_HD void pass_store(double2 *s, int q, double2 v[8], OFF off, TW tw) {
#pragma unroll
  for (int b = 0; b < 8 / R; ++b) {
    const int j = q + b * (N / 8);
    const int k = j % NS;
    if (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const double2 w = tw(r * k * (N / (NS * R)));
        v[b * R + r] = INV ? cmulc(v[b * R + r], w) : cmul(v[b * R + r], w);
      }
    }
    bfly<R, INV>(&v[b * R]);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) s[off(base + r * NS)] = v[b * R + r];
  }
}


// ---------------------------------------------------------------------------------------------
// radix-16 Stockham passes (16 points per thread) for the persistent z kernel: N = 16*16 or 8*16,
// two passes instead of three -> fewer shared-memory round trips (the z pass is shared-memory bound).
// ---------------------------------------------------------------------------------------------
// 16-point DFT as 4x4: in place, natural order out.  Forward W16 = exp(-i pi/8).
template <bool INV>
EVP_HD void bfly16(double2 *v) {
  // step 1: DFT4 over n2 for every n1 (elements n1 + 4 n2) -> Y[n1][k2] at position n1 + 4 k2
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    double2 t[4] = {v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]};
    bfly4<INV>(t);
    v[n1] = t[0]; v[n1 + 4] = t[1]; v[n1 + 8] = t[2]; v[n1 + 12] = t[3];
  }
  // step 2: twiddles W16^(n1*k2)
  constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
  auto mulw = [](double2 a, double wr, double wi) {   // a * (wr - i wi) forward, a * (wr + i wi) inverse
    return INV ? make_double2(a.x * wr - a.y * wi, a.x * wi + a.y * wr) : make_double2(a.x * wr + a.y * wi, a.y * wr - a.x * wi);
  };
  v[1 + 4] = mulw(v[1 + 4], c1, s1);                 // W16^1
  v[1 + 8] = mulw(v[1 + 8], kRSQ2, kRSQ2);           // W16^2
  v[1 + 12] = mulw(v[1 + 12], s1, c1);               // W16^3
  v[2 + 4] = mulw(v[2 + 4], kRSQ2, kRSQ2);           // W16^2
  v[2 + 8] = mul_mi<INV>(v[2 + 8]);                  // W16^4 = -i
  v[2 + 12] = mulw(v[2 + 12], -kRSQ2, kRSQ2);        // W16^6
  v[3 + 4] = mulw(v[3 + 4], s1, c1);                 // W16^3
  v[3 + 8] = mulw(v[3 + 8], -kRSQ2, kRSQ2);          // W16^6
  v[3 + 12] = mulw(v[3 + 12], -c1, -s1);             // W16^9 = -W16^1
  // step 3: DFT4 over n1 for every k2 -> X[4 k1 + k2] at position k1 + 4 k2
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) bfly4<INV>(&v[4 * k2]);
}

// thread (q in [0, N/16)) loads its 16 points for a pass of radix R (16 or 8): butterfly b handles j = q + b*N/16
template <int N, int R, class OFF>
EVP_HD void pass16_load(const double2 *s, int q, double2 v[16], OFF off) {
#pragma unroll
  for (int b = 0; b < 16 / R; ++b) {
    const int j = q + b * (N / 16);
#pragma unroll
    for (int r = 0; r < R; ++r) v[b * R + r] = s[off(j + r * (N / R))];
  }
}
// One Stockham pass of the 16-points-per-thread kernels: radix R in {2, 8, 16}, NS = product of the radices before it.
// Twiddles W_N^(r*k*N/(NS*R)) with k = j % NS come from the accessor tw(r, k): the hoisted per-thread array of the last
// pass (k = q there) or a table lookup for a middle pass.  SWZ: the outputs of a lane are stored in a permuted order
// for lanes whose output group (j - k)/NS is odd — without it lanes of one warp write 64-byte pieces that all start in
// the same half of the 128-byte bank row (two-way conflict on every store).
template <int N, int R, int NS, bool INV, bool SWZ, class OFF, class TW>
EVP_HD void pass16_store_t(double2 *s, int q, double2 v[16], OFF off, TW tw) {
#pragma unroll
  for (int b = 0; b < 16 / R; ++b) {
    const int j = q + b * (N / 16);
    const int k = j % NS;
    if (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const double2 w = tw(r, k);
        v[b * R + r] = INV ? cmulc(v[b * R + r], w) : cmul(v[b * R + r], w);
      }
    }
    if (R == 16) bfly16<INV>(&v[b * R]);
    else if (R == 8) bfly8<INV>(&v[b * R]);
    else bfly2<INV>(&v[b * R]);
    const int base = (j - k) * R + k;
    if (SWZ) {
      const bool odd = (((j - k) / NS) & 1) != 0;
      constexpr int SW = (R == 16) ? 4 : 1;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ra = r, rb = r ^ SW;
        const int xa = (R == 16) ? (4 * (ra & 3) + (ra >> 2)) : ra;   // radix 16 leaves X[4 k1 + k2] at position k1 + 4 k2
        const int xb = (R == 16) ? (4 * (rb & 3) + (rb >> 2)) : rb;
        const double2 va = v[b * R + ra], vb = v[b * R + rb];
        const double2 val = odd ? vb : va;
        s[off(base + (odd ? xb : xa) * NS)] = val;
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int xr = (R == 16) ? (4 * (r & 3) + (r >> 2)) : r;
        s[off(base + xr * NS)] = v[b * R + r];
      }
    }
  }
}
struct TwArr {   // hoisted twiddles of the last pass: tw[r] = W_N^(r*q)
  const double2 *t;
  EVP_HD double2 operator()(int r, int) const { return t[r]; }
};
// first pass (NS = 1, no twiddles) or last pass (NS = N/16, R = 16, hoisted twiddles tw[r] = W_N^(r*k), k = q)
template <int N, int R, int NS, bool INV, class OFF>
EVP_HD void pass16_store(double2 *s, int q, double2 v[16], OFF off, const double2 *tw) {
  pass16_store_t<N, R, NS, INV, NS == 1>(s, q, v, off, TwArr{tw});
}

}  // namespace evp
