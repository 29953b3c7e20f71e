// tests/emu/emu.cpp — TEST INFRASTRUCTURE ONLY.  Runs the thread-level math of the CUDA kernels
// (lapx_b200/csrc/evp_core.h, the very same inline functions the __global__ kernels call) on the
// CPU, so that the arithmetic can be checked in the GPU-less build container.  It is not part of
// the product library and is not reachable from the product API.
#include <cstring>
#include <vector>

#include "../../lapx_b200/csrc/host_math.h"

using namespace evp;
using namespace evp::host;

namespace {

struct OffLin { int base; int operator()(int i) const { return base + i; } };
struct TwTab { const double2 *t; double2 operator()(int k) const { return t[k]; } };

template <int N, int R, int NS, bool INV>
void emu_pass(double2 *s, int nlines, const double2 *tw) {
  std::vector<double2> regs((size_t)nlines * (N / 8) * 8);
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 8; ++q) pass_load<N, R>(s, q, &regs[((size_t)l * (N / 8) + q) * 8], OffLin{l * N});
  // (the kernel has __syncthreads() here)
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 8; ++q) pass_store<N, R, NS, INV>(s, q, &regs[((size_t)l * (N / 8) + q) * 8], OffLin{l * N}, TwTab{tw});
}
template <int N, int NS, bool INV>
void emu_rest(double2 *s, int nlines, const double2 *tw) {
  if constexpr (NS < N) {
    emu_pass<N, 8, NS, INV>(s, nlines, tw);
    emu_rest<N, NS * 8, INV>(s, nlines, tw);
  }
}
template <int N, bool INV>
void emu_fft_n(double2 *s, int nlines, const double2 *tw) {
  constexpr int R0 = first_radix(N);
  emu_pass<N, R0, 1, INV>(s, nlines, tw);
  emu_rest<N, R0, INV>(s, nlines, tw);
}

struct ItcArr { const double *p; double operator()(int s) const { return p[s]; } };

}  // namespace

namespace {
struct ArrAcc {
  double *p;
  double operator()(int k) const { return p[k]; }
  void operator()(int k, double v) const { p[k] = v; }
};
template <int NS_T, int NPOW_T>
int run_t(const PhaseDev &P, const ConstParams &cp, const double *R, double *sig, const double *em, double *itc, double *ds, double *de,
          int *bad) {
  double M[25], jb[21], g[6], so[6], sc[6];
  increment_invariants(P, cp, R, M, jb);                       // k_prep_increment
  constitutive_prep(cp, ArrAcc{M}, sig, em, ArrAcc{g}, ArrAcc{so}, sc);  // k_constitutive_t
  const int nit = newton_crystal_t<NS_T, NPOW_T>(P, ArrAcc{jb}, ArrAcc{g}, sc, cp.dt, cp.tol_newton, cp.newton_itmax, ArrAcc{itc}, bad);
  constitutive_finish(P, ArrAcc{M}, sc, ArrAcc{jb}, ArrAcc{so}, sig, ds, de);
  return nit;
}
// uniform-exponent fast path (k_constitutive_p): phased system loop, residual from the tangent
template <int NS_T, int NPOW_T, bool TWIN, int G, int TAB = 0>
int run_p(const PhaseDev &P, ConstParams &cp, const double *R, double *sig, const double *em, double *itc, double *ds, double *de,
          int *bad) {
  double M[25], jb[21], g[6], so[6], sc[6];
  fill_uniform_rate(P, cp);
  double kn[EVP_MAX_SYS];
  for (int q = 0; q < NS_T; ++q) kn[q] = rate_factor(cp.dtg0n[q], 1.0 / itc[q], NPOW_T);   // k_prep_itc
  increment_invariants(P, cp, R, M, jb);
  jb_eliminate_hydrostatic(jb);                                 // k_prep_orient, fast path
  constitutive_prep(cp, ArrAcc{M}, sig, em, ArrAcc{g}, ArrAcc{so}, sc);
  const int nit = newton_crystal_p<NS_T, NPOW_T, TWIN, G, TAB>(P, cp, ArrAcc{jb}, ArrAcc{g}, sc, ArrAcc{kn}, bad);
  constitutive_finish_p(P, ArrAcc{M}, sc, ArrAcc{jb}, ArrAcc{so}, sig, ds, de);
  return nit;
}
}  // namespace


// two-pass radix-16 path of the persistent z kernel (N = 128: 8x16, N = 256: 16x16)
template <int N, bool INV>
void emu_fft16_n(double2 *s, int nlines, const double2 *twt) {
  constexpr int R1 = N / 16;
  std::vector<double2> regs((size_t)nlines * (N / 16) * 16);
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 16; ++q) pass16_load<N, R1>(s, q, &regs[((size_t)l * (N / 16) + q) * 16], OffLin{l * N});
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 16; ++q) pass16_store<N, R1, 1, INV>(s, q, &regs[((size_t)l * (N / 16) + q) * 16], OffLin{l * N}, nullptr);
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 16; ++q) pass16_load<N, 16>(s, q, &regs[((size_t)l * (N / 16) + q) * 16], OffLin{l * N});
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 16; ++q) {
      double2 tw[16];
      for (int r = 0; r < 16; ++r) tw[r] = twt[(r * q * (N / (R1 * 16))) % N];   // hoisted: k = q, NS = R1
      pass16_store<N, 16, R1, INV>(s, q, &regs[((size_t)l * (N / 16) + q) * 16], OffLin{l * N}, tw);
    }
}

// three-pass path of the persistent z kernel for N = 512: radix 2, 16, 16 (same 16 points per thread)
struct TwMid { const double2 *t; int mult, n; double2 operator()(int r, int k) const { return t[(mult * r * k) % n]; } };
template <int N, bool INV>
void emu_fft16x3_n(double2 *s, int nlines, const double2 *twt) {
  std::vector<double2> regs((size_t)nlines * (N / 16) * 16);
  auto R = [&](int l, int q) { return &regs[((size_t)l * (N / 16) + q) * 16]; };
  for (int l = 0; l < nlines; ++l) for (int q = 0; q < N / 16; ++q) pass16_load<N, 2>(s, q, R(l, q), OffLin{l * N});
  for (int l = 0; l < nlines; ++l) for (int q = 0; q < N / 16; ++q) pass16_store_t<N, 2, 1, INV, true>(s, q, R(l, q), OffLin{l * N}, TwArr{nullptr});
  for (int l = 0; l < nlines; ++l) for (int q = 0; q < N / 16; ++q) pass16_load<N, 16>(s, q, R(l, q), OffLin{l * N});
  for (int l = 0; l < nlines; ++l) for (int q = 0; q < N / 16; ++q) pass16_store_t<N, 16, 2, INV, true>(s, q, R(l, q), OffLin{l * N}, TwMid{twt, N / 32, N});
  for (int l = 0; l < nlines; ++l) for (int q = 0; q < N / 16; ++q) pass16_load<N, 16>(s, q, R(l, q), OffLin{l * N});
  for (int l = 0; l < nlines; ++l)
    for (int q = 0; q < N / 16; ++q) {
      double2 tw[16];
      for (int r = 0; r < 16; ++r) tw[r] = twt[(r * q) % N];   // NS = N/16: W_N^(r*k), k = q
      pass16_store_t<N, 16, N / 16, INV, false>(s, q, R(l, q), OffLin{l * N}, TwArr{tw});
    }
}

extern "C" {

int emu_fft16(int n, int inv, int nlines, double *data) {
  std::vector<double2> tw(n);
  for (int k = 0; k < n; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  double2 *s = reinterpret_cast<double2 *>(data);
  if (n == 128) { if (inv) emu_fft16_n<128, true>(s, nlines, tw.data()); else emu_fft16_n<128, false>(s, nlines, tw.data()); return 0; }
  if (n == 256) { if (inv) emu_fft16_n<256, true>(s, nlines, tw.data()); else emu_fft16_n<256, false>(s, nlines, tw.data()); return 0; }
  if (n == 512) { if (inv) emu_fft16x3_n<512, true>(s, nlines, tw.data()); else emu_fft16x3_n<512, false>(s, nlines, tw.data()); return 0; }
  return -1;
}

// in-place FFT of nlines contiguous lines of length n (interleaved re,im); same pass sequence as block_fft
int emu_fft(int n, int inv, int nlines, double *data) {
  std::vector<double2> tw(n);
  for (int k = 0; k < n; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  double2 *s = reinterpret_cast<double2 *>(data);
#define E_(N) case N: if (inv) emu_fft_n<N, true>(s, nlines, tw.data()); else emu_fft_n<N, false>(s, nlines, tw.data()); return 0;
  switch (n) { E_(8) E_(16) E_(32) E_(64) E_(128) E_(256) E_(512) E_(1024) default: return -1; }
#undef E_
}

// one voxel of the generic kernel variant; force_aniso exercises the rotated-S0 path on an isotropic medium
int emu_constitutive(const evp_phase *ph, const double *c0_voigt, const double *R, double *sig, const double *e, const double *epsp,
                     const double *crss, double dt, double tol, int itmax, double *ds, double *de, int *bad, int force_aniso) {
  PhaseDev P;
  build_phase_dev(*ph, P);
  double C0m[36], S0m[36];
  voigt_to_mandel(c0_voigt, C0m);
  inv6(C0m, S0m);
  ConstParams cp;
  build_s0b(S0m, cp.S0b, &cp.iso_c0);
  if (force_aniso) cp.iso_c0 = 0;
  cp.dt = dt; cp.tol_newton = tol; cp.newton_itmax = itmax;
  double itc[EVP_MAX_SYS], em[6];
  for (int s = 0; s < ph->nsys; ++s) itc[s] = 1.0 / crss[s];
  for (int c = 0; c < 6; ++c) em[c] = e[c] - epsp[c];
  *bad = 0;
  return run_t<0, -2>(P, cp, R, sig, em, itc, ds, de, bad);
}

// Green operator at one frequency (k_zfused inner stage).  lam/out: 6 complex (re,im interleaved)
int emu_green_point(const double *c0_voigt, double x, double y, double z, int zero, int nyq, double scale, const double *lam, double *out) {
  double C0m[36], S0m[36];
  voigt_to_mandel(c0_voigt, C0m);
  inv6(C0m, S0m);
  GreenConst G;
  build_green_const(C0m, S0m, G);
  double2 l[6], o[6];
  for (int c = 0; c < 6; ++c) l[c] = make_double2(lam[2 * c], lam[2 * c + 1]);
  green_point(G, x, y, z, zero != 0, nyq != 0, scale, l, o);
  for (int c = 0; c < 6; ++c) { out[2 * c] = o[c].x; out[2 * c + 1] = o[c].y; }
  return 0;
}

int emu_rot_b5(const double *R, double *M) { rot_b5(R, M); return 0; }

// production (templated) form of k_constitutive_t: variant 0 = generic, 11..15 = uniform-exponent fast path, 1 = <12,9>, 2 = <12,-2>, 3 = <24,9>, 4 = <12,19>, 5 = <24,19>, 6 = <24,-2>
int emu_constitutive_t(int variant, const evp_phase *ph, const double *c0_voigt, const double *R, double *sig, const double *e,
                       const double *epsp, const double *crss, double dt, double tol, int itmax, double *ds, double *de, int *bad) {
  PhaseDev P;
  build_phase_dev(*ph, P);
  double C0m[36], S0m[36];
  voigt_to_mandel(c0_voigt, C0m);
  inv6(C0m, S0m);
  ConstParams cp;
  build_s0b(S0m, cp.S0b, &cp.iso_c0);
  cp.dt = dt; cp.tol_newton = tol; cp.newton_itmax = itmax;
  double itc[EVP_MAX_SYS], em[6];
  for (int s = 0; s < ph->nsys; ++s) itc[s] = 1.0 / crss[s];
  for (int c = 0; c < 6; ++c) em[c] = e[c] - epsp[c];
  *bad = 0;
  switch (variant) {
    case 1: return run_t<12, 9>(P, cp, R, sig, em, itc, ds, de, bad);
    case 2: return run_t<12, -2>(P, cp, R, sig, em, itc, ds, de, bad);
    case 3: return run_t<24, 9>(P, cp, R, sig, em, itc, ds, de, bad);
    case 4: return run_t<12, 19>(P, cp, R, sig, em, itc, ds, de, bad);
    case 5: return run_t<24, 19>(P, cp, R, sig, em, itc, ds, de, bad);
    case 6: return run_t<24, -2>(P, cp, R, sig, em, itc, ds, de, bad);
    case 11: return run_p<12, 9, false, 6>(P, cp, R, sig, em, itc, ds, de, bad);
    case 12: return run_p<12, 9, true, 4>(P, cp, R, sig, em, itc, ds, de, bad);
    case 13: return run_p<24, 9, true, 6>(P, cp, R, sig, em, itc, ds, de, bad);
    case 14: return run_p<12, 19, false, 6>(P, cp, R, sig, em, itc, ds, de, bad);
    case 15: return run_p<24, 19, true, 6>(P, cp, R, sig, em, itc, ds, de, bad);
    case 20: return run_p<30, 9, true, 10>(P, cp, R, sig, em, itc, ds, de, bad);
    case 16: if (!fcc_table_matches(P)) return -1; return run_p<12, 9, false, 12, 1>(P, cp, R, sig, em, itc, ds, de, bad);
    case 17: if (!fcc_table_matches(P)) return -1; return run_p<12, 19, false, 12, 1>(P, cp, R, sig, em, itc, ds, de, bad);
    case 18: if (!hcp24_pattern_matches(P)) return -1; return run_p<24, 9, true, 12, 2>(P, cp, R, sig, em, itc, ds, de, bad);
    case 19: if (!hcp24_pattern_matches(P)) return -1; return run_p<24, 19, true, 12, 2>(P, cp, R, sig, em, itc, ds, de, bad);
    default: return run_t<0, -2>(P, cp, R, sig, em, itc, ds, de, bad);
  }
}

}  // extern "C"
