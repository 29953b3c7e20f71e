// evp_oracle.cpp — CPU ORACLE for the EVPFFT equilibrium loop.  TEST INFRASTRUCTURE ONLY.
//
//   *** PARITY UNPINNED ***  /root/reference contains only LICENSE (LICENSE:1 "BSD 3-Clause
//   License", LICENSE:3 "Copyright (c) 2023, Los Alamos National Laboratory"); there is no LApx
//   source, test, golden vector or fixture to follow or to pin against.  This file restates the
//   PUBLISHED algorithm named by BASELINE.json:5 ("Lebensohn-style (EVP)FFT fixed-point
//   iteration"), as tabulated in SURVEY.md §8(a) rows a1..a7:
//     - Lebensohn, Kanjarla, Eisenlohr (2012) IJP 32-33: EVPFFT augmented-Lagrangian iteration
//     - Moulinec & Suquet (1998): Green operator of the reference medium, Nyquist treatment
//     - Tome, Canova, Kocks et al. (1984): extended Voce hardening
//   Every parity claim made with this oracle reads "vs our CPU restatement", never "vs LApx".
//
//   Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
//   load this library.  The product path (lapx_b200/libevpfft_b200.so) never does.
//
// It implements the ABI of include/evpfft.h (single rank), deliberately with DIFFERENT internal
// choices than the CUDA path so that agreement means something:
//   oracle: Mandel basis, sample-frame Newton, full 6x6 Jacobian + pivoted Gauss elimination,
//           explicit 4th-order Green tensor, recursive mixed-radix FFT (any n = 2^a 3^b 5^c ...).
//   CUDA  : deviatoric/hydrostatic b-basis, crystal-frame Newton, LDL^T, vector form of
//           Gamma:lambda fused in the z pass, shared-memory Stockham power-of-two FFT.
//
// Build: make -C oracle    (g++ -O3 -march=native -fopenmp -shared -fPIC; the library is always built on the host that runs it)

#include "../include/evpfft.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using cplx = std::complex<double>;
constexpr double kPi = 3.14159265358979323846264338327950288;
const double kS2 = std::sqrt(2.0);

// Cartesian component order 11,22,33,23,13,12  (include/evpfft.h)
const int kI[6] = {0, 1, 2, 1, 0, 0};
const int kJ[6] = {0, 1, 2, 2, 2, 1};
const double kW[6] = {1.0, 1.0, 1.0, std::sqrt(2.0), std::sqrt(2.0), std::sqrt(2.0)};  // Mandel weights

std::string g_create_error;

// ------------------------------------------------------------------------------------------
// small dense helpers (Mandel 6-vectors / 6x6 matrices, row major)
// ------------------------------------------------------------------------------------------
inline void mat6_vec(const double *A, const double *x, double *y) {
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int j = 0; j < 6; ++j) s += A[6 * i + j] * x[j];
    y[i] = s;
  }
}

// Solve A x = b (n<=6), Gaussian elimination with partial pivoting. Returns false if singular.
bool gauss_solve(int n, double *A /* n*n, destroyed */, double *b /* in: rhs, out: x */) {
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = std::fabs(A[n * c + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[n * r + c]) > best) { best = std::fabs(A[n * r + c]); p = r; }
    if (!(best > 0.0)) return false;
    if (p != c) {
      for (int k = 0; k < n; ++k) std::swap(A[n * c + k], A[n * p + k]);
      std::swap(b[c], b[p]);
    }
    const double inv = 1.0 / A[n * c + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = A[n * r + c] * inv;
      if (f == 0.0) continue;
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

bool mat6_inverse(const double *A, double *Ainv) {
  for (int c = 0; c < 6; ++c) {
    double M[36], rhs[6] = {0, 0, 0, 0, 0, 0};
    std::memcpy(M, A, sizeof(M));
    rhs[c] = 1.0;
    if (!gauss_solve(6, M, rhs)) return false;
    for (int r = 0; r < 6; ++r) Ainv[6 * r + c] = rhs[r];
  }
  return true;
}

// Voigt (engineering) 6x6 -> Mandel 6x6:  C^M_ab = w_a w_b C^V_ab
void voigt_to_mandel(const double *cv, double *cm) {
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) cm[6 * a + b] = kW[a] * kW[b] * cv[6 * a + b];
}
void mandel_to_voigt(const double *cm, double *cv) {
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) cv[6 * a + b] = cm[6 * a + b] / (kW[a] * kW[b]);
}

// Mandel rotation matrix Q(R):  mandel(R A R^T) = Q mandel(A), built by projecting the rotated
// basis tensors (slow, obviously right; used per voxel only through rot_cache below).
void mandel_rotation(const double *R /*9 row major*/, double *Q /*36*/) {
  for (int mu = 0; mu < 6; ++mu) {
    double B[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const double v = (mu < 3) ? 1.0 : 1.0 / kS2;
    B[kI[mu]][kJ[mu]] = v;
    B[kJ[mu]][kI[mu]] = v;
    double T[3][3], RB[3][3];
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

// x^n for the power law.  Integer n <= 64: binary powering (the CUDA path uses the same
// multiplication tree); otherwise pow().
inline double pow_rate(double x, double n) {
  const int ni = (int)n;
  if ((double)ni == n && ni >= 0 && ni <= 64) {
    double r = 1.0, b = x;
    int k = ni;
    while (k) {
      if (k & 1) r *= b;
      b *= b;
      k >>= 1;
    }
    return r;
  }
  return std::pow(x, n);
}

// ------------------------------------------------------------------------------------------
// FFT: recursive mixed-radix decimation in time, any length (radix 4/2/3/5, generic primes)
// ------------------------------------------------------------------------------------------
struct FftPlan {
  int n = 0;
  std::vector<cplx> tw;  // tw[k] = exp(-2 pi i k / n)
  explicit FftPlan(int n_ = 0) : n(n_), tw((size_t)std::max(n_, 1)) {
    for (int k = 0; k < n; ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
      tw[k] = cplx((double)std::cos(a), (double)std::sin(a));
    }
  }
};

int pick_radix(int n) {
  if (n % 4 == 0) return 4;
  if (n % 2 == 0) return 2;
  if (n % 3 == 0) return 3;
  if (n % 5 == 0) return 5;
  for (int p = 7; p * p <= n; p += 2)
    if (n % p == 0) return p;
  return n;
}

// out[0..n) = DFT of in[0], in[stride], ...; twiddle W_n^k = plan.tw[k * (plan.n / n)] (conj if inverse)
void fft_rec(const FftPlan &plan, const cplx *in, cplx *out, int n, int stride, bool inverse) {
  if (n == 1) { out[0] = in[0]; return; }
  const int p = pick_radix(n);
  const int m = n / p;
  for (int r = 0; r < p; ++r) fft_rec(plan, in + (size_t)r * stride, out + (size_t)r * m, m, stride * p, inverse);
  const int tws = plan.n / n;
  auto W = [&](long idx) {  // W_n^idx
    const cplx w = plan.tw[(size_t)((idx % n) * tws)];
    return inverse ? std::conj(w) : w;
  };
  std::vector<cplx> tmp_dyn;
  cplx tmp_st[8];
  cplx *t = tmp_st;
  if (p > 8) { tmp_dyn.resize(p); t = tmp_dyn.data(); }
  for (int k = 0; k < m; ++k) {
    for (int r = 0; r < p; ++r) t[r] = out[(size_t)r * m + k] * W((long)r * k);
    if (p == 2) {
      out[k] = t[0] + t[1];
      out[k + m] = t[0] - t[1];
    } else if (p == 4) {
      const cplx a = t[0] + t[2], b = t[0] - t[2], c = t[1] + t[3], d = t[1] - t[3];
      const cplx jd = inverse ? cplx(-d.imag(), d.real()) : cplx(d.imag(), -d.real());  // -i*d (fwd)
      out[k] = a + c;
      out[k + m] = b + jd;
      out[k + 2 * m] = a - c;
      out[k + 3 * m] = b - jd;
    } else {
      for (int q = 0; q < p; ++q) {
        cplx s = t[0];
        for (int r = 1; r < p; ++r) s += t[r] * W((long)r * q * m);
        out[(size_t)q * m + k] = s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// solver state
// ------------------------------------------------------------------------------------------
struct PhaseData {
  evp_phase in;
  double Cm[36];                 // crystal stiffness, Mandel
  double Sm[36];                 // crystal compliance, Mandel
  double schmid[EVP_MAX_SYS][6]; // crystal-frame Schmid tensors, Mandel
  double alpha[EVP_MAX_SYS][3];  // skew part of b(x)n: (32,13,21) axial vector, crystal frame
};

}  // namespace

struct evp_solver {
  evp_grid g{};
  int nx = 0, ny = 0, nz = 0, nxh = 0;
  size_t N = 0;
  int nphases = 0, nsmax = 0;
  std::vector<PhaseData> ph;
  std::vector<int32_t> grain, phase;
  std::vector<double> rot, sig, e, epsp, edotp, crss, gacc, twinf, de, wrot;
  std::vector<int32_t> twinned;
  long long ntwinned = 0;
  double facc = 0.0;             // F_acc: twin volume fraction accumulated over the history (never decreases)
  long long last_unconverged = 0;
  double C0[36]{}, S0[36]{};
  bool have_micro = false, have_c0 = false, have_loading = false, in_incr = false;
  evp_ctrl ctrl{1e-6, 1e-6, 100, 1, 1e-6, 100, 0, 0};
  // loading
  int iudot[9]{}, iscau[6]{};
  double udot[9]{}, scau[6]{};
  bool strain_ctl[6]{};
  // macro state (Cartesian components)
  double Et[6]{}, E[6]{}, dEpend[6]{}, Edot_prev[6]{}, savg[6]{};
  double dt = 0;
  int iter = 0;
  double last_err_s = 0, last_err_e = 0;
  std::string err;
  FftPlan px, py, pz;
  std::vector<cplx> spec;  // 6 * nz*ny*nxh half spectra
};

namespace {

int fail(evp_handle h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

void build_phase(const evp_phase &in, PhaseData &pd) {
  pd.in = in;
  voigt_to_mandel(in.c_voigt, pd.Cm);
  mat6_inverse(pd.Cm, pd.Sm);
  for (int s = 0; s < in.nsys; ++s) {
    double b[3], n[3], bl = 0, nl = 0;
    for (int k = 0; k < 3; ++k) { bl += in.b[s][k] * in.b[s][k]; nl += in.n[s][k] * in.n[s][k]; }
    bl = std::sqrt(bl); nl = std::sqrt(nl);
    for (int k = 0; k < 3; ++k) { b[k] = in.b[s][k] / bl; n[k] = in.n[s][k] / nl; }
    for (int c = 0; c < 6; ++c)
      pd.schmid[s][c] = kW[c] * 0.5 * (b[kI[c]] * n[kJ[c]] + b[kJ[c]] * n[kI[c]]);
    // skew part q = (b n^T - n b^T)/2 ; axial storage (q32, q13, q21)
    pd.alpha[s][0] = 0.5 * (b[2] * n[1] - n[2] * b[1]);
    pd.alpha[s][1] = 0.5 * (b[0] * n[2] - n[0] * b[2]);
    pd.alpha[s][2] = 0.5 * (b[1] * n[0] - n[1] * b[0]);
  }
}

// ---- 3-D transforms ----------------------------------------------------------------------
// forward: real field f[z][y][x] pairs (a,b) -> half spectra A,B [z][y][kx], kx in [0,nx/2]
void fft3_forward_pair(evp_solver *S, const double *a, const double *b, cplx *A, cplx *B) {
  const int nx = S->nx, ny = S->ny, nz = S->nz, nxh = S->nxh;
#pragma omp parallel
  {
    std::vector<cplx> in(std::max({nx, ny, nz})), out(std::max({nx, ny, nz}));
#pragma omp for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z)
      for (int y = 0; y < ny; ++y) {
        const size_t r = ((size_t)z * ny + y);
        for (int x = 0; x < nx; ++x) in[x] = cplx(a[r * nx + x], b[r * nx + x]);
        fft_rec(S->px, in.data(), out.data(), nx, 1, false);
        for (int k = 0; k < nxh; ++k) {
          const cplx zk = out[k], zmk = std::conj(out[(nx - k) % nx]);
          A[r * nxh + k] = 0.5 * (zk + zmk);
          const cplx d = 0.5 * (zk - zmk);         // = i * B
          B[r * nxh + k] = cplx(d.imag(), -d.real());
        }
      }
    for (int pass = 0; pass < 2; ++pass) {
      cplx *F = pass ? B : A;
#pragma omp for collapse(2) schedule(static)
      for (int z = 0; z < nz; ++z)
        for (int k = 0; k < nxh; ++k) {
          for (int y = 0; y < ny; ++y) in[y] = F[((size_t)z * ny + y) * nxh + k];
          fft_rec(S->py, in.data(), out.data(), ny, 1, false);
          for (int y = 0; y < ny; ++y) F[((size_t)z * ny + y) * nxh + k] = out[y];
        }
#pragma omp for collapse(2) schedule(static)
      for (int y = 0; y < ny; ++y)
        for (int k = 0; k < nxh; ++k) {
          for (int z = 0; z < nz; ++z) in[z] = F[((size_t)z * ny + y) * nxh + k];
          fft_rec(S->pz, in.data(), out.data(), nz, 1, false);
          for (int z = 0; z < nz; ++z) F[((size_t)z * ny + y) * nxh + k] = out[z];
        }
    }
  }
}

// inverse: half spectra A,B -> real fields a,b (scaled by 1/N). A and B are destroyed.
void fft3_inverse_pair(evp_solver *S, cplx *A, cplx *B, double *a, double *b) {
  const int nx = S->nx, ny = S->ny, nz = S->nz, nxh = S->nxh;
  const double scale = 1.0 / ((double)nx * ny * nz);
#pragma omp parallel
  {
    std::vector<cplx> in(std::max({nx, ny, nz})), out(std::max({nx, ny, nz}));
    for (int pass = 0; pass < 2; ++pass) {
      cplx *F = pass ? B : A;
#pragma omp for collapse(2) schedule(static)
      for (int y = 0; y < ny; ++y)
        for (int k = 0; k < nxh; ++k) {
          for (int z = 0; z < nz; ++z) in[z] = F[((size_t)z * ny + y) * nxh + k];
          fft_rec(S->pz, in.data(), out.data(), nz, 1, true);
          for (int z = 0; z < nz; ++z) F[((size_t)z * ny + y) * nxh + k] = out[z];
        }
#pragma omp for collapse(2) schedule(static)
      for (int z = 0; z < nz; ++z)
        for (int k = 0; k < nxh; ++k) {
          for (int y = 0; y < ny; ++y) in[y] = F[((size_t)z * ny + y) * nxh + k];
          fft_rec(S->py, in.data(), out.data(), ny, 1, true);
          for (int y = 0; y < ny; ++y) F[((size_t)z * ny + y) * nxh + k] = out[y];
        }
    }
#pragma omp for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z)
      for (int y = 0; y < ny; ++y) {
        const size_t r = ((size_t)z * ny + y);
        // Z(k) = A(k) + i B(k); Hermitian completion for k > nx/2
        for (int k = 0; k < nx; ++k) {
          if (k < nxh) {
            const cplx Ak = A[r * nxh + k], Bk = B[r * nxh + k];
            in[k] = Ak + cplx(-Bk.imag(), Bk.real());
          } else {
            const cplx Ak = std::conj(A[r * nxh + (nx - k)]), Bk = std::conj(B[r * nxh + (nx - k)]);
            in[k] = Ak + cplx(-Bk.imag(), Bk.real());
          }
        }
        fft_rec(S->px, in.data(), out.data(), nx, 1, true);
        for (int x = 0; x < nx; ++x) {
          a[r * nx + x] = out[x].real() * scale;
          b[r * nx + x] = out[x].imag() * scale;
        }
      }
  }
}

inline int freq_index(int k, int n) { return (k <= n / 2) ? k : k - n; }

// Rows a1-a3: de = FFT^-1[ Gamma0^ : FFT(sig) ],  e <- e - de + dE_pending
void op_green(evp_solver *S) {
  const int nx = S->nx, ny = S->ny, nz = S->nz, nxh = S->nxh;
  const size_t N = S->N, NS = (size_t)nz * ny * nxh;
  S->spec.resize(6 * NS);
  cplx *sp = S->spec.data();
  for (int p = 0; p < 3; ++p)
    fft3_forward_pair(S, &S->sig[(2 * p) * N], &S->sig[(2 * p + 1) * N], sp + (2 * p) * NS, sp + (2 * p + 1) * NS);

  // C0 as a full 3x3x3x3 tensor
  double C4[3][3][3][3];
  {
    int vm[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) {
            const int a = vm[i][j], b = vm[k][l];
            C4[i][j][k][l] = S->C0[6 * a + b] / (kW[a] * kW[b]);
          }
  }
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int kx = 0; kx < nxh; ++kx) {
        const size_t idx = ((size_t)z * ny + y) * nxh + kx;
        const int fx = freq_index(kx, nx), fy = freq_index(y, ny), fz = freq_index(z, nz);
        cplx lam[6], out[6];
        for (int c = 0; c < 6; ++c) lam[c] = kW[c] * sp[c * NS + idx];  // Mandel
        if (fx == 0 && fy == 0 && fz == 0) {
          for (int c = 0; c < 6; ++c) out[c] = 0.0;
        } else if ((nx % 2 == 0 && kx == nx / 2) || (ny % 2 == 0 && y == ny / 2) || (nz % 2 == 0 && z == nz / 2)) {
          // Nyquist planes: Gamma^ := S0 (Moulinec & Suquet 1998; zero stress at the Nyquist frequency)
          for (int a = 0; a < 6; ++a) {
            cplx s = 0.0;
            for (int b = 0; b < 6; ++b) s += S->S0[6 * a + b] * lam[b];
            out[a] = s;
          }
        } else {
          const double xi[3] = {(double)fx / (nx * S->g.dx), (double)fy / (ny * S->g.dy), (double)fz / (nz * S->g.dz)};
          double A[9], G[9];
          for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) {
              double s = 0;
              for (int j = 0; j < 3; ++j)
                for (int l = 0; l < 3; ++l) s += C4[i][j][k][l] * xi[j] * xi[l];
              A[3 * i + k] = s;
            }
          // 3x3 inverse by cofactors
          const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
          const double id = 1.0 / det;
          G[0] = (A[4] * A[8] - A[5] * A[7]) * id; G[1] = (A[2] * A[7] - A[1] * A[8]) * id; G[2] = (A[1] * A[5] - A[2] * A[4]) * id;
          G[3] = (A[5] * A[6] - A[3] * A[8]) * id; G[4] = (A[0] * A[8] - A[2] * A[6]) * id; G[5] = (A[2] * A[3] - A[0] * A[5]) * id;
          G[6] = (A[3] * A[7] - A[4] * A[6]) * id; G[7] = (A[1] * A[6] - A[0] * A[7]) * id; G[8] = (A[0] * A[4] - A[1] * A[3]) * id;
          // Gamma_ijkl = 1/4 (G_ik xj xl + G_jk xi xl + G_il xj xk + G_jl xi xk), Mandel 6x6
          for (int a = 0; a < 6; ++a) {
            const int i = kI[a], j = kJ[a];
            cplx s = 0.0;
            for (int b = 0; b < 6; ++b) {
              const int k = kI[b], l = kJ[b];
              const double gam = 0.25 * (G[3 * i + k] * xi[j] * xi[l] + G[3 * j + k] * xi[i] * xi[l] +
                                         G[3 * i + l] * xi[j] * xi[k] + G[3 * j + l] * xi[i] * xi[k]);
              s += (kW[a] * kW[b] * gam) * lam[b];
            }
            out[a] = s;
          }
        }
        for (int c = 0; c < 6; ++c) sp[c * NS + idx] = out[c] / kW[c];  // back to Cartesian components
      }
  for (int p = 0; p < 3; ++p)
    fft3_inverse_pair(S, sp + (2 * p) * NS, sp + (2 * p + 1) * NS, &S->de[(2 * p) * N], &S->de[(2 * p + 1) * N]);
#pragma omp parallel for schedule(static)
  for (size_t v = 0; v < N; ++v)
    for (int c = 0; c < 6; ++c) S->e[c * N + v] += S->dEpend[c] - S->de[c * N + v];
  for (int c = 0; c < 6; ++c) S->dEpend[c] = 0.0;
}

// Local rotation fluctuation of the compatible strain field e: for e^ = sym(u (x) xi),
//   w^_ij = (e^_ik xi_k xi_j - e^_jk xi_k xi_i) / |xi|^2 = skew(u (x) xi);  zero at xi = 0 and on the Nyquist planes.
// Output: axial components (w32, w13, w21) per voxel.
void local_rotation(evp_solver *S, std::vector<double> &w) {
  const int nx = S->nx, ny = S->ny, nz = S->nz, nxh = S->nxh;
  const size_t N = S->N, NS = (size_t)nz * ny * nxh;
  std::vector<cplx> sp(6 * NS);
  for (int p = 0; p < 3; ++p)
    fft3_forward_pair(S, &S->e[(2 * p) * N], &S->e[(2 * p + 1) * N], sp.data() + (2 * p) * NS, sp.data() + (2 * p + 1) * NS);
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int kx = 0; kx < nxh; ++kx) {
        const size_t idx = ((size_t)z * ny + y) * nxh + kx;
        const int fx = freq_index(kx, nx), fy = freq_index(y, ny), fz = freq_index(z, nz);
        cplx out[3] = {0.0, 0.0, 0.0};
        const bool nyq = (nx % 2 == 0 && kx == nx / 2) || (ny % 2 == 0 && y == ny / 2) || (nz % 2 == 0 && z == nz / 2);
        if (!(fx == 0 && fy == 0 && fz == 0) && !nyq) {
          const double xi[3] = {(double)fx / (nx * S->g.dx), (double)fy / (ny * S->g.dy), (double)fz / (nz * S->g.dz)};
          const double x2 = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2];
          cplx E[3][3];
          for (int c = 0; c < 6; ++c) { E[kI[c]][kJ[c]] = sp[c * NS + idx]; E[kJ[c]][kI[c]] = sp[c * NS + idx]; }
          cplx t[3];
          for (int i = 0; i < 3; ++i) t[i] = E[i][0] * xi[0] + E[i][1] * xi[1] + E[i][2] * xi[2];
          out[0] = (t[2] * xi[1] - t[1] * xi[2]) / x2;   // w32
          out[1] = (t[0] * xi[2] - t[2] * xi[0]) / x2;   // w13
          out[2] = (t[1] * xi[0] - t[0] * xi[1]) / x2;   // w21
        }
        for (int c = 0; c < 3; ++c) sp[c * NS + idx] = out[c];
        sp[3 * NS + idx] = 0.0;
      }
  w.assign(4 * N, 0.0);
  fft3_inverse_pair(S, sp.data(), sp.data() + NS, &w[0], &w[N]);
  fft3_inverse_pair(S, sp.data() + 2 * NS, sp.data() + 3 * NS, &w[2 * N], &w[3 * N]);
  w.resize(3 * N);
}

// R <- exp([w]x) R  (Rodrigues), w = axial vector (w32, w13, w21)
void rotate_lattice(double *R, const double w[3]) {
  const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double a, b;   // exp = I + a K + b K^2
  if (th < 1e-8) { a = 1.0 - th * th / 6.0; b = 0.5 - th * th / 24.0; }
  else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / (th * th); }
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double K2[9], Q[9], Rn[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += K[3 * i + k] * K[3 * k + j];
      K2[3 * i + j] = s;
    }
  for (int k = 0; k < 9; ++k) Q[k] = ((k % 4 == 0) ? 1.0 : 0.0) + a * K[k] + b * K2[k];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Q[3 * i + k] * R[3 * k + j];
      Rn[3 * i + j] = s;
    }
  for (int k = 0; k < 9; ++k) R[k] = Rn[k];
}

// plastic strain rate (Mandel, sample frame) and its stress derivative at stress s6 (Mandel)
struct VoxelFrame {
  double Ssample[36];
  double msample[EVP_MAX_SYS][6];
};

void voxel_frame(const evp_solver *S, size_t v, VoxelFrame &F) {
  const PhaseData &pd = S->ph[S->phase[v]];
  double R[9], Q[36], T[36];
  for (int k = 0; k < 9; ++k) R[k] = S->rot[k * S->N + v];
  mandel_rotation(R, Q);
  // S_sample = Q S_c Q^T
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += Q[6 * i + k] * pd.Sm[6 * k + j];
      T[6 * i + j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += T[6 * i + k] * Q[6 * j + k];
      F.Ssample[6 * i + j] = s;
    }
  for (int s = 0; s < pd.in.nsys; ++s) mat6_vec(Q, pd.schmid[s], F.msample[s]);
}

// gamma_dot per system and d(gamma_dot)/d(tau) at resolved shear stress tau
inline void slip_rate(const evp_phase &p, int s, double tau, double tauc, double &gd, double &dgd) {
  const int m = p.mode[s];
  const double n = p.nrate[m], g0 = p.gamma0[m];
  if (p.twin[m] && tau <= 0.0) { gd = 0.0; dgd = 0.0; return; }
  const double x = std::fabs(tau) / tauc;
  const double xn1 = pow_rate(x, n - 1.0);
  gd = g0 * xn1 * x * (tau >= 0.0 ? 1.0 : -1.0);
  dgd = g0 * n * xn1 / tauc;
}

void plastic_rate(const evp_solver *S, size_t v, const VoxelFrame &F, const double *s6, double *edp6, double *dedp /*36 or null*/) {
  const PhaseData &pd = S->ph[S->phase[v]];
  for (int c = 0; c < 6; ++c) edp6[c] = 0.0;
  if (dedp) for (int c = 0; c < 36; ++c) dedp[c] = 0.0;
  for (int s = 0; s < pd.in.nsys; ++s) {
    const double *m = F.msample[s];
    double tau = 0;
    for (int c = 0; c < 6; ++c) tau += m[c] * s6[c];
    double gd, dgd;
    slip_rate(pd.in, s, tau, S->crss[(size_t)s * S->N + v], gd, dgd);
    for (int c = 0; c < 6; ++c) edp6[c] += gd * m[c];
    if (dedp)
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) dedp[6 * a + b] += dgd * m[a] * m[b];
  }
}

// Rows a4-a6 for all voxels.
void op_constitutive(evp_solver *S, evp_iter_report *rep) {
  const size_t N = S->N;
  double sum_ds = 0, sum_de = 0, ssum[6] = {0, 0, 0, 0, 0, 0};
  long nit_sum = 0, unconv = 0;
  int nit_max = 0, nonfinite = 0;
#pragma omp parallel for schedule(static) reduction(+ : sum_ds, sum_de, nit_sum, nonfinite, unconv, ssum[:6]) reduction(max : nit_max)
  for (size_t v = 0; v < N; ++v) {
    VoxelFrame F;
    voxel_frame(S, v, F);
    double so[6], e6[6], ep6[6], s6[6];
    for (int c = 0; c < 6; ++c) {
      so[c] = kW[c] * S->sig[c * N + v];
      e6[c] = kW[c] * S->e[c * N + v];
      ep6[c] = kW[c] * S->epsp[c * N + v];
      s6[c] = so[c];
    }
    int it = 0;
    bool done = false;
    for (; it < S->ctrl.newton_itmax;) {
      // F(s) = S0 (s - so) + Sx s + ep + dt edp(s) - e
      double edp[6], dedp[36], res[6], J[36], d[6], t1[6], t2[6];
      plastic_rate(S, v, F, s6, edp, dedp);
      for (int c = 0; c < 6; ++c) d[c] = s6[c] - so[c];
      mat6_vec(S->S0, d, t1);
      mat6_vec(F.Ssample, s6, t2);
      for (int c = 0; c < 6; ++c) res[c] = -(t1[c] + t2[c] + ep6[c] + S->dt * edp[c] - e6[c]);
      for (int k = 0; k < 36; ++k) J[k] = S->S0[k] + F.Ssample[k] + S->dt * dedp[k];
      if (!gauss_solve(6, J, res)) { nonfinite += 1; done = true; break; }
      double dn = 0, sn = 0;
      for (int c = 0; c < 6; ++c) { s6[c] += res[c]; dn += res[c] * res[c]; sn += s6[c] * s6[c]; }
      ++it;
      if (!(dn == dn) || !(sn == sn)) { nonfinite += 1; done = true; break; }
      if (std::sqrt(dn) <= S->ctrl.tol_newton * std::sqrt(sn)) { done = true; break; }
    }
    if (!done) unconv += 1;   // newton_itmax exhausted: sigma is not the converged multiplier for this voxel
    // error norms (published definitions): |sig_new - lambda_old| and |eps(sig_new) - e|
    double edp[6], t2[6], ds = 0, de = 0;
    plastic_rate(S, v, F, s6, edp, nullptr);
    mat6_vec(F.Ssample, s6, t2);
    for (int c = 0; c < 6; ++c) {
      const double a = s6[c] - so[c];
      const double b = t2[c] + ep6[c] + S->dt * edp[c] - e6[c];
      ds += a * a;
      de += b * b;
    }
    sum_ds += std::sqrt(ds);
    sum_de += std::sqrt(de);
    for (int c = 0; c < 6; ++c) {
      S->sig[c * N + v] = s6[c] / kW[c];
      S->edotp[c * N + v] = edp[c] / kW[c];
      ssum[c] += s6[c] / kW[c];
    }
    nit_sum += it;
    nit_max = std::max(nit_max, it);
  }
  double sn = 0, en = 0;
  for (int c = 0; c < 6; ++c) {
    S->savg[c] = ssum[c] / (double)N;
    const double w = (c < 3) ? 1.0 : 2.0;
    sn += w * S->savg[c] * S->savg[c];
    en += w * S->E[c] * S->E[c];
  }
  S->last_err_s = (sn > 0) ? (sum_ds / (double)N) / std::sqrt(sn) : (sum_ds / (double)N);
  S->last_err_e = (en > 0) ? (sum_de / (double)N) / std::sqrt(en) : (sum_de / (double)N);
  if (rep) {
    rep->newton_max = nit_max;
    rep->newton_mean = (double)nit_sum / (double)N;
    rep->nonfinite = nonfinite;
    rep->unconverged = unconv;
  }
  S->last_unconverged = unconv;
}

// Row a7: macro strain correction for stress-controlled components,
// dE_T = (C0_TT)^-1 (Sigma_T - <sig>_T), Mandel scaling.
void macro_update(evp_solver *S) {
  int idx[6], nt = 0;
  for (int c = 0; c < 6; ++c)
    if (!S->strain_ctl[c]) idx[nt++] = c;
  for (int c = 0; c < 6; ++c) S->dEpend[c] = 0.0;
  if (nt == 0) return;
  double A[36], r[6];
  for (int a = 0; a < nt; ++a) {
    r[a] = kW[idx[a]] * (S->scau[idx[a]] - S->savg[idx[a]]);
    for (int b = 0; b < nt; ++b) A[nt * a + b] = S->C0[6 * idx[a] + idx[b]];
  }
  gauss_solve(nt, A, r);
  for (int a = 0; a < nt; ++a) {
    S->dEpend[idx[a]] = r[a] / kW[idx[a]];
    S->E[idx[a]] += S->dEpend[idx[a]];
  }
}

void fill_report(evp_solver *S, evp_iter_report *rep) {
  if (!rep) return;
  rep->iter = S->iter;
  rep->err_stress = S->last_err_s;
  rep->err_strain = S->last_err_e;
  for (int c = 0; c < 6; ++c) { rep->savg[c] = S->savg[c]; rep->emacro[c] = S->E[c]; }
  rep->converged = (S->iter >= S->ctrl.itmin && S->last_err_s <= S->ctrl.tol_stress && S->last_err_e <= S->ctrl.tol_strain &&
                    S->last_unconverged == 0) ? 1 : 0;
}

double voce_tau(const evp_phase &p, int m, double G) {
  const double t0 = p.tau0[m], t1 = p.tau1[m], h0 = p.theta0[m], h1 = p.theta1[m];
  if (std::fabs(t1) < 1e-300) return t0 + h1 * G;
  return t0 + (t1 + h1 * G) * (1.0 - std::exp(-G * std::fabs(h0 / t1)));
}

size_t field_comps(const evp_solver *S, int f) {
  switch (f) {
    case EVP_FIELD_STRESS: case EVP_FIELD_STRAIN: case EVP_FIELD_PLASTIC_STRAIN:
    case EVP_FIELD_PLASTIC_RATE: case EVP_FIELD_STRAIN_INCR: return 6;
    case EVP_FIELD_CRSS: case EVP_FIELD_TWIN_FRACTION: return (size_t)S->nsmax;
    case EVP_FIELD_ROTATION: return 9;
    case EVP_FIELD_GRAIN: case EVP_FIELD_PHASE: case EVP_FIELD_GAMMA_ACC: case EVP_FIELD_TWINNED: return 1;
    case EVP_FIELD_LOCAL_ROTATION: return 3;
    default: return 0;
  }
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

int evp_abi_version(void) { return EVP_ABI_VERSION; }
const char *evp_backend(void) { return "cpu-oracle"; }

int evp_create(const evp_grid *grid, const evp_phase *phases, int32_t nphases, const evp_dist *dist, evp_handle *out) {
  if (!grid || !phases || !out || nphases < 1 || nphases > EVP_MAX_PHASES) return fail(nullptr, EVP_ERR_ARG, "evp_create: bad argument");
  if (dist && dist->nranks != 1) return fail(nullptr, EVP_ERR_UNSUPPORTED, "oracle is single rank");
  if (grid->nx < 2 || grid->ny < 2 || grid->nz < 2) return fail(nullptr, EVP_ERR_ARG, "grid must be >= 2 in every direction");
  evp_solver *S = new evp_solver();
  S->g = *grid;
  if (!(S->g.dx > 0)) S->g.dx = 1.0;
  if (!(S->g.dy > 0)) S->g.dy = 1.0;
  if (!(S->g.dz > 0)) S->g.dz = 1.0;
  S->nx = grid->nx; S->ny = grid->ny; S->nz = grid->nz; S->nxh = grid->nx / 2 + 1;
  S->N = (size_t)S->nx * S->ny * S->nz;
  S->nphases = nphases;
  S->ph.resize(nphases);
  for (int p = 0; p < nphases; ++p) {
    if (phases[p].nsys < 0 || phases[p].nsys > EVP_MAX_SYS || phases[p].nmodes < 0 || phases[p].nmodes > EVP_MAX_MODES) {
      delete S;
      return fail(nullptr, EVP_ERR_ARG, "phase: nsys/nmodes out of range");
    }
    build_phase(phases[p], S->ph[p]);
    S->nsmax = std::max(S->nsmax, (int)phases[p].nsys);
  }
  const size_t N = S->N;
  S->grain.assign(N, 0); S->phase.assign(N, 0);
  S->rot.assign(9 * N, 0.0); S->sig.assign(6 * N, 0.0); S->e.assign(6 * N, 0.0);
  S->epsp.assign(6 * N, 0.0); S->edotp.assign(6 * N, 0.0); S->de.assign(6 * N, 0.0);
  S->crss.assign((size_t)std::max(S->nsmax, 1) * N, 1.0); S->gacc.assign(N, 0.0);
  S->twinf.assign((size_t)std::max(S->nsmax, 1) * N, 0.0);
  S->wrot.assign(3 * N, 0.0); S->twinned.assign(N, 0);
  S->px = FftPlan(S->nx); S->py = FftPlan(S->ny); S->pz = FftPlan(S->nz);
  *out = S;
  return EVP_OK;
}

int evp_destroy(evp_handle h) { delete h; return EVP_OK; }

const char *evp_last_error(evp_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int evp_local_slab(evp_handle h, int32_t *z0, int32_t *nzl) {
  if (!h) return EVP_ERR_ARG;
  if (z0) *z0 = 0;
  if (nzl) *nzl = h->nz;
  return EVP_OK;
}
int evp_local_block(evp_handle h, int32_t *y0, int32_t *nyl, int32_t *z0, int32_t *nzl) {
  if (!h) return EVP_ERR_ARG;
  if (y0) *y0 = 0;
  if (nyl) *nyl = h->ny;
  if (z0) *z0 = 0;
  if (nzl) *nzl = h->nz;
  return EVP_OK;
}
int evp_nsys_max(evp_handle h) { return h ? h->nsmax : EVP_ERR_ARG; }
int evp_transport(evp_handle) { return EVP_TRANSPORT_NCCL; }
#ifndef EVP_ORACLE_BUILD
#define EVP_ORACLE_BUILD "EVPORACLE:manual"
#endif
const char *evp_build_id(void) { return EVP_ORACLE_BUILD; }
int64_t evp_launch_count(evp_handle) { return 0; }
int evp_debug_fp64_peak(evp_handle, int32_t, double *tflops) { if (tflops) *tflops = 0.0; return EVP_ERR_UNSUPPORTED; }

int evp_set_microstructure(evp_handle h, const int32_t *grain, const int32_t *phase, const double *rot9) {
  if (!h || !grain || !rot9) return fail(h, EVP_ERR_ARG, "set_microstructure: null pointer");
  const size_t N = h->N;
  for (size_t v = 0; v < N; ++v) {
    const int p = phase ? phase[v] : 0;
    if (p < 0 || p >= h->nphases) return fail(h, EVP_ERR_ARG, "set_microstructure: phase id out of range");
  }
  std::memcpy(h->grain.data(), grain, N * sizeof(int32_t));
  if (phase) std::memcpy(h->phase.data(), phase, N * sizeof(int32_t)); else std::fill(h->phase.begin(), h->phase.end(), 0);
  std::memcpy(h->rot.data(), rot9, 9 * N * sizeof(double));
  std::fill(h->sig.begin(), h->sig.end(), 0.0); std::fill(h->e.begin(), h->e.end(), 0.0);
  std::fill(h->epsp.begin(), h->epsp.end(), 0.0); std::fill(h->edotp.begin(), h->edotp.end(), 0.0);
  std::fill(h->gacc.begin(), h->gacc.end(), 0.0); std::fill(h->twinf.begin(), h->twinf.end(), 0.0);
  std::fill(h->de.begin(), h->de.end(), 0.0);
  std::fill(h->wrot.begin(), h->wrot.end(), 0.0); std::fill(h->twinned.begin(), h->twinned.end(), 0); h->ntwinned = 0; h->facc = 0.0;
  for (size_t v = 0; v < N; ++v) {
    const PhaseData &pd = h->ph[h->phase[v]];
    for (int s = 0; s < h->nsmax; ++s)
      h->crss[(size_t)s * N + v] = (s < pd.in.nsys) ? pd.in.tau0[pd.in.mode[s]] : 1.0;
  }
  for (int c = 0; c < 6; ++c) { h->Et[c] = h->E[c] = h->dEpend[c] = h->Edot_prev[c] = h->savg[c] = 0.0; }
  h->have_micro = true; h->in_incr = false; h->iter = 0;
  return EVP_OK;
}

int evp_set_reference_medium(evp_handle h, const double *c0) {
  if (!h) return EVP_ERR_ARG;
  if (c0) {
    voigt_to_mandel(c0, h->C0);
  } else {
    if (!h->have_micro) return fail(h, EVP_ERR_STATE, "reference medium average needs the microstructure");
    double acc[36] = {0};
    const size_t N = h->N;
#pragma omp parallel for schedule(static) reduction(+ : acc[:36])
    for (size_t v = 0; v < N; ++v) {
      const PhaseData &pd = h->ph[h->phase[v]];
      double R[9], Q[36], T[36];
      for (int k = 0; k < 9; ++k) R[k] = h->rot[k * N + v];
      mandel_rotation(R, Q);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += Q[6 * i + k] * pd.Cm[6 * k + j];
          T[6 * i + j] = s;
        }
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += T[6 * i + k] * Q[6 * j + k];
          acc[6 * i + j] += s;
        }
    }
    for (int k = 0; k < 36; ++k) h->C0[k] = acc[k] / (double)N;
    for (int i = 0; i < 6; ++i)
      for (int j = i + 1; j < 6; ++j) h->C0[6 * i + j] = h->C0[6 * j + i] = 0.5 * (h->C0[6 * i + j] + h->C0[6 * j + i]);
  }
  if (!mat6_inverse(h->C0, h->S0)) return fail(h, EVP_ERR_NUMERIC, "reference medium is singular");
  h->have_c0 = true;
  return EVP_OK;
}

int evp_get_reference_medium(evp_handle h, double *c0) {
  if (!h || !c0 || !h->have_c0) return EVP_ERR_STATE;
  mandel_to_voigt(h->C0, c0);
  return EVP_OK;
}

int evp_set_control(evp_handle h, const evp_ctrl *c) {
  if (!h || !c) return EVP_ERR_ARG;
  h->ctrl = *c;
  return EVP_OK;
}

int evp_set_loading(evp_handle h, const int32_t iudot[9], const double udot[9], const int32_t iscau[6], const double scau[6]) {
  if (!h || !iudot || !udot || !iscau || !scau) return fail(h, EVP_ERR_ARG, "set_loading: null pointer");
  for (int c = 0; c < 6; ++c) {
    const int i = kI[c], j = kJ[c];
    const bool sc = iudot[3 * i + j] && iudot[3 * j + i];
    if (sc == (iscau[c] != 0)) return fail(h, EVP_ERR_ARG, "set_loading: each symmetric component needs exactly one of strain-rate / stress imposed");
    h->strain_ctl[c] = sc;
  }
  std::memcpy(h->iudot, iudot, sizeof(h->iudot)); std::memcpy(h->udot, udot, sizeof(h->udot));
  std::memcpy(h->iscau, iscau, sizeof(h->iscau)); std::memcpy(h->scau, scau, sizeof(h->scau));
  h->have_loading = true;
  return EVP_OK;
}

int evp_begin_increment(evp_handle h, double dt) {
  if (!h) return EVP_ERR_ARG;
  if (!h->have_micro || !h->have_c0 || !h->have_loading) return fail(h, EVP_ERR_STATE, "begin_increment: microstructure, reference medium and loading must be set");
  if (!(dt > 0)) return fail(h, EVP_ERR_ARG, "dt must be positive");
  h->dt = dt;
  for (int c = 0; c < 6; ++c) {
    const int i = kI[c], j = kJ[c];
    const double rate = h->strain_ctl[c] ? 0.5 * (h->udot[3 * i + j] + h->udot[3 * j + i]) : h->Edot_prev[c];
    h->dEpend[c] = dt * rate;
    h->E[c] = h->Et[c] + h->dEpend[c];
  }
  h->iter = 0; h->in_incr = true;
  return EVP_OK;
}

int evp_op_green(evp_handle h) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "op_green outside an increment");
  op_green(h);
  return EVP_OK;
}

int evp_op_constitutive(evp_handle h, evp_iter_report *rep) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "op_constitutive outside an increment");
  evp_iter_report tmp{};
  op_constitutive(h, &tmp);
  h->iter += 1;
  macro_update(h);
  fill_report(h, &tmp);
  if (rep) *rep = tmp;
  return tmp.nonfinite ? EVP_ERR_NUMERIC : EVP_OK;
}

int evp_equilibrium_iter(evp_handle h, evp_iter_report *rep) {
  int rc = evp_op_green(h);
  if (rc) return rc;
  return evp_op_constitutive(h, rep);
}

int evp_equilibrium_iters(evp_handle h, int32_t n, evp_iter_report *last) {
  int rc = EVP_OK;
  for (int i = 0; i < n && rc == EVP_OK; ++i) rc = evp_equilibrium_iter(h, last);
  return rc;
}

int evp_end_increment(evp_handle h, evp_step_report *rep) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "end_increment outside an increment");
  const size_t N = h->N;
  const bool tex = h->ctrl.update_texture != 0, twn = h->ctrl.update_twinning != 0;
  std::vector<double> wnew;
  if (tex) local_rotation(h, wnew);
  double wapp[3] = {0.5 * (h->udot[3 * 2 + 1] - h->udot[3 * 1 + 2]), 0.5 * (h->udot[3 * 0 + 2] - h->udot[3 * 2 + 0]),
                    0.5 * (h->udot[3 * 1 + 0] - h->udot[3 * 0 + 1])};
  double epsum[6] = {0, 0, 0, 0, 0, 0}, dfsum = 0;
#pragma omp parallel for schedule(static) reduction(+ : epsum[:6], dfsum)
  for (size_t v = 0; v < N; ++v) {
    const PhaseData &pd = h->ph[h->phase[v]];
    VoxelFrame F;
    voxel_frame(h, v, F);
    double s6[6], edp[6] = {0, 0, 0, 0, 0, 0}, dg[EVP_MAX_SYS], gdv[EVP_MAX_SYS], dG = 0, wpc[3] = {0, 0, 0};
    for (int c = 0; c < 6; ++c) s6[c] = kW[c] * h->sig[c * N + v];
    for (int s = 0; s < pd.in.nsys; ++s) {
      double tau = 0, gd, dgd;
      for (int c = 0; c < 6; ++c) tau += F.msample[s][c] * s6[c];
      slip_rate(pd.in, s, tau, h->crss[(size_t)s * N + v], gd, dgd);
      for (int c = 0; c < 6; ++c) edp[c] += gd * F.msample[s][c];
      gdv[s] = gd;
      dg[s] = std::fabs(gd) * h->dt;
      dG += dg[s];
      for (int k = 0; k < 3; ++k) wpc[k] += pd.alpha[s][k] * gd;   // plastic spin, crystal frame
    }
    for (int c = 0; c < 6; ++c) {
      h->edotp[c * N + v] = edp[c] / kW[c];
      h->epsp[c * N + v] += h->dt * edp[c] / kW[c];
      epsum[c] += h->epsp[c * N + v];
    }
    // extended Voce, integrated analytically over the accumulated-shear step (Tome et al. 1984)
    const double G0 = h->gacc[v];
    if (dG > 0) {
      double dtau[EVP_MAX_SYS];
      for (int s = 0; s < pd.in.nsys; ++s) {
        const int m = pd.in.mode[s];
        const double dvoce = voce_tau(pd.in, m, G0 + dG) - voce_tau(pd.in, m, G0);
        double hs = 0;
        for (int s2 = 0; s2 < pd.in.nsys; ++s2) hs += pd.in.hlat[m][pd.in.mode[s2]] * dg[s2];
        dtau[s] = dvoce * hs / dG;
      }
      for (int s = 0; s < pd.in.nsys; ++s) h->crss[(size_t)s * N + v] += dtau[s];
      h->gacc[v] = G0 + dG;
    }
    if (twn) {  // twin volume fractions: df = dgamma / S_tw
      for (int s = 0; s < pd.in.nsys; ++s) {
        const int m = pd.in.mode[s];
        if (pd.in.twin[m] && pd.in.twin_shear[m] > 0) {
          const double df = gdv[s] * h->dt / pd.in.twin_shear[m];   // reoriented voxels keep contributing: F_acc is a history sum
          h->twinf[(size_t)s * N + v] += df;
          dfsum += df;
        }
      }
    }
    if (tex) {  // lattice spin = applied spin + local (FFT) spin - plastic spin
      double R[9], wps[3], dw[3];
      for (int k = 0; k < 9; ++k) R[k] = h->rot[k * N + v];
      for (int i = 0; i < 3; ++i) wps[i] = R[3 * i] * wpc[0] + R[3 * i + 1] * wpc[1] + R[3 * i + 2] * wpc[2];
      for (int k = 0; k < 3; ++k) {
        dw[k] = h->dt * wapp[k] + (wnew[k * N + v] - h->wrot[k * N + v]) - h->dt * wps[k];
        h->wrot[k * N + v] = wnew[k * N + v];
      }
      rotate_lattice(R, dw);
      for (int k = 0; k < 9; ++k) h->rot[k * N + v] = R[k];
    }
  }
  // PTR: a voxel whose predominant twin system exceeds thr1 + thr2 * F_eff / F_acc takes the twin orientation
  long long nre = 0;
  if (twn) h->facc += dfsum / (double)N;   // Tome, Lebensohn, Kocks 1991: accumulated, never reduced by a reorientation
  const double Facc = h->facc, Feff = (double)h->ntwinned / (double)N;
  if (twn) {
#pragma omp parallel for schedule(static) reduction(+ : nre)
    for (size_t v = 0; v < N; ++v) {
      if (h->twinned[v]) continue;
      const PhaseData &pd = h->ph[h->phase[v]];
      const double thr = pd.in.twin_thr1 + ((Facc > 0) ? pd.in.twin_thr2 * Feff / Facc : 0.0);
      int best = -1;
      double fb = 0;
      for (int s = 0; s < pd.in.nsys; ++s)
        if (pd.in.twin[pd.in.mode[s]] && h->twinf[(size_t)s * N + v] > fb) { fb = h->twinf[(size_t)s * N + v]; best = s; }
      if (best < 0 || !(fb > thr)) continue;
      double n[3], nl = 0, R[9], Rn[9];
      for (int k = 0; k < 3; ++k) nl += pd.in.n[best][k] * pd.in.n[best][k];
      nl = std::sqrt(nl);
      for (int k = 0; k < 3; ++k) n[k] = pd.in.n[best][k] / nl;
      for (int k = 0; k < 9; ++k) R[k] = h->rot[k * N + v];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {  // R (2 n n^T - I)
          double s = 0;
          for (int k = 0; k < 3; ++k) s += R[3 * i + k] * (2.0 * n[k] * n[j] - (k == j ? 1.0 : 0.0));
          Rn[3 * i + j] = s;
        }
      for (int k = 0; k < 9; ++k) h->rot[k * N + v] = Rn[k];
      for (int s = 0; s < pd.in.nsys; ++s) h->twinf[(size_t)s * N + v] = 0.0;
      h->twinned[v] = 1;
      nre += 1;
    }
    h->ntwinned += nre;
  }
  // the pending macro correction belongs to an iteration that will not run: drop it, so that
  // the committed E is the mean of the committed strain field e
  for (int c = 0; c < 6; ++c) {
    h->E[c] -= h->dEpend[c];
    h->dEpend[c] = 0.0;
    h->Edot_prev[c] = (h->E[c] - h->Et[c]) / h->dt;
    h->Et[c] = h->E[c];
  }
  h->in_incr = false;
  if (rep) {
    rep->iters = h->iter;
    rep->err_stress = h->last_err_s; rep->err_strain = h->last_err_e;
    rep->converged = (h->last_err_s <= h->ctrl.tol_stress && h->last_err_e <= h->ctrl.tol_strain && h->last_unconverged == 0) ? 1 : 0;
    for (int c = 0; c < 6; ++c) { rep->savg[c] = h->savg[c]; rep->emacro[c] = h->E[c]; rep->epavg[c] = epsum[c] / (double)N; }
    rep->twin_acc = Facc; rep->twin_eff = (double)h->ntwinned / (double)N; rep->reoriented = nre;
  }
  return EVP_OK;
}

int evp_step(evp_handle h, double dt, evp_step_report *rep) {
  const auto t0 = std::chrono::steady_clock::now();
  int rc = evp_begin_increment(h, dt);
  if (rc) return rc;
  evp_iter_report ir{};
  for (int it = 0; it < h->ctrl.itmax; ++it) {
    rc = evp_equilibrium_iter(h, &ir);
    if (rc) return rc;
    if (ir.converged) break;
  }
  rc = evp_end_increment(h, rep);
  if (rep) rep->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

int evp_field_components(evp_handle h, evp_field f) { return h ? (int)field_comps(h, f) : EVP_ERR_ARG; }

static void *field_ptr(evp_solver *S, int f, size_t *elem) {
  *elem = sizeof(double);
  switch (f) {
    case EVP_FIELD_STRESS: return S->sig.data();
    case EVP_FIELD_STRAIN: return S->e.data();
    case EVP_FIELD_PLASTIC_STRAIN: return S->epsp.data();
    case EVP_FIELD_PLASTIC_RATE: return S->edotp.data();
    case EVP_FIELD_CRSS: return S->crss.data();
    case EVP_FIELD_ROTATION: return S->rot.data();
    case EVP_FIELD_GAMMA_ACC: return S->gacc.data();
    case EVP_FIELD_TWIN_FRACTION: return S->twinf.data();
    case EVP_FIELD_STRAIN_INCR: return S->de.data();
    case EVP_FIELD_LOCAL_ROTATION: return S->wrot.data();
    case EVP_FIELD_TWINNED: *elem = sizeof(int32_t); return S->twinned.data();
    case EVP_FIELD_GRAIN: *elem = sizeof(int32_t); return S->grain.data();
    case EVP_FIELD_PHASE: *elem = sizeof(int32_t); return S->phase.data();
    default: return nullptr;
  }
}

int evp_get_field(evp_handle h, evp_field f, void *host, size_t bytes) {
  if (!h || !host) return EVP_ERR_ARG;
  size_t el;
  void *p = field_ptr(h, f, &el);
  if (!p) return fail(h, EVP_ERR_ARG, "unknown field");
  const size_t need = field_comps(h, f) * h->N * el;
  if (bytes != need) return fail(h, EVP_ERR_ARG, "get_field: size mismatch");
  std::memcpy(host, p, need);
  return EVP_OK;
}

int evp_set_field(evp_handle h, evp_field f, const void *host, size_t bytes) {
  if (!h || !host) return EVP_ERR_ARG;
  size_t el;
  void *p = field_ptr(h, f, &el);
  if (!p) return fail(h, EVP_ERR_ARG, "unknown field");
  const size_t need = field_comps(h, f) * h->N * el;
  if (bytes != need) return fail(h, EVP_ERR_ARG, "set_field: size mismatch");
  std::memcpy(p, host, need);
  return EVP_OK;
}

int evp_get_macro(evp_handle h, double emacro[6], double savg[6]) {
  if (!h) return EVP_ERR_ARG;
  for (int c = 0; c < 6; ++c) { if (emacro) emacro[c] = h->E[c]; if (savg) savg[c] = h->savg[c]; }
  return EVP_OK;
}


// restart file format "EVPCKPT2", identical to the product library's (include/evpfft.h)
struct CkptHeader {
  char magic[8];
  int32_t nx, ny, nz, y0, nyl, z0, nzl, nsmax, nranks, rank;
  double Et[6], Edot_prev[6];
  int64_t ntwinned;
  double facc;
};

int evp_save_state(evp_handle h, const char *path) {
  if (!h || !path) return EVP_ERR_ARG;
  std::FILE *f = std::fopen(path, "wb");
  if (!f) return fail(h, EVP_ERR_ARG, std::string("cannot write ") + path);
  CkptHeader hd{};
  std::memcpy(hd.magic, "EVPCKPT2", 8);
  hd.nx = h->nx; hd.ny = h->ny; hd.nz = h->nz; hd.y0 = 0; hd.nyl = h->ny; hd.z0 = 0; hd.nzl = h->nz; hd.nsmax = h->nsmax; hd.nranks = 1; hd.rank = 0;
  for (int c = 0; c < 6; ++c) { hd.Et[c] = h->Et[c]; hd.Edot_prev[c] = h->Edot_prev[c]; }
  hd.ntwinned = h->ntwinned;
  hd.facc = h->facc;
  bool ok = std::fwrite(&hd, sizeof(hd), 1, f) == 1;
  auto wr = [&](const void *p, size_t bytes) { ok = ok && std::fwrite(p, 1, bytes, f) == bytes; };
  const size_t N = h->N, ns = (size_t)std::max(h->nsmax, 1);
  wr(h->sig.data(), 6 * N * 8); wr(h->e.data(), 6 * N * 8); wr(h->epsp.data(), 6 * N * 8); wr(h->crss.data(), ns * N * 8);
  wr(h->rot.data(), 9 * N * 8); wr(h->gacc.data(), N * 8); wr(h->twinf.data(), ns * N * 8); wr(h->wrot.data(), 3 * N * 8);
  wr(h->grain.data(), N * 4); wr(h->phase.data(), N * 4); wr(h->twinned.data(), N * 4);
  std::fclose(f);
  return ok ? EVP_OK : fail(h, EVP_ERR_ARG, "short write");
}

int evp_load_state(evp_handle h, const char *path) {
  if (!h || !path) return EVP_ERR_ARG;
  std::FILE *f = std::fopen(path, "rb");
  if (!f) return fail(h, EVP_ERR_ARG, std::string("cannot read ") + path);
  CkptHeader hd{};
  bool ok = std::fread(&hd, sizeof(hd), 1, f) == 1 && std::memcmp(hd.magic, "EVPCKPT2", 8) == 0;
  if (ok && (hd.nx != h->nx || hd.ny != h->ny || hd.nz != h->nz || hd.nzl != h->nz || hd.nyl != h->ny || hd.nsmax != h->nsmax)) ok = false;
  if (!ok) { std::fclose(f); return fail(h, EVP_ERR_ARG, "checkpoint does not match this handle"); }
  {
    const size_t N = h->N, ns = (size_t)std::max(h->nsmax, 1);
    const size_t need = sizeof(hd) + (size_t)8 * N * (6 + 6 + 6 + ns + 9 + 1 + ns + 3) + (size_t)4 * N * 3;
    std::fseek(f, 0, SEEK_END);
    const long have = std::ftell(f);
    std::fseek(f, (long)sizeof(hd), SEEK_SET);
    if (have < 0 || (size_t)have != need) { std::fclose(f); return fail(h, EVP_ERR_ARG, "checkpoint file is truncated or has trailing bytes"); }
  }
  auto rd = [&](void *p, size_t bytes) { ok = ok && std::fread(p, 1, bytes, f) == bytes; };
  const size_t N = h->N, ns = (size_t)std::max(h->nsmax, 1);
  rd(h->sig.data(), 6 * N * 8); rd(h->e.data(), 6 * N * 8); rd(h->epsp.data(), 6 * N * 8); rd(h->crss.data(), ns * N * 8);
  rd(h->rot.data(), 9 * N * 8); rd(h->gacc.data(), N * 8); rd(h->twinf.data(), ns * N * 8); rd(h->wrot.data(), 3 * N * 8);
  rd(h->grain.data(), N * 4); rd(h->phase.data(), N * 4); rd(h->twinned.data(), N * 4);
  std::fclose(f);
  if (!ok) return fail(h, EVP_ERR_ARG, "short read");
  for (int c = 0; c < 6; ++c) { h->Et[c] = hd.Et[c]; h->E[c] = hd.Et[c]; h->Edot_prev[c] = hd.Edot_prev[c]; h->dEpend[c] = 0; }
  h->ntwinned = hd.ntwinned;
  h->facc = hd.facc;
  h->have_micro = true; h->in_incr = false;
  return EVP_OK;
}

int evp_debug_spectrum(evp_handle h, int32_t comp, double *out) {
  if (!h || !out || comp < 0 || comp > 5) return EVP_ERR_ARG;
  const size_t NS = (size_t)h->nz * h->ny * h->nxh;
  std::vector<cplx> A(NS), B(NS);
  const int other = comp ^ 1;
  const int a = std::min(comp, other), b = std::max(comp, other);
  fft3_forward_pair(h, &h->sig[(size_t)a * h->N], &h->sig[(size_t)b * h->N], A.data(), B.data());
  const cplx *src = (comp == a) ? A.data() : B.data();
  for (size_t i = 0; i < NS; ++i) { out[2 * i] = src[i].real(); out[2 * i + 1] = src[i].imag(); }
  return EVP_OK;
}

void *evp_stream(evp_handle) { return nullptr; }
int evp_set_profiling(evp_handle, int32_t) { return EVP_OK; }
int evp_last_kernel_ms(evp_handle, double ms[8]) { for (int i = 0; i < 8; ++i) ms[i] = 0; return EVP_OK; }

int evp_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// n <= 0: every online core.  bench.py calls this so that a launcher's OMP_NUM_THREADS=1 (torchrun) does not
// silently turn the CPU baseline into a single-core run.
int evp_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n <= 0) n = omp_get_num_procs();
  omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

}  // extern "C"
