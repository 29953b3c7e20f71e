// kernels.cu — see kernels.cuh for the kernel list.  sm_100a only.
#include "kernels.cuh"

#include <cstdio>

namespace evp {

__constant__ PhaseDev c_phase[EVP_MAX_PHASES];
__constant__ GreenConst c_green;
__constant__ ConstParams c_cp;

void upload_phase_tables(const PhaseDev *ph, int nph) { cudaMemcpyToSymbol(c_phase, ph, sizeof(PhaseDev) * nph); }
void upload_green(const GreenConst &g) { cudaMemcpyToSymbol(c_green, &g, sizeof(g)); }
void upload_const_params(const ConstParams &p) { cudaMemcpyToSymbol(c_cp, &p, sizeof(p)); }

// ---------------------------------------------------------------------------------------------
// block-level Stockham FFT over lines in shared memory.  Thread owns (line via `base`, q).
// ---------------------------------------------------------------------------------------------
struct TwLdg {
  const double2 *t;
  __device__ __forceinline__ double2 operator()(int k) const { return __ldg(t + k); }
};
template <int ES>
struct OffES {
  int base;
  __device__ __forceinline__ int operator()(int i) const { return base + i * ES; }
};

template <int N, int R, int NS, bool INV, class OFF>
__device__ __forceinline__ void fft_pass(double2 *s, int q, OFF off, TwLdg tw) {
  double2 v[8];
  pass_load<N, R>(s, q, v, off);
  __syncthreads();
  pass_store<N, R, NS, INV>(s, q, v, off, tw);
  __syncthreads();
}

template <int N, int NS, bool INV, class OFF>
__device__ __forceinline__ void fft_rest(double2 *s, int q, OFF off, TwLdg tw) {
  if constexpr (NS < N) {
    fft_pass<N, 8, NS, INV>(s, q, off, tw);
    fft_rest<N, NS * 8, INV>(s, q, off, tw);
  }
}

// all threads of the block must call this (contains __syncthreads); threads with active == false
// only take part in the barriers.
template <int N, bool INV, class OFF>
__device__ __forceinline__ void block_fft(double2 *s, int q, OFF off, TwLdg tw) {
  constexpr int R0 = first_radix(N);
  fft_pass<N, R0, 1, INV>(s, q, off, tw);
  fft_rest<N, R0, INV>(s, q, off, tw);
}

// ---------------------------------------------------------------------------------------------
// K2: x forward.  Block = XL rows of one component pair.
// ---------------------------------------------------------------------------------------------
template <int NX>
struct XCfg {
  static constexpr int L = (2048 / NX > 8) ? 2048 / NX : 8;
  static constexpr int T = L * NX / 8;
  static constexpr int LS = NX + 1;
  static constexpr size_t smem = (size_t)L * LS * sizeof(double2);
};

template <int NX>
__global__ void __launch_bounds__(XCfg<NX>::T) k_xfwd(const double *__restrict__ sig, double2 *__restrict__ W, long long N,
                                                      int nrows, SpecLayout Lay, int ny, const double2 *__restrict__ twp) {
  using C = XCfg<NX>;
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * C::L;
  const int pair = blockIdx.y;
  const double *a = sig + (long long)(2 * pair) * N + (long long)row0 * NX;
  const double *b = a + N;
  const int nrl = min(C::L, nrows - row0);
  // coalesced load of L consecutive rows of both fields
  for (int idx = tid * 2; idx < nrl * NX; idx += C::T * 2) {
    const int l = idx / NX, x = idx % NX;
    const double2 av = *reinterpret_cast<const double2 *>(a + idx);
    const double2 bv = *reinterpret_cast<const double2 *>(b + idx);
    sm[l * C::LS + x] = make_double2(av.x, bv.x);
    sm[l * C::LS + x + 1] = make_double2(av.y, bv.y);
  }
  for (int idx = nrl * NX + tid; idx < C::L * NX; idx += C::T) sm[(idx / NX) * C::LS + idx % NX] = make_double2(0.0, 0.0);
  __syncthreads();
  const int l = tid % C::L, q = tid / C::L;
  block_fft<NX, false>(sm, q, OffES<1>{l * C::LS}, TwLdg{twp});
  // separate the two spectra: A = (Z(k) + conj Z(-k))/2,  B = (Z(k) - conj Z(-k))/(2i)
  const int nxh = NX / 2 + 1;
  for (int idx = tid; idx < nrl * nxh; idx += C::T) {
    const int ll = idx / nxh, k = idx % nxh;
    const double2 zk = sm[ll * C::LS + k];
    const double2 zm = sm[ll * C::LS + ((NX - k) & (NX - 1))];
    const double2 A = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
    const double2 B = make_double2(0.5 * (zk.y + zm.y), 0.5 * (zm.x - zk.x));
    const int row = row0 + ll;
    const int zl = row / ny, y = row % ny;
    const long long o = Lay.row_ysplit(2 * pair, zl, y) + k;
    W[o] = A;
    W[o + Lay.cstride] = B;
  }
}

// ---------------------------------------------------------------------------------------------
// K6: x inverse fused with the strain update  e <- e - de + dE  (row a3)
// ---------------------------------------------------------------------------------------------
template <int NX>
__global__ void __launch_bounds__(XCfg<NX>::T) k_xinv(const double2 *__restrict__ W, double *__restrict__ e, double *__restrict__ de_dbg,
                                                      const MacroDev *__restrict__ macro, long long N, int nrows, SpecLayout Lay, int ny,
                                                      const double2 *__restrict__ twp) {
  using C = XCfg<NX>;
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * C::L;
  const int pair = blockIdx.y;
  const int nrl = min(C::L, nrows - row0);
  const int nxh = NX / 2 + 1;
  // Z(k) = A(k) + i B(k), Hermitian completion Z(N-k) = conj(A(k)) + i conj(B(k))
  for (int idx = tid; idx < C::L * nxh; idx += C::T) {
    const int ll = idx / nxh, k = idx % nxh;
    double2 A = make_double2(0.0, 0.0), B = A;
    if (ll < nrl) {
      const int row = row0 + ll;
      const int zl = row / ny, y = row % ny;
      const long long o = Lay.row_ysplit(2 * pair, zl, y) + k;
      A = W[o];
      B = W[o + Lay.cstride];
    }
    if (k == 0 || k == NX / 2) { A.y = 0.0; B.y = 0.0; }  // real by Hermitian symmetry
    sm[ll * C::LS + k] = make_double2(A.x - B.y, A.y + B.x);
    if (k > 0 && k < NX / 2) sm[ll * C::LS + NX - k] = make_double2(A.x + B.y, B.x - A.y);
  }
  __syncthreads();
  const int l = tid % C::L, q = tid / C::L;
  block_fft<NX, true>(sm, q, OffES<1>{l * C::LS}, TwLdg{twp});
  const double dEa = macro->dEpend[2 * pair], dEb = macro->dEpend[2 * pair + 1];
  double *ea = e + (long long)(2 * pair) * N + (long long)row0 * NX;
  double *eb = ea + N;
  for (int idx = tid * 2; idx < nrl * NX; idx += C::T * 2) {
    const int ll = idx / NX, x = idx % NX;
    const double2 z0 = sm[ll * C::LS + x], z1 = sm[ll * C::LS + x + 1];
    double2 va = *reinterpret_cast<double2 *>(ea + idx);
    double2 vb = *reinterpret_cast<double2 *>(eb + idx);
    va.x += dEa - z0.x; va.y += dEa - z1.x;
    vb.x += dEb - z0.y; vb.y += dEb - z1.y;
    *reinterpret_cast<double2 *>(ea + idx) = va;
    *reinterpret_cast<double2 *>(eb + idx) = vb;
    if (de_dbg) {
      double *da = de_dbg + (long long)(2 * pair) * N + (long long)row0 * NX;
      *reinterpret_cast<double2 *>(da + idx) = make_double2(z0.x, z1.x);
      *reinterpret_cast<double2 *>(da + N + idx) = make_double2(z0.y, z1.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3 / K5: y pass.  Block = one component, ZT planes, 8 kx columns, all ny rows.
// ---------------------------------------------------------------------------------------------
template <int NY>
struct YCfg {
  static constexpr int TX = 8;
  static constexpr int ZT = (256 / NY > 1) ? 256 / NY : 1;
  static constexpr int L = TX * ZT;
  static constexpr int T = L * NY / 8;
  static constexpr size_t smem = (size_t)ZT * NY * TX * sizeof(double2);
};

template <int NY, bool INV>
__global__ void __launch_bounds__(YCfg<NY>::T) k_ypass(const double2 *__restrict__ in, double2 *__restrict__ out, SpecLayout Lin,
                                                       SpecLayout Lout, int nzl, const double2 *__restrict__ twp) {
  using C = YCfg<NY>;
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * C::TX;
  const int z0 = blockIdx.y * C::ZT;
  const int c = blockIdx.z;
  const int nxh = Lin.nxh;
  for (int idx = tid; idx < C::ZT * NY * C::TX; idx += C::T) {
    const int col = idx % C::TX, row = (idx / C::TX) % NY, zt = idx / (C::TX * NY);
    double2 v = make_double2(0.0, 0.0);
    if (k0 + col < nxh && z0 + zt < nzl) v = in[Lin.row_ysplit(c, z0 + zt, row) + k0 + col];
    sm[idx] = v;
  }
  __syncthreads();
  const int l = tid % C::L, q = tid / C::L;
  block_fft<NY, INV>(sm, q, OffES<C::TX>{(l / C::TX) * (NY * C::TX) + (l % C::TX)}, TwLdg{twp});
  for (int idx = tid; idx < C::ZT * NY * C::TX; idx += C::T) {
    const int col = idx % C::TX, row = (idx / C::TX) % NY, zt = idx / (C::TX * NY);
    if (k0 + col < nxh && z0 + zt < nzl) out[Lout.row_ysplit(c, z0 + zt, row) + k0 + col] = sm[idx];
  }
}

// ---------------------------------------------------------------------------------------------
// K4: fused z pass:  forward FFT of all 6 components, Green operator, inverse FFT.
// Block = one ky row, TX kx columns, 6 components, all nz points.
// ---------------------------------------------------------------------------------------------
template <int NZ>
struct ZCfg {
  static constexpr int TX = (NZ >= 1024) ? 2 : ((NZ >= 256) ? 4 : 8);
  static constexpr int TPC = TX * NZ / 8;                       // threads per component
  static constexpr int CG = (TPC >= 384) ? 1 : (2 * TPC >= 384 ? 2 : (3 * TPC >= 384 ? 3 : 6));
  static constexpr int T = CG * TPC;
  static constexpr int CS = NZ * TX;                            // component stride in smem (elements)
  static constexpr size_t smem = (size_t)6 * CS * sizeof(double2);
};

template <int NZ, int MODE>  // MODE 0: fused fwd+Green+inv; MODE 1: forward only (evp_debug_spectrum)
__global__ void __launch_bounds__(ZCfg<NZ>::T) k_zfused(double2 *__restrict__ Wt, SpecLayout Lay, int ky0, int nx, int ny, double rx,
                                                        double ry, double rz, double scale, const double2 *__restrict__ twp) {
  using C = ZCfg<NZ>;
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * C::TX;
  const int yl = blockIdx.y;
  const int nxh = Lay.nxh;
  // load [c][z][TX]
  for (int idx = tid; idx < 6 * C::CS; idx += C::T) {
    const int col = idx % C::TX, z = (idx / C::TX) % NZ, c = idx / C::CS;
    double2 v = make_double2(0.0, 0.0);
    if (k0 + col < nxh) v = Wt[Lay.row_zsplit(c, z, yl) + k0 + col];
    sm[idx] = v;
  }
  __syncthreads();
  const int cg = tid / C::TPC, t = tid % C::TPC;
  const int col = t % C::TX, q = t / C::TX;
#pragma unroll 1
  for (int c = cg; c < 6; c += C::CG) block_fft<NZ, false>(sm, q, OffES<C::TX>{c * C::CS + col}, TwLdg{twp});
  if (MODE == 0) {
  // Green operator per frequency (row a2)
  const int ky = ky0 + yl;
  const int fy = (ky <= ny / 2) ? ky : ky - ny;
  for (int idx = tid; idx < C::CS; idx += C::T) {
    const int cc = idx % C::TX, kz = idx / C::TX;
    const int kx = k0 + cc;
    if (kx < nxh) {
      const int fz = (kz <= NZ / 2) ? kz : kz - NZ;
      double2 lam[6], o[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) lam[c] = sm[c * C::CS + idx];
      const bool zero = (kx == 0) && (ky == 0) && (kz == 0);
      const bool nyq = (kx * 2 == nx) || (ky * 2 == ny) || (kz * 2 == NZ);
      green_point(c_green, kx * rx, fy * ry, fz * rz, zero, nyq, scale, lam, o);
#pragma unroll
      for (int c = 0; c < 6; ++c) sm[c * C::CS + idx] = o[c];
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int c = cg; c < 6; c += C::CG) block_fft<NZ, true>(sm, q, OffES<C::TX>{c * C::CS + col}, TwLdg{twp});
  }
  for (int idx = tid; idx < 6 * C::CS; idx += C::T) {
    const int cc = idx % C::TX, z = (idx / C::TX) % NZ, c = idx / C::CS;
    if (k0 + cc < nxh) Wt[Lay.row_zsplit(c, z, yl) + k0 + cc] = sm[idx];
  }
}

// ---------------------------------------------------------------------------------------------
// K1: constitutive update (rows a4, a5, a6).  One thread per voxel, 128 threads per block.
// ---------------------------------------------------------------------------------------------
constexpr int kCB = 128;
int constitutive_block() { return kCB; }

struct ItcSmem {
  const double *p;  // &smem[tid], stride kCB
  __device__ __forceinline__ double operator()(int s) const { return p[s * kCB]; }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block reduction of NV values (sum) + one max; result written by thread 0 to out[0..NV], out[NV] = max
template <int NV>
__device__ __forceinline__ void block_reduce_store(double vals[NV], int vmax, double *out) {
  __shared__ double red[kCB / 32][NV + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(vals[k]);
    if (lane == 0) red[w][k] = s;
  }
  const int m = warp_max(vmax);
  if (lane == 0) red[w][NV] = (double)m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < kCB / 32; ++ww) s += red[ww][k];
      out[k] = s;
    }
    double mm = 0.0;
#pragma unroll
    for (int ww = 0; ww < kCB / 32; ++ww) mm = fmax(mm, red[ww][NV]);
    out[NV] = mm;
  }
}

__global__ void __launch_bounds__(kCB) k_constitutive(Fields f, double *__restrict__ partials) {
  extern __shared__ double itc_sm[];  // [nsmax][kCB]
  const int tid = threadIdx.x;
  const long long v = (long long)blockIdx.x * kCB + tid;
  const long long N = f.N;
  double vals[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // ds, de, sig[6], nit, bad
  int nit = 0;
  if (v < N) {
    const PhaseDev &P = c_phase[f.phase[v]];
    double R[9], sig[6], em[6];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = __ldg(f.rot + k * N + v);
#pragma unroll
    for (int c = 0; c < 6; ++c) sig[c] = f.sig[c * N + v];
#pragma unroll
    for (int c = 0; c < 6; ++c) em[c] = f.e[c * N + v] - __ldg(f.epsp + c * N + v);
    const int ns = P.nsys;
    for (int s = 0; s < ns; ++s) itc_sm[s * kCB + tid] = 1.0 / __ldg(f.crss + (long long)s * N + v);
    int bad = 0;
    double ds, de;
    nit = constitutive_voxel(P, c_cp, R, sig, em, ItcSmem{itc_sm + tid}, &ds, &de, &bad);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      f.sig[c * N + v] = sig[c];
      vals[2 + c] = sig[c];
    }
    vals[0] = ds;
    vals[1] = de;
    vals[8] = (double)nit;
    vals[9] = (double)bad;
  }
  block_reduce_store<10>(vals, nit, partials + (long long)blockIdx.x * kPartial);
}

// second stage of the reductions: fixed-order sum of the block partials (deterministic)
__global__ void __launch_bounds__(256) k_reduce(const double *__restrict__ partials, int nblocks, double *__restrict__ totals) {
  __shared__ double red[256][12];
  const int tid = threadIdx.x;
  double acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  for (int b = tid; b < nblocks; b += 256) {
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] += partials[(long long)b * kPartial + k];
    acc[10] = fmax(acc[10], partials[(long long)b * kPartial + 10]);
  }
#pragma unroll
  for (int k = 0; k < 11; ++k) red[tid][k] = acc[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
#pragma unroll
      for (int k = 0; k < 10; ++k) red[tid][k] += red[tid + s][k];
      red[tid][10] = fmax(red[tid][10], red[tid + s][10]);
    }
    __syncthreads();
  }
  if (tid < 11) totals[tid] = red[0][tid];
}

// rows a6 (normalisation) + a7 (macro strain correction) on the device.
// totals: [0] sum|dsig| [1] sum|S0 dsig| [2..7] sum sig [8] sum nit [9] bad [10] max nit
__global__ void k_macro(const double *__restrict__ totals, MacroDev *__restrict__ m, double ntot) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sn = 0.0, en = 0.0;
  for (int c = 0; c < 6; ++c) {
    m->savg[c] = totals[2 + c] / ntot;
    const double w = (c < 3) ? 1.0 : 2.0;
    sn += w * m->savg[c] * m->savg[c];
    en += w * m->E[c] * m->E[c];
  }
  const double es = totals[0] / ntot, ee = totals[1] / ntot;
  m->err_s = (sn > 0.0) ? es / sqrt(sn) : es;
  m->err_e = (en > 0.0) ? ee / sqrt(en) : ee;
  m->newton_mean = totals[8] / ntot;
  m->newton_max = (int)totals[10];
  m->nonfinite = (int)totals[9];
  m->iter += 1;
  for (int a = 0; a < 6; ++a) {
    double acc = 0.0;
    for (int b = 0; b < 6; ++b) acc += m->Mmac[6 * a + b] * (m->scau[b] - m->savg[b]);
    m->dEpend[a] = acc;
    m->E[a] += acc;
  }
}

// ---------------------------------------------------------------------------------------------
// per-increment commit (§8(f).1): eps_p += dt*edp(sig), extended Voce hardening
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double voce_tau(const PhaseDev &P, int m, double G) {
  const double t0 = P.tau0[m], t1 = P.tau1[m], h0 = P.theta0[m], h1 = P.theta1[m];
  if (fabs(t1) < 1e-300) return t0 + h1 * G;
  return t0 + (t1 + h1 * G) * (1.0 - exp(-G * fabs(h0 / t1)));
}

__global__ void __launch_bounds__(kCB) k_commit(Fields f, double dt, double *__restrict__ partials) {
  extern __shared__ double dg_sm[];  // [nsmax][kCB] |dgamma|
  const int tid = threadIdx.x;
  const long long v = (long long)blockIdx.x * kCB + tid;
  const long long N = f.N;
  double sums[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (v < N) {
    const PhaseDev &P = c_phase[f.phase[v]];
    double R[9], M[25], t[6], sb[6], sc[6];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = f.rot[k * N + v];
    rot_b5(R, M);
#pragma unroll
    for (int c = 0; c < 6; ++c) t[c] = f.sig[c * N + v];
    cart_to_b(t, sb);
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      double x = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) x += M[b * 5 + a] * sb[b];
      sc[a] = x;
    }
    double edc[5] = {0, 0, 0, 0, 0}, dG = 0.0;
    const int ns = P.nsys;
    for (int s = 0; s < ns; ++s) {
      double tau = 0.0;
#pragma unroll
      for (int c = 0; c < 5; ++c) tau += P.m[s][c] * sc[c];
      double gd, dgd;
      slip_rate(P, s, tau, 1.0 / f.crss[(long long)s * N + v], gd, dgd);
#pragma unroll
      for (int c = 0; c < 5; ++c) edc[c] += gd * P.m[s][c];
      const double dg = fabs(gd) * dt;
      dg_sm[s * kCB + tid] = dg;
      dG += dg;
    }
    double eds[6];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      double x = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) x += M[a * 5 + b] * edc[b];
      eds[a] = x;
    }
    eds[5] = 0.0;
    b_to_cart(eds, t);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      f.edotp[c * N + v] = t[c];
      const double ep = f.epsp[c * N + v] + dt * t[c];
      f.epsp[c * N + v] = ep;
      sums[2 + c] = ep;
    }
    const double G0 = f.gacc[v];
    if (dG > 0.0) {
      for (int s = 0; s < ns; ++s) {
        const int m = P.mode[s];
        const double dv = voce_tau(P, m, G0 + dG) - voce_tau(P, m, G0);
        double hs = 0.0;
        for (int s2 = 0; s2 < ns; ++s2) hs += P.hlat[m][P.mode[s2]] * dg_sm[s2 * kCB + tid];
        f.crss[(long long)s * N + v] += dv * hs / dG;
      }
      f.gacc[v] = G0 + dG;
    }
  }
  block_reduce_store<10>(sums, 0, partials + (long long)blockIdx.x * kPartial);
}

__global__ void k_fill(double *p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void k_init_crss(Fields f, int nsmax) {
  const long long N = f.N;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += (long long)gridDim.x * blockDim.x) {
    const PhaseDev &P = c_phase[f.phase[v]];
    for (int s = 0; s < nsmax; ++s) f.crss[(long long)s * N + v] = (s < P.nsys) ? P.tau0[P.mode[s]] : 1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
bool fft_size_supported(int n) { return n >= 8 && n <= 1024 && (n & (n - 1)) == 0; }

template <class K>
static void set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define EVP_DISPATCH_N(n, MACRO) \
  switch (n) {                   \
    case 8: MACRO(8); break;     \
    case 16: MACRO(16); break;   \
    case 32: MACRO(32); break;   \
    case 64: MACRO(64); break;   \
    case 128: MACRO(128); break; \
    case 256: MACRO(256); break; \
    case 512: MACRO(512); break; \
    case 1024: MACRO(1024); break; \
    default: break;              \
  }

void launch_xfwd(int nx, const double *sig, double2 *W, long long N, int nrows, SpecLayout L, const double2 *tw, cudaStream_t st) {
  const int ny = L.nyl;  // plain layout: nyl == ny
#define X_(NX)                                                                                        \
  {                                                                                                   \
    using C = XCfg<NX>;                                                                               \
    set_smem(k_xfwd<NX>, C::smem);                                                                    \
    dim3 grid((nrows + C::L - 1) / C::L, 3);                                                          \
    k_xfwd<NX><<<grid, C::T, C::smem, st>>>(sig, W, N, nrows, L, ny, tw);                              \
  }
  EVP_DISPATCH_N(nx, X_)
#undef X_
}

void launch_xinv(int nx, const double2 *W, double *e, double *de_dbg, const MacroDev *macro, long long N, int nrows, SpecLayout L,
                 const double2 *tw, cudaStream_t st) {
  const int ny = L.nyl;
#define X_(NX)                                                                                        \
  {                                                                                                   \
    using C = XCfg<NX>;                                                                               \
    set_smem(k_xinv<NX>, C::smem);                                                                    \
    dim3 grid((nrows + C::L - 1) / C::L, 3);                                                          \
    k_xinv<NX><<<grid, C::T, C::smem, st>>>(W, e, de_dbg, macro, N, nrows, L, ny, tw);                 \
  }
  EVP_DISPATCH_N(nx, X_)
#undef X_
}

void launch_ypass(int ny, bool inv, const double2 *in, double2 *out, SpecLayout Lin, SpecLayout Lout, int nzl,
                  const double2 *tw, cudaStream_t st) {
#define Y_(NY)                                                                                        \
  {                                                                                                   \
    using C = YCfg<NY>;                                                                               \
    dim3 grid((Lin.nxh + C::TX - 1) / C::TX, (nzl + C::ZT - 1) / C::ZT, 6);                           \
    if (inv) {                                                                                        \
      set_smem(k_ypass<NY, true>, C::smem);                                                           \
      k_ypass<NY, true><<<grid, C::T, C::smem, st>>>(in, out, Lin, Lout, nzl, tw);                    \
    } else {                                                                                          \
      set_smem(k_ypass<NY, false>, C::smem);                                                          \
      k_ypass<NY, false><<<grid, C::T, C::smem, st>>>(in, out, Lin, Lout, nzl, tw);                   \
    }                                                                                                 \
  }
  EVP_DISPATCH_N(ny, Y_)
#undef Y_
}

void launch_zfused(int nz, bool fwd_only, double2 *Wt, SpecLayout L, int nyl, int ky0, int nx, int ny, double dx, double dy, double dz,
                   const double2 *tw, cudaStream_t st) {
  const double rx = 1.0 / (nx * dx), ry = 1.0 / (ny * dy), rz = 1.0 / (nz * dz);
  const double scale = 1.0 / ((double)nx * ny * nz);
#define Z_(NZ)                                                                                        \
  {                                                                                                   \
    using C = ZCfg<NZ>;                                                                               \
    dim3 grid((L.nxh + C::TX - 1) / C::TX, nyl);                                                      \
    if (fwd_only) {                                                                                   \
      set_smem(k_zfused<NZ, 1>, C::smem);                                                             \
      k_zfused<NZ, 1><<<grid, C::T, C::smem, st>>>(Wt, L, ky0, nx, ny, rx, ry, rz, scale, tw);        \
    } else {                                                                                          \
      set_smem(k_zfused<NZ, 0>, C::smem);                                                             \
      k_zfused<NZ, 0><<<grid, C::T, C::smem, st>>>(Wt, L, ky0, nx, ny, rx, ry, rz, scale, tw);        \
    }                                                                                                 \
  }
  EVP_DISPATCH_N(nz, Z_)
#undef Z_
}

void launch_constitutive(const Fields &f, int nsmax, double *partials, int *nblocks_out, cudaStream_t st) {
  const int nb = (int)((f.N + kCB - 1) / kCB);
  const size_t smem = (size_t)(nsmax > 0 ? nsmax : 1) * kCB * sizeof(double);
  k_constitutive<<<nb, kCB, smem, st>>>(f, partials);
  if (nblocks_out) *nblocks_out = nb;
}

void launch_commit(const Fields &f, int nsmax, double dt, double *partials, int *nblocks_out, cudaStream_t st) {
  const int nb = (int)((f.N + kCB - 1) / kCB);
  const size_t smem = (size_t)(nsmax > 0 ? nsmax : 1) * kCB * sizeof(double);
  k_commit<<<nb, kCB, smem, st>>>(f, dt, partials);
  if (nblocks_out) *nblocks_out = nb;
}

void launch_reduce(const double *partials, int nblocks, double *totals, cudaStream_t st) {
  k_reduce<<<1, 256, 0, st>>>(partials, nblocks, totals);
}
void launch_macro(const double *totals, MacroDev *macro, double ntot, cudaStream_t st) { k_macro<<<1, 32, 0, st>>>(totals, macro, ntot); }
void launch_fill(double *p, long long n, double v, cudaStream_t st) { k_fill<<<592, 256, 0, st>>>(p, n, v); }
void launch_init_crss(const Fields &f, int nsmax, cudaStream_t st) { k_init_crss<<<592, 256, 0, st>>>(f, nsmax); }

}  // namespace evp
