// kernels.cu — see kernels.cuh for the kernel list.  sm_100a only.
#include "kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <string>

namespace evp {

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) + mbarrier wrappers: spectral tiles are staged global <-> shared by
// the tensor memory accelerator; one elected thread issues the copies, the block waits on an mbarrier.
// ---------------------------------------------------------------------------------------------
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void load5(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void store5(const CUtensorMap *tm, const void *src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"((uint64_t)tm),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16), completion on an mbarrier
__device__ __forceinline__ void load1(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }   // writes performed (peer stores)
}  // namespace tma

__constant__ PhaseDev c_phase[EVP_MAX_PHASES];
__constant__ GreenConst c_green;
__constant__ ConstParams c_cp;

// structure of c_phase[0] the fast path may compile in: 1 = canonical FCC {111}<110> table (literals), 2 = zero pattern of
// the 24-system HCP table, 0 = none
static int g_table = 0;
void upload_phase_tables(const PhaseDev *ph, int nph) {
  cudaMemcpyToSymbol(c_phase, ph, sizeof(PhaseDev) * nph);
  const bool off = getenv("EVP_K1_FCC") && atoi(getenv("EVP_K1_FCC")) == 0;   // read per upload: tests switch it per solver
  g_table = (off || nph != 1) ? 0 : (fcc_table_matches(ph[0]) ? 1 : (hcp24_pattern_matches(ph[0]) ? 2 : 0));
}
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
                                                      int rowbase, int nrows, SpecLayout Lay, const double2 *__restrict__ twp) {
  using C = XCfg<NX>;
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * C::L;
  const int pair = blockIdx.y;
  const double *a = sig + (long long)(2 * pair) * N + (long long)(rowbase + row0) * NX;   // rowbase: first row of this z-chunk
  const double *b = a + N;
  const int nrl = min(C::L, nrows - row0);
  // coalesced load of L consecutive rows of both fields
  if (nrl == C::L) {   // full tile: compile-time trip count, every load issued before the first use
    constexpr int NIT = C::L * NX / (2 * C::T);
    double2 av[NIT], bv[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = (tid + it * C::T) * 2;
      av[it] = *reinterpret_cast<const double2 *>(a + idx);
      bv[it] = *reinterpret_cast<const double2 *>(b + idx);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = (tid + it * C::T) * 2;
      const int l = idx / NX, x = idx % NX;
      sm[l * C::LS + x] = make_double2(av[it].x, bv[it].x);
      sm[l * C::LS + x + 1] = make_double2(av[it].y, bv[it].y);
    }
  } else {
    for (int idx = tid * 2; idx < nrl * NX; idx += C::T * 2) {
      const int l = idx / NX, x = idx % NX;
      const double2 av = *reinterpret_cast<const double2 *>(a + idx);
      const double2 bv = *reinterpret_cast<const double2 *>(b + idx);
      sm[l * C::LS + x] = make_double2(av.x, bv.x);
      sm[l * C::LS + x + 1] = make_double2(av.y, bv.y);
    }
    for (int idx = nrl * NX + tid; idx < C::L * NX; idx += C::T) sm[(idx / NX) * C::LS + idx % NX] = make_double2(0.0, 0.0);
  }
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
    const long long o = Lay.row_x(2 * pair, row0 + ll, k);
    W[o] = A;
    W[o + Lay.cstride] = B;
  }
}

// ---------------------------------------------------------------------------------------------
// K6: x inverse fused with the strain update  e <- e - de + dE  (row a3)
// ---------------------------------------------------------------------------------------------
template <int NX>
__global__ void __launch_bounds__(XCfg<NX>::T) k_xinv(const double2 *__restrict__ W, double *__restrict__ e, double *__restrict__ de_dbg,
                                                      const MacroDev *__restrict__ macro, long long N, int rowbase, int nrows, SpecLayout Lay,
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
      const long long o = Lay.row_x(2 * pair, row0 + ll, k);
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
  if (e == nullptr) {   // plain output (commit step: local rotation field) into de_dbg
    double *da = de_dbg + (long long)(2 * pair) * N + (long long)(rowbase + row0) * NX;
    for (int idx = tid * 2; idx < nrl * NX; idx += C::T * 2) {
      const int ll = idx / NX, x = idx % NX;
      const double2 z0 = sm[ll * C::LS + x], z1 = sm[ll * C::LS + x + 1];
      *reinterpret_cast<double2 *>(da + idx) = make_double2(z0.x, z1.x);
      *reinterpret_cast<double2 *>(da + N + idx) = make_double2(z0.y, z1.y);
    }
    return;
  }
  const double dEa = macro->dEpend[2 * pair], dEb = macro->dEpend[2 * pair + 1];
  double *ea = e + (long long)(2 * pair) * N + (long long)(rowbase + row0) * NX;
  double *eb = ea + N;
  if (nrl == C::L && !de_dbg) {   // full tile: all e loads in flight before the first use
    constexpr int NIT = C::L * NX / (2 * C::T);
    double2 va[NIT], vb[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = (tid + it * C::T) * 2;
      va[it] = *reinterpret_cast<double2 *>(ea + idx);
      vb[it] = *reinterpret_cast<double2 *>(eb + idx);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = (tid + it * C::T) * 2;
      const int ll = idx / NX, x = idx % NX;
      const double2 z0 = sm[ll * C::LS + x], z1 = sm[ll * C::LS + x + 1];
      va[it].x += dEa - z0.x; va[it].y += dEa - z1.x;
      vb[it].x += dEb - z0.y; vb[it].y += dEb - z1.y;
      *reinterpret_cast<double2 *>(ea + idx) = va[it];
      *reinterpret_cast<double2 *>(eb + idx) = vb[it];
    }
    return;
  }
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
      double *da = de_dbg + (long long)(2 * pair) * N + (long long)(rowbase + row0) * NX;
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

// Tiles move by TMA: one op per (plane, chunk of yc rows); the split (all-to-all send) layout is the
// 5-D tensor [kx][y % nyl][zl][c][y / nyl], the plain layout is the same with nyl = ny.
template <int NY, bool INV>
__global__ void __launch_bounds__(YCfg<NY>::T) k_ypass(const __grid_constant__ PeerMaps tin, const __grid_constant__ PeerMaps tout,
                                                       int lg_nyl_in, int yc_in, int lg_nyl_out, int yc_out, int nzc, int p2p, int pull,
                                                       const double2 *__restrict__ twp) {
  using C = YCfg<NY>;
  extern __shared__ __align__(128) double2 sm[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * C::TX;
  const int z0 = blockIdx.y * C::ZT;
  const int c = blockIdx.z;
  if (tid == 0) {
    tma::mbar_init(&bar, 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  const int nzt = min(C::ZT, nzc - z0);   // planes of this tile that exist in the chunk
  if (tid == 0) {
    tma::mbar_expect_tx(&bar, (uint32_t)(nzt * NY * C::TX * sizeof(double2)));
    for (int zt = 0; zt < nzt; ++zt)
      for (int y0 = 0; y0 < NY; y0 += yc_in) {
        // pull: the rows that rank d transformed along z are read straight out of rank d's buffer over NVLink (TMA load on the
        // peer mapping): the way-back FFT transpose is fused into this kernel's load phase.  Otherwise: local split / plain layout.
        const int d = y0 >> lg_nyl_in;
        tma::load5(sm + (zt * NY + y0) * C::TX, &tin.m[pull ? d : 0], &bar, 2 * k0, y0 & ((1 << lg_nyl_in) - 1), z0 + zt, c, pull ? 0 : d);
      }
  }
  tma::mbar_wait(&bar, 0);
  const int l = tid % C::L, q = tid / C::L;
  block_fft<NY, INV>(sm, q, OffES<C::TX>{(l / C::TX) * (NY * C::TX) + (l % C::TX)}, TwLdg{twp});
  tma::fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    for (int zt = 0; zt < nzt; ++zt)
      for (int y0 = 0; y0 < NY; y0 += yc_out)
      {
          // p2p: the rows of destination rank d go straight into rank d's receive buffer over NVLink (TMA store on the
          // peer mapping): the FFT transpose is fused into this kernel's store phase.  Otherwise: local send layout.
          const int d = y0 >> lg_nyl_out;
          tma::store5(&tout.m[p2p ? d : 0], sm + (zt * NY + y0) * C::TX, 2 * k0, y0 & ((1 << lg_nyl_out) - 1), z0 + zt, c, p2p ? 0 : d);
      }
    tma::commit();
    if (p2p == 1) tma::wait_all0(); else tma::wait_read0();
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
  static constexpr int CS = NZ * TX;                            // component stride in smem (elements)
  static constexpr size_t smem = (size_t)6 * CS * sizeof(double2);
  static constexpr int MINB = (smem <= 100 * 1024) ? 2 : 1;     // two resident blocks when shared memory allows
  static constexpr int TT = (MINB == 2) ? 256 : 512;            // thread target: 128 registers per thread either way
  static constexpr int CG = (TPC >= TT) ? 1 : (2 * TPC >= TT ? 2 : (3 * TPC >= TT ? 3 : 6));
  static constexpr int T = CG * TPC;
};

template <int NZ, int MODE>  // MODE 0: fused fwd+Green+inv; 1: forward only (evp_debug_spectrum); 2: fwd + local-rotation spectrum + inv
__global__ void __launch_bounds__(ZCfg<NZ>::T, ZCfg<NZ>::MINB) k_zfused(const __grid_constant__ ZMaps tz, const __grid_constant__ ZOutMaps tzo,
                                                                        int p2p, int lg_nzl, int lg_nzc, int zc,
                                                                        int ky0, int kx0, int nx, int ny, double rx, double ry, double rz,
                                                                        double scale, const double2 *__restrict__ twp) {
  using C = ZCfg<NZ>;
  extern __shared__ __align__(128) double2 sm[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * C::TX;
  const int yl = blockIdx.y;
  const int nxh = nx / 2 + 1;
  if (tid == 0) {
    tma::mbar_init(&bar, 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  // tile [c][z][TX] by TMA: one op per (component, run of zc planes).  Global z = r*nzl + i*nzc + zz with r the source
  // rank, i the pipeline chunk and zz the plane inside the chunk; each chunk is its own 5-D tensor [kx][yl][zz][c][r]
  if (tid == 0) {
    tma::mbar_expect_tx(&bar, (uint32_t)C::smem);
#pragma unroll 1
    for (int c = 0; c < 6; ++c)
#pragma unroll 1
      for (int z0 = 0; z0 < NZ; z0 += zc)
        tma::load5(sm + c * C::CS + z0 * C::TX, &tz.m[(z0 & ((1 << lg_nzl) - 1)) >> lg_nzc], &bar, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c,
                   z0 >> lg_nzl);
  }
  tma::mbar_wait(&bar, 0);
  const int cg = tid / C::TPC, t = tid % C::TPC;
  const int col = t % C::TX, q = t / C::TX;
#pragma unroll 1
  for (int c = cg; c < 6; c += C::CG) block_fft<NZ, false>(sm, q, OffES<C::TX>{c * C::CS + col}, TwLdg{twp});
  if (MODE == 0 || MODE == 2) {
    // Green operator per frequency (row a2); real and imaginary parts are transformed one after the other
    // (MODE 2, commit step: strain spectrum -> local rotation spectrum in components 0..2, zeros in 3..5)
    const int ky = ky0 + yl;
    const int fy = (ky <= ny / 2) ? ky : ky - ny;
#pragma unroll 1
    for (int idx = tid; idx < C::CS; idx += C::T) {
      const int cc = idx % C::TX, kz = idx / C::TX;
      const int kx = kx0 + k0 + cc;
      if (kx < nxh) {
        const int fz = (kz <= NZ / 2) ? kz : kz - NZ;
        const double x = kx * rx, y = fy * ry, z = fz * rz;
        const bool zero = (kx == 0) && (ky == 0) && (kz == 0);
        const bool nyq = (kx * 2 == nx) || (ky * 2 == ny) || (kz * 2 == NZ);
        double g[6];
        if (MODE == 0 && !nyq && !zero) green_G(c_green, x, y, z, scale, g);
        double *smd = reinterpret_cast<double *>(sm);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          double lam[6], o[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) lam[c] = smd[2 * (c * C::CS + idx) + part];
          if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 6; ++c) o[c] = 0.0;
            if (!zero && !nyq) rot_apply(x, y, z, scale, lam, o);
          } else if (zero) {
#pragma unroll
            for (int c = 0; c < 6; ++c) o[c] = 0.0;
          } else if (nyq) {
            green_nyquist(c_green, scale, lam, o);
          } else {
            green_apply(g, x, y, z, lam, o);
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) smd[2 * (c * C::CS + idx) + part] = o[c];
        }
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int c = cg; c < 6; c += C::CG) block_fft<NZ, true>(sm, q, OffES<C::TX>{c * C::CS + col}, TwLdg{twp});
  }
  tma::fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
#pragma unroll 1
    for (int c = 0; c < 6; ++c)
#pragma unroll 1
      for (int z0 = 0; z0 < NZ; z0 += zc)
      {
        const int r = z0 >> lg_nzl, i = (z0 & ((1 << lg_nzl) - 1)) >> lg_nzc;
        if (p2p)   // planes of rank r go straight into rank r's way-back buffer (peer mapping)
          tma::store5(&tzo.m[r * kMaxChunksP2P + i], sm + c * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c, 0);
        else
          tma::store5(&tz.m[i], sm + c * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c, r);
      }
    tma::commit();
    if (p2p == 1) tma::wait_all0(); else tma::wait_read0();
  }
}

// ---------------------------------------------------------------------------------------------
// K4, production form for nz = 128 / 256: PERSISTENT blocks, tiles double-buffered by TMA, two radix-16
// Stockham passes per direction with the second-pass twiddles hoisted into registers for the whole kernel.
// Per tile: 10 shared-memory round trips instead of 14 and no exposed load/store (ncu on the one-shot kernel:
// 65 % LSU-shared wavefronts, 38 % of the HBM roofline).
// ---------------------------------------------------------------------------------------------
template <int NZ>
struct Z2Cfg {
  static constexpr int TX = (NZ >= 256) ? 4 : 8;
  static constexpr int R1 = NZ / 16;                    // first-pass radix (16 or 8)
  static constexpr int TPC = TX * NZ / 16;              // threads per component (64)
  static constexpr int T = 6 * TPC;                     // NB = 1: all six components at once
  static constexpr int T2 = 3 * TPC;                    // NB = 2: a thread transforms components c and c + 3 in turn
  static constexpr int CS = NZ * TX;
  static constexpr size_t tile = (size_t)6 * CS * sizeof(double2);
  static constexpr size_t smem = 2 * tile;              // NB = 1: two tile buffers
  static constexpr size_t smem2 = tile;                 // NB = 2: one tile buffer
};

// NB = 1: one block per SM, the next tile is prefetched into a second buffer while the current one is transformed.
// NB = 2: two blocks per SM with one tile buffer each; a thread transforms components c and c + 3 one after the other
//         (half the threads, same registers).  The two blocks drift apart, so the shared-memory-bound load/store phases
//         of one overlap the fp64 butterflies, the Green operator and the TMA traffic of the other.
template <int NZ, int NB>
__global__ void __launch_bounds__(NB == 1 ? Z2Cfg<NZ>::T : Z2Cfg<NZ>::T2, NB) k_zfused2(const __grid_constant__ ZMaps tz, const __grid_constant__ ZOutMaps tzo, int p2p,
                                                             int lg_nzl, int lg_nzc, int zc, int ky0, int kx0, int nx,
                                                             int ny, double rx, double ry, double rz, double scale, int nkx, int ntiles,
                                                             const double2 *__restrict__ twp) {
  using C = Z2Cfg<NZ>;
  constexpr int T = (NB == 1) ? C::T : C::T2;
  constexpr int CPT = (NB == 1) ? 1 : 2;                // components per thread
  extern __shared__ __align__(128) double2 sm[];
  __shared__ uint64_t full[2];
  const int tid = threadIdx.x;
  const int nxh = nx / 2 + 1;
  const int c = tid / C::TPC, t = tid % C::TPC;
  const int col = t % C::TX, q = t / C::TX;             // q in [0, NZ/16)
  // hoisted second-pass twiddles: W_NZ^(r*q*NZ/(R1*16)), r = 1..15
  double2 tw[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) tw[r] = __ldg(twp + ((r * q * (NZ / (C::R1 * 16))) & (NZ - 1)));
  if (tid == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  auto issue_load = [&](int tile, int buf) {
    const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
    double2 *dst = sm + (size_t)buf * 6 * C::CS;
    tma::mbar_expect_tx(&full[buf], (uint32_t)C::tile);
#pragma unroll 1
    for (int cc = 0; cc < 6; ++cc)
#pragma unroll 1
      for (int z0 = 0; z0 < NZ; z0 += zc)
        tma::load5(dst + cc * C::CS + z0 * C::TX, &tz.m[(z0 & ((1 << lg_nzl) - 1)) >> lg_nzc], &full[buf], 2 * k0, yl,
                   z0 & ((1 << lg_nzc) - 1), cc, z0 >> lg_nzl);
  };
  int tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) issue_load(tile, 0);
  int n = 0;
  for (; tile < ntiles; tile += gridDim.x, ++n) {
    const int buf = (NB == 1) ? (n & 1) : 0;
    double2 *s = sm + (size_t)buf * 6 * C::CS;
    const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
    if constexpr (NB == 1) {
      // prefetch the next tile into the other buffer (its previous store has been read out: wait_read0 below)
      if (tid == 0 && tile + gridDim.x < ntiles) {
        tma::wait_read0();
        issue_load(tile + gridDim.x, buf ^ 1);
      }
      tma::mbar_wait(&full[buf], (n >> 1) & 1);
    } else {
      tma::mbar_wait(&full[0], n & 1);
    }
    double2 v[16];
    // forward: pass 1 (radix R1, no twiddles), pass 2 (radix 16).  The components of one thread take turns: the barrier
    // between the loads and the stores of one component also orders the other component's stores before its next loads
#pragma unroll
    for (int h = 0; h < CPT; ++h) {
      const OffES<C::TX> off{(c + 3 * h) * C::CS + col};
      pass16_load<NZ, C::R1>(s, q, v, off);
      __syncthreads();
      pass16_store<NZ, C::R1, 1, false>(s, q, v, off, tw);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < CPT; ++h) {
      const OffES<C::TX> off{(c + 3 * h) * C::CS + col};
      pass16_load<NZ, 16>(s, q, v, off);
      __syncthreads();
      pass16_store<NZ, 16, C::R1, false>(s, q, v, off, tw);
    }
    __syncthreads();
    // Green operator per frequency (row a2)
    {
      const int ky = ky0 + yl;
      const int fy = (ky <= ny / 2) ? ky : ky - ny;
#pragma unroll 1
      for (int idx = tid; idx < C::CS; idx += T) {
        const int cc = idx % C::TX, kz = idx / C::TX;
        const int kx = kx0 + k0 + cc;
        if (kx < nxh) {
          const int fz = (kz <= NZ / 2) ? kz : kz - NZ;
          const double x = kx * rx, y = fy * ry, z = fz * rz;
          const bool zero = (kx == 0) && (ky == 0) && (kz == 0);
          const bool nyq = (kx * 2 == nx) || (ky * 2 == ny) || (kz * 2 == NZ);
          double g[6];
          if (!nyq && !zero) green_G(c_green, x, y, z, scale, g);
          double2 l2[6];   // 16-byte accesses: a warp reads 512 contiguous bytes per component
#pragma unroll
          for (int a = 0; a < 6; ++a) l2[a] = s[a * C::CS + idx];
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            double lam[6], o[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) lam[a] = part ? l2[a].y : l2[a].x;
            if (zero) {
#pragma unroll
              for (int a = 0; a < 6; ++a) o[a] = 0.0;
            } else if (nyq) {
              green_nyquist(c_green, scale, lam, o);
            } else {
              green_apply(g, x, y, z, lam, o);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) { if (part) l2[a].y = o[a]; else l2[a].x = o[a]; }
          }
#pragma unroll
          for (int a = 0; a < 6; ++a) s[a * C::CS + idx] = l2[a];
        }
      }
    }
    __syncthreads();
    // inverse
#pragma unroll
    for (int h = 0; h < CPT; ++h) {
      const OffES<C::TX> off{(c + 3 * h) * C::CS + col};
      pass16_load<NZ, C::R1>(s, q, v, off);
      __syncthreads();
      pass16_store<NZ, C::R1, 1, true>(s, q, v, off, tw);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < CPT; ++h) {
      const OffES<C::TX> off{(c + 3 * h) * C::CS + col};
      pass16_load<NZ, 16>(s, q, v, off);
      __syncthreads();
      pass16_store<NZ, 16, C::R1, true>(s, q, v, off, tw);
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
#pragma unroll 1
      for (int cc = 0; cc < 6; ++cc)
#pragma unroll 1
        for (int z0 = 0; z0 < NZ; z0 += zc)
        {
          const int r = z0 >> lg_nzl, i = (z0 & ((1 << lg_nzl) - 1)) >> lg_nzc;
          if (p2p)
            tma::store5(&tzo.m[r * kMaxChunksP2P + i], s + cc * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), cc, 0);
          else
            tma::store5(&tz.m[i], s + cc * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), cc, r);
        }
      tma::commit();
      if constexpr (NB == 2) {
        // single buffer: the next tile can land once the store has read this one out; the other block of the SM
        // computes meanwhile
        if (tile + gridDim.x < ntiles) {
          tma::wait_read0();
          issue_load(tile + gridDim.x, 0);
        }
      }
    }
  }
  if (tid == 0) { if (p2p == 1) tma::wait_all0(); else tma::wait_read0(); }
}

// ---------------------------------------------------------------------------------------------
// K4 for nz = 512 (every multi-GPU bench configuration; 512^3 on one GPU): PERSISTENT, one block per SM, one 192 KB tile
// (6 components x 4 kx x 512 z) whose COMPONENT slots are pipelined individually by TMA:
//   * two compute groups of 256 threads (components 0-2 / 3-5), each transforming its components ONE AFTER THE OTHER with
//     three IN-PLACE radix-8 passes (decimation in frequency forward, in time backward; the Green stage works on the
//     digit-reversed spectrum): a thread reads and writes the same elements, so a component direction needs ONE barrier of
//     the group and one __syncwarp, and the warps run through load / butterfly / store phases unsynchronised — the
//     shared-memory phases of some warps overlap the fp64 phases of others (ncu on the Stockham version of this kernel, which
//     needed five group barriers per component direction: issue slots 34 % busy, shared memory 48 %, fp64 30 %);
//   * rows z and z^1 are kept swapped where bit 3 of z is set between the first and the last pass of a direction, which makes all
//     three passes bank-conflict free on the dense TMA tile (z3_swap);
//   * one producer warp per group: as soon as component c of tile n has been inverse-transformed (mbarrier `done[c]`), its
//     slot is stored (locally or straight into the owning rank's way-back buffer over NVLink) and, once the store has read the
//     slot out, reloaded with component c of tile n+1.  A group needs its components in the order it released them one tile
//     earlier, so every load has about a third of a tile of slack: neither the loads nor the peer stores are on the critical
//     path (the one-shot kernel it replaces waits for both: 1 block/SM, load -> compute -> store -> cp.async.bulk.wait_group).
//   18 warps: 5 on one scheduler -> 96 registers per thread.
// ---------------------------------------------------------------------------------------------
template <int NZ>
struct Z3Cfg {
  static constexpr int TX = 4;
  static constexpr int TPC = TX * NZ / 8;            // threads per component: 8 points each
  static constexpr int TC = 2 * TPC;                 // compute threads (two groups)
  static constexpr int T = TC + 64;                  // + one producer warp per group
  static constexpr int CS = NZ * TX;                 // elements per component slot
  static constexpr size_t tile = (size_t)6 * CS * sizeof(double2);
  static constexpr size_t smem = tile + 16 * sizeof(uint64_t);
};

namespace tma {
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tma
__device__ __forceinline__ void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// read-only load the compiler must not hoist out of a loop (the value would be spilled across the Green stage)
__device__ __forceinline__ double2 ldg_here(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// 512-point FFT of one column IN PLACE: decimation in frequency forward (natural order in, base-8 digit-reversed order out:
// frequency 64c + 8b + a ends at position 64a + 8b + c) and decimation in time backward (digit-reversed in, natural out).
// A thread reads and writes the same eight elements in every pass, so no barrier separates the loads from the stores of a pass:
//   stride 64 (thread u = q)            <- group barrier ->   stride 8 (thread u: block u/8, j = u%8)   <- __syncwarp ->   stride 1
// (the 32 lanes of a warp own one block of 64 points x 4 columns in the last two passes).
// Bank conflicts: the tile is dense [z][4 kx] (64-byte rows, written by TMA).  In the stride-1 pass the lanes of a quarter-warp
// belong to two threads u whose rows are 512 bytes apart: same banks.  Between the first pass and the last the rows z and z^1 are
// therefore stored SWAPPED wherever bit 3 of z is set (position z' = z ^ ((z >> 3) & 1)): odd u then touch the other half of the
// 128-byte bank row, and every pass is conflict free.  The first forward pass reads dense rows and writes swapped ones (the
// target belongs to the neighbouring thread u^1 of the same warp: a __syncwarp orders it), the last inverse pass does the opposite.
// tw64[r] = W^(r u) (stride-64 pass); the stride-8 twiddles W^(8 r (u%8)) are read from the table (eight rows per warp, L1).
__device__ __forceinline__ int z3_swap(int z) { return z ^ ((z >> 3) & 1); }

template <int NZ, int TX, int NT>
__device__ __forceinline__ void z3_forward(double2 *__restrict__ s /* slot + col */, int u, const double2 *__restrict__ twp, const double2 *tw64, int barid) {
  double2 v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = s[(u + 64 * r) * TX];
  bfly8<false>(v);
#pragma unroll
  for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tw64[r]);
  __syncwarp();
  {
    const int us = z3_swap(u);
#pragma unroll
    for (int r = 0; r < 8; ++r) s[(us + 64 * r) * TX] = v[r];
  }
  bar_named(barid, NT);
  {
    const int j = u & 7, base = (u >> 3) * 64 + j;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = s[((base + 8 * r) ^ (r & 1)) * TX];
    bfly8<false>(v);
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], __ldg(twp + 8 * j * r));
#pragma unroll
    for (int r = 0; r < 8; ++r) s[((base + 8 * r) ^ (r & 1)) * TX] = v[r];
  }
  __syncwarp();
  {
    const int o = u & 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = s[((8 * u + r) ^ o) * TX];
    bfly8<false>(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) s[((8 * u + r) ^ o) * TX] = v[r];
  }
}
template <int NZ, int TX, int NT>
__device__ __forceinline__ void z3_inverse(double2 *__restrict__ s, int u, const double2 *__restrict__ twp, const double2 *tw64, int barid) {
  double2 v[8];
  {
    const int o = u & 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = s[((8 * u + r) ^ o) * TX];
    bfly8<true>(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) s[((8 * u + r) ^ o) * TX] = v[r];
  }
  __syncwarp();
  {
    const int j = u & 7, base = (u >> 3) * 64 + j;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = s[((base + 8 * r) ^ (r & 1)) * TX];
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmulc(v[r], __ldg(twp + 8 * j * r));
    bfly8<true>(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) s[((base + 8 * r) ^ (r & 1)) * TX] = v[r];
  }
  bar_named(barid, NT);
  {
    const int us = z3_swap(u);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = s[(us + 64 * r) * TX];
  }
#pragma unroll
  for (int r = 1; r < 8; ++r) v[r] = cmulc(v[r], tw64[r]);
  bfly8<true>(v);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; ++r) s[(u + 64 * r) * TX] = v[r];
}

template <int NZ>
__global__ void __launch_bounds__(Z3Cfg<NZ>::T, 1) k_zfused3(const __grid_constant__ ZMaps tz, const __grid_constant__ ZOutMaps tzo, int p2p,
                                                             int lg_nzl, int lg_nzc, int zc, int ky0, int kx0, int nx, int ny, double rx, double ry,
                                                             double rz, double scale, int nkx, int ntiles, const double2 *__restrict__ twp) {
  using C = Z3Cfg<NZ>;
  static_assert(NZ == 512, "three radix-8 passes");
  extern __shared__ __align__(128) double2 sm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(sm + 6 * C::CS);   // full[c]: component c of the current tile has landed
  uint64_t *done = full + 6;                                        // done[c]: component c has been transformed back (TPC arrivals)
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      tma::mbar_init(&full[c], 1);
      tma::mbar_init(&done[c], C::TPC);
    }
    tma::fence_mbar_init();
  }
  __syncthreads();
  if (tid >= C::TC) {
    // ---- producer warp of group g: stores and loads of components 3g .. 3g+2 ----
    if ((tid & 31) != 0) return;
    const int g = (tid - C::TC) >> 5;
    auto issue_load = [&](int tile, int c) {
      const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
      tma::mbar_expect_tx(&full[c], (uint32_t)(C::CS * sizeof(double2)));
#pragma unroll 1
      for (int z0 = 0; z0 < NZ; z0 += zc)
        tma::load5(sm + c * C::CS + z0 * C::TX, &tz.m[(z0 & ((1 << lg_nzl) - 1)) >> lg_nzc], &full[c], 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c,
                   z0 >> lg_nzl);
    };
    int tile = blockIdx.x;
    if (tile < ntiles)
      for (int h = 0; h < 3; ++h) issue_load(tile, 3 * g + h);
    for (int n = 0; tile < ntiles; tile += gridDim.x, ++n) {
      const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
      const bool more = tile + (int)gridDim.x < ntiles;
#pragma unroll 1
      for (int h = 0; h < 3; ++h) {
        const int c = 3 * g + h;
        tma::mbar_wait(&done[c], n & 1);
#pragma unroll 1
        for (int z0 = 0; z0 < NZ; z0 += zc) {
          const int r = z0 >> lg_nzl, i = (z0 & ((1 << lg_nzl) - 1)) >> lg_nzc;
          if (p2p)   // planes of rank r go straight into rank r's way-back buffer (peer mapping): the transpose is this store
            tma::store5(&tzo.m[r * kMaxChunksP2P + i], sm + c * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c, 0);
          else
            tma::store5(&tz.m[i], sm + c * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c, r);
        }
        tma::commit();
        if (more) {
          tma::wait_read0();                 // the slot has been read out: it can take component c of the next tile
          issue_load(tile + gridDim.x, c);
        }
      }
    }
    if (p2p == 1) tma::wait_all0(); else tma::wait_read0();
    return;
  }
  // ---- compute groups ----
  const int g = tid / C::TPC, t = tid % C::TPC;
  const int col = t % C::TX, q = t / C::TX;          // q in [0, NZ/8)
  const int nxh = nx / 2 + 1;
  int n = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
    const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
    double2 tw3[8];                                   // stride-64 twiddles W^(r q): in registers across the three components of a
                                                      // direction, re-read (L1) after the Green stage rather than spilled across it
#pragma unroll
    for (int r = 1; r < 8; ++r) tw3[r] = ldg_here(twp + ((r * q) & (NZ - 1)));
    tw3[0] = make_double2(1.0, 0.0);
#pragma unroll 1
    for (int h = 0; h < 3; ++h) {
      const int c = 3 * g + h;
      tma::mbar_wait(&full[c], n & 1);
      z3_forward<NZ, C::TX, C::TPC>(sm + c * C::CS + col, q, twp, tw3, 2 + g);
    }
    bar_named(1, C::TC);
    // Green operator per frequency (row a2)
    {
      const int ky = ky0 + yl;
      const int fy = (ky <= ny / 2) ? ky : ky - ny;
#pragma unroll 1
      for (int idx = tid; idx < C::CS; idx += C::TC) {
        // slot index -> (row, column) -> position (rows with bit 3 set are stored pairwise swapped) -> frequency (the forward
        // transform leaves the spectrum in base-8 digit-reversed order)
        const int cc = idx % C::TX, pos = z3_swap(idx / C::TX);
        const int kz = ((pos & 7) << 6) | (pos & 0x38) | (pos >> 6);
        const int kx = kx0 + k0 + cc;
        if (kx < nxh) {
          const int fz = (kz <= NZ / 2) ? kz : kz - NZ;
          const double x = kx * rx, y = fy * ry, z = fz * rz;
          const bool zero = (kx == 0) && (ky == 0) && (kz == 0);
          const bool nyq = (kx * 2 == nx) || (ky * 2 == ny) || (kz * 2 == NZ);
          double gg[6];
          if (!nyq && !zero) green_G(c_green, x, y, z, scale, gg);
          double2 l2[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) l2[a] = sm[a * C::CS + idx];
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            double lam[6], o[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) lam[a] = part ? l2[a].y : l2[a].x;
            if (zero) {
#pragma unroll
              for (int a = 0; a < 6; ++a) o[a] = 0.0;
            } else if (nyq) {
              green_nyquist(c_green, scale, lam, o);
            } else {
              green_apply(gg, x, y, z, lam, o);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) { if (part) l2[a].y = o[a]; else l2[a].x = o[a]; }
          }
#pragma unroll
          for (int a = 0; a < 6; ++a) sm[a * C::CS + idx] = l2[a];
        }
      }
    }
    bar_named(1, C::TC);
#pragma unroll
    for (int r = 1; r < 8; ++r) tw3[r] = ldg_here(twp + ((r * q) & (NZ - 1)));
#pragma unroll 1
    for (int h = 0; h < 3; ++h) {
      const int c = 3 * g + h;
      z3_inverse<NZ, C::TX, C::TPC>(sm + c * C::CS + col, q, twp, tw3, 2 + g);
      tma::fence_proxy_async();          // this thread's generic-proxy writes, before the producer's TMA store reads them
      tma::mbar_arrive(&done[c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4 for nz = 256, round-2 variant (k_zfused4, opt-in with EVP_Z4=1; not faster than k_zfused2, see launch_zfused): the design of
// k_zfused3 on the 96 KB tile of nz = 256, two blocks per SM.
//   * per block: three compute groups of 64 threads (group g transforms components g and g + 3, 16 points per thread) and one
//     producer warp that stores / reloads every component slot as soon as its group has released it (k_zfused2 waits for the
//     whole 96 KB tile after its last store: ncu, 23 % of the samples on that mbarrier);
//   * IN-PLACE radix-16 passes: stride 16 with twiddles, then stride 1 (decimation in frequency; the spectrum is left in base-16
//     digit-reversed order, frequency k1 + 16 k2 at position 16 k1 + k2, which the Green stage undoes), and the mirror image
//     backward.  One barrier of the 64-thread group per component direction instead of four block barriers per pass pair;
//   * rows z and z^1 are kept swapped where bit 4 of z is set between the first and the last pass (z4_swap): the stride-1 pass
//     would otherwise hit every bank twice (two threads of a quarter-warp 1 KB apart).
// ---------------------------------------------------------------------------------------------
template <int NZ>
struct Z4Cfg {
  static constexpr int TX = 4;
  static constexpr int TPC = TX * NZ / 16;           // 64 threads transform one component (16 points each)
  static constexpr int NG = 3;                       // compute groups; group g owns components g and g + 3
  static constexpr int TC = NG * TPC;                // 192 compute threads
  static constexpr int T = TC + 32;                  // + one producer warp
  static constexpr int CS = NZ * TX;
  static constexpr size_t tile = (size_t)6 * CS * sizeof(double2);
  static constexpr size_t smem = tile + 16 * sizeof(uint64_t);
};
__device__ __forceinline__ int z4_swap(int z) { return z ^ ((z >> 4) & 1); }
__device__ __forceinline__ constexpr int z4_xr(int p) { return 4 * (p & 3) + (p >> 2); }   // bfly16 leaves X[z4_xr(p)] in v[p]

template <int NZ, int TX, int NT>
__device__ __forceinline__ void z4_forward(double2 *__restrict__ s /* slot + col */, int u, const double2 *__restrict__ twp, int barid) {
  double2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = s[(u + 16 * r) * TX];
  bfly16<false>(v);
#pragma unroll
  for (int p = 1; p < 16; ++p) v[p] = cmul(v[p], __ldg(twp + u * z4_xr(p)));   // W_256^(u k1), k1 = z4_xr(p)
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 16; ++p) s[z4_swap(u + 16 * z4_xr(p)) * TX] = v[p];
  bar_named(barid, NT);
  {
    const int o = u & 1;     // bit 4 of 16 u + r
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = s[((16 * u + r) ^ o) * TX];
    bfly16<false>(v);
#pragma unroll
    for (int p = 0; p < 16; ++p) s[((16 * u + z4_xr(p)) ^ o) * TX] = v[p];
  }
}
template <int NZ, int TX, int NT>
__device__ __forceinline__ void z4_inverse(double2 *__restrict__ s, int u, const double2 *__restrict__ twp, int barid) {
  double2 v[16];
  {
    const int o = u & 1;
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = s[((16 * u + r) ^ o) * TX];
    bfly16<true>(v);
#pragma unroll
    for (int p = 0; p < 16; ++p) s[((16 * u + z4_xr(p)) ^ o) * TX] = v[p];
  }
  bar_named(barid, NT);
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = s[z4_swap(u + 16 * r) * TX];
#pragma unroll
  for (int r = 1; r < 16; ++r) v[r] = cmulc(v[r], __ldg(twp + u * r));
  bfly16<true>(v);
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 16; ++p) s[(u + 16 * z4_xr(p)) * TX] = v[p];
}

template <int NZ>
__global__ void __launch_bounds__(Z4Cfg<NZ>::T, 2) k_zfused4(const __grid_constant__ ZMaps tz, const __grid_constant__ ZOutMaps tzo, int p2p,
                                                             int lg_nzl, int lg_nzc, int zc, int ky0, int kx0, int nx, int ny, double rx, double ry,
                                                             double rz, double scale, int nkx, int ntiles, const double2 *__restrict__ twp) {
  using C = Z4Cfg<NZ>;
  static_assert(NZ == 256, "two radix-16 passes");
  extern __shared__ __align__(128) double2 sm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(sm + 6 * C::CS);   // full[c]: component c of the current tile has landed
  uint64_t *done = full + 6;                                        // done[c]: component c has been transformed back (TPC arrivals)
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      tma::mbar_init(&full[c], 1);
      tma::mbar_init(&done[c], C::TPC);
    }
    tma::fence_mbar_init();
  }
  __syncthreads();
  if (tid >= C::TC) {
    // ---- producer warp: stores and loads of all six component slots, in the order the groups release them ----
    if ((tid & 31) != 0) return;
    auto issue_load = [&](int tile, int c) {
      const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
      tma::mbar_expect_tx(&full[c], (uint32_t)(C::CS * sizeof(double2)));
#pragma unroll 1
      for (int z0 = 0; z0 < NZ; z0 += zc)
        tma::load5(sm + c * C::CS + z0 * C::TX, &tz.m[(z0 & ((1 << lg_nzl) - 1)) >> lg_nzc], &full[c], 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c,
                   z0 >> lg_nzl);
    };
    int tile = blockIdx.x;
    if (tile < ntiles)
      for (int c = 0; c < 6; ++c) issue_load(tile, c);
    for (int n = 0; tile < ntiles; tile += gridDim.x, ++n) {
      const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
      const bool more = tile + (int)gridDim.x < ntiles;
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {
        tma::mbar_wait(&done[c], n & 1);
#pragma unroll 1
        for (int z0 = 0; z0 < NZ; z0 += zc) {
          const int r = z0 >> lg_nzl, i = (z0 & ((1 << lg_nzl) - 1)) >> lg_nzc;
          if (p2p)
            tma::store5(&tzo.m[r * kMaxChunksP2P + i], sm + c * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c, 0);
          else
            tma::store5(&tz.m[i], sm + c * C::CS + z0 * C::TX, 2 * k0, yl, z0 & ((1 << lg_nzc) - 1), c, r);
        }
        tma::commit();
        if (more) {
          tma::wait_read0();
          issue_load(tile + gridDim.x, c);
        }
      }
    }
    if (p2p == 1) tma::wait_all0(); else tma::wait_read0();
    return;
  }
  // ---- compute groups ----
  const int g = tid / C::TPC, t = tid % C::TPC;
  const int col = t % C::TX, u = t / C::TX;          // u in [0, 16)
  const int nxh = nx / 2 + 1;
  int n = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
    const int k0 = (tile % nkx) * C::TX, yl = tile / nkx;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int c = g + 3 * h;
      tma::mbar_wait(&full[c], n & 1);
      z4_forward<NZ, C::TX, C::TPC>(sm + c * C::CS + col, u, twp, 2 + g);
    }
    bar_named(1, C::TC);
    {
      const int ky = ky0 + yl;
      const int fy = (ky <= ny / 2) ? ky : ky - ny;
#pragma unroll 1
      for (int idx = tid; idx < C::CS; idx += C::TC) {
        // slot index -> (row, column) -> position (rows with bit 4 set are stored pairwise swapped) -> frequency (base-16 digits reversed)
        const int cc = idx % C::TX, pos = z4_swap(idx / C::TX);
        const int kz = ((pos & 15) << 4) | (pos >> 4);
        const int kx = kx0 + k0 + cc;
        if (kx < nxh) {
          const int fz = (kz <= NZ / 2) ? kz : kz - NZ;
          const double x = kx * rx, y = fy * ry, z = fz * rz;
          const bool zero = (kx == 0) && (ky == 0) && (kz == 0);
          const bool nyq = (kx * 2 == nx) || (ky * 2 == ny) || (kz * 2 == NZ);
          double gg[6];
          if (!nyq && !zero) green_G(c_green, x, y, z, scale, gg);
          double2 l2[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) l2[a] = sm[a * C::CS + idx];
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            double lam[6], o[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) lam[a] = part ? l2[a].y : l2[a].x;
            if (zero) {
#pragma unroll
              for (int a = 0; a < 6; ++a) o[a] = 0.0;
            } else if (nyq) {
              green_nyquist(c_green, scale, lam, o);
            } else {
              green_apply(gg, x, y, z, lam, o);
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) { if (part) l2[a].y = o[a]; else l2[a].x = o[a]; }
          }
#pragma unroll
          for (int a = 0; a < 6; ++a) sm[a * C::CS + idx] = l2[a];
        }
      }
    }
    bar_named(1, C::TC);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int c = g + 3 * h;
      z4_inverse<NZ, C::TX, C::TPC>(sm + c * C::CS + col, u, twp, 2 + g);
      tma::fence_proxy_async();
      tma::mbar_arrive(&done[c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1: constitutive update (rows a4, a5, a6).  One thread per voxel, 128 threads per block.
// ---------------------------------------------------------------------------------------------
constexpr int kCB = 128;
constexpr int kNSum = 11;   // per-voxel sums reduced over the grid (+ one max): partial slot layout [kNSum + 1][warps]
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

// Recursive-halving butterfly: sums N per-lane values over the warp in N + O(log N) exchanges instead of 5 N.  At every
// level the lanes of the upper half hand their lower slots to the partner and keep the upper ones (and vice versa), so
// the number of live slots halves; from one slot on the exchange is a plain xor sum.  The summation tree is fixed.
template <int N, int M>
struct WarpHalve {
  static __device__ __forceinline__ void run(double *v, int lane, int &base, int &cnt, int &dup) {
    if constexpr (M > 0) {
      if constexpr (N == 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], M);
        dup |= M;
        WarpHalve<1, M / 2>::run(v, lane, base, cnt, dup);
      } else {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const double lo = v[i];
          const double hi = (H + i < N) ? v[H + i] : 0.0;
          const double send = up ? lo : hi, keep = up ? hi : lo;
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
        }
        if (up) { base += H; cnt -= H; } else { cnt = min(cnt, H); }
        WarpHalve<H, M / 2>::run(v, lane, base, cnt, dup);
      }
    }
  }
};

// per-warp partial sums (no block barrier): NV sums and one max go into the SoA partial buffer
// partials[k * nw + warp]; sum k is stored by the lane that ends up owning it in the butterfly
template <int NV>
__device__ __forceinline__ void warp_partials_store(const double vals[NV], int vmax, double *__restrict__ partials, long long nw,
                                                    long long gw0 = 0) {
  const int lane = threadIdx.x & 31;
  const long long gw = gw0 + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = vals[k];
  int base = 0, cnt = NV, dup = 0;
  WarpHalve<NV, 16>::run(v, lane, base, cnt, dup);
  if (cnt >= 1 && (lane & dup) == 0) partials[(long long)base * nw + gw] = v[0];
  const int m = warp_max(vmax);
  if (lane == 0) partials[(long long)NV * nw + gw] = (double)m;
}

struct RegAcc25 {   // register-resident 5x5 rotation (indices are compile-time constants after unrolling)
  const double *m;
  __device__ __forceinline__ double operator()(int k) const { return m[k]; }
};
struct SmAcc {      // strided per-thread array in shared memory: element k of this thread
  double *p;
  __device__ __forceinline__ double operator()(int k) const { return p[k * kCB]; }
  __device__ __forceinline__ void operator()(int k, double v) const { p[k * kCB] = v; }
};

// once per increment: orientation-dependent invariants (M, Jb) of every ORIENTATION CLASS and 1/tau_c of
// every voxel.  Orientation classes are grains while the texture has not evolved per voxel, voxels after.
__global__ void __launch_bounds__(kCB) k_prep_orient(Fields f, int fast) {
  const long long o = (long long)blockIdx.x * kCB + threadIdx.x;
  const long long NO = f.norient, N = f.N;
  if (o >= NO) return;
  const long long v = f.orient_rep[o];   // a voxel that carries this orientation (-1: class absent on this rank)
  if (v < 0) return;
  const PhaseDev &P = c_phase[f.phase[v]];
  double R[9], M[25], Jb[21];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = f.rot[k * N + v];
  increment_invariants(P, c_cp, R, M, Jb);
  if (fast) jb_eliminate_hydrostatic(Jb);   // fast path: table [K' | b | d] (5x5 Newton, evp_core.h)
#pragma unroll
  for (int k = 0; k < 25; ++k) f.mrot[k * NO + o] = M[k];
#pragma unroll
  for (int k = 0; k < 21; ++k) f.jb[k * NO + o] = Jb[k];
}
// fast_npow < 0: itc = 1/tau_c (generic kernels);  >= 0: itc = dt*gamma0*n / tau_c^n (uniform-exponent fast path, n = fast_npow + 1)
__global__ void __launch_bounds__(256) k_prep_itc(Fields f, int nsmax, int fast_npow) {
  const long long n = (long long)nsmax * f.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    f.itc[i] = (fast_npow >= 0) ? rate_factor(c_cp.dtg0n[i / f.N], f.crss[i], fast_npow) : 1.0 / f.crss[i];
}

// K1.  NS_T > 0: unrolled system loop; NPOW_T >= 0: compile-time rate exponent; ONEPH: single phase
// (tables addressed as c_phase[0], i.e. immediate constant-bank operands).
template <int NS_T, int NPOW_T, bool ONEPH, int MINB>
__global__ void __launch_bounds__(kCB, MINB) k_constitutive_t(Fields f, long long vbase, long long count, double *__restrict__ partials,
                                                              long long nw, long long gw0, int pf_dist) {
  extern __shared__ double smd[];  // [21 Jb | 6 g | 6 s_old | nsmax 1/tau_c] x kCB
  const int tid = threadIdx.x;
  const long long vl = (long long)blockIdx.x * kCB + tid;   // index inside this z-chunk
  const long long v = vbase + vl;
  const long long N = f.N;
  double vals[kNSum] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // ds, de, sig[6], nit, nonfinite, unconverged
  int nit = 0;
  // L2 prefetch of the per-voxel streams of the block that runs one residency wave later: with ~12 warps per SM
  // the loads below would otherwise expose DRAM latency (ncu: long-scoreboard stalls on their first use)
  {
    const long long vp = vbase + ((long long)blockIdx.x + pf_dist) * kCB;
    if (vp + kCB <= N) {
      const int nstream = 18 + ((NS_T > 0) ? NS_T : 0);
      if (tid < nstream) {
        const double *base = (tid < 6) ? f.sig + (long long)tid * N : (tid < 12) ? f.e + (long long)(tid - 6) * N
                             : (tid < 18) ? f.epsp + (long long)(tid - 12) * N : f.itc + (long long)(tid - 18) * N;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + vp), "r"(kCB * 8) : "memory");
      } else if (tid == nstream) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(f.orient + vp), "r"(kCB * 4) : "memory");
      }
    }
  }
  if (vl < count) {
    const PhaseDev &P = ONEPH ? c_phase[0] : c_phase[f.phase[v]];
    const SmAcc jb{smd + tid}, gv{smd + 21 * kCB + tid}, so{smd + 27 * kCB + tid}, itc{smd + 33 * kCB + tid};
    const long long NO = f.norient;
    const long long oid = f.orient[v];   // orientation class: grain while the texture is per grain, voxel afterwards
    double sc[6];
    {
      // every load of the voxel is issued before the first use (39+ independent loads in flight)
      double M[25], sig[6], em[6], ep[6], jbv[21];
#pragma unroll
      for (int k = 0; k < 25; ++k) M[k] = __ldg(f.mrot + k * NO + oid);
#pragma unroll
      for (int c = 0; c < 6; ++c) sig[c] = f.sig[c * N + v];
#pragma unroll
      for (int c = 0; c < 6; ++c) em[c] = f.e[c * N + v];
#pragma unroll
      for (int c = 0; c < 6; ++c) ep[c] = __ldg(f.epsp + c * N + v);
#pragma unroll
      for (int k = 0; k < 21; ++k) jbv[k] = __ldg(f.jb + k * NO + oid);
      const int ns = (NS_T > 0) ? NS_T : P.nsys;
      if (NS_T > 0) {
        double tc[NS_T > 0 ? NS_T : 1];
#pragma unroll
        for (int s = 0; s < NS_T; ++s) tc[s] = __ldg(f.itc + (long long)s * N + v);
#pragma unroll
        for (int s = 0; s < NS_T; ++s) itc(s, tc[s]);
      } else {
        for (int s = 0; s < ns; ++s) itc(s, __ldg(f.itc + (long long)s * N + v));
      }
#pragma unroll
      for (int k = 0; k < 21; ++k) jb(k, jbv[k]);
#pragma unroll
      for (int c = 0; c < 6; ++c) em[c] -= ep[c];
      constitutive_prep(c_cp, RegAcc25{M}, sig, em, gv, so, sc);
    }
    int bad = 0;
    nit = newton_crystal_t<NS_T, NPOW_T>(P, jb, gv, sc, c_cp.dt, c_cp.tol_newton, c_cp.newton_itmax, itc, &bad);
    double M[25], sig[6], ds, de;
#pragma unroll
    for (int k = 0; k < 25; ++k) M[k] = __ldg(f.mrot + k * NO + oid);   // second touch: L1/L2 hit
    constitutive_finish(P, RegAcc25{M}, sc, jb, so, sig, &ds, &de);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      f.sig[c * N + v] = sig[c];
      vals[2 + c] = sig[c];
    }
    vals[0] = ds;
    vals[1] = de;
    vals[8] = (double)nit;
    vals[9] = (double)(bad & 1);
    vals[10] = (double)(bad >> 1);
  }
  warp_partials_store<kNSum>(vals, nit, partials, nw, gw0);
}


// K1, uniform-exponent fast path (one phase, NS_T systems, integer exponent NPOW_T + 1 for all of them).
//  * the 18 + NS_T per-voxel streams (sig, e, eps_p, 1/tau_c) of the block are staged by 1 KB bulk copies
//    (cp.async.bulk, one stream per thread, mbarrier) while every thread gathers its orientation-class tables:
//    no register is tied up by loads in flight, which is what allows MINB = 4 resident blocks without spills;
//  * the staging area of sig/e/eps_p is reused for g and s_old once the thread has consumed its column;
//  * Newton: newton_crystal_p (evp_core.h); TAB: structure of the Schmid tables known at compile time (FCC: literal
//    values; HCP-24: zero pattern) — zero terms are never issued.
// Shared memory (doubles x kCB): [21 Jb | 18 streams -> 6 g, 6 s_old | NS_T rate factors] + mbarrier.
// f.itc holds the rate factors dt*gamma0*n/tau_c^n here (k_prep_itc with fast_npow >= 0).
template <int NS_T, int NPOW_T, bool TWIN, int MINB, int G, int TAB>
__global__ void __launch_bounds__(kCB, MINB) k_constitutive_p(Fields f, long long vbase, long long count, double *__restrict__ partials,
                                                              long long nw, long long gw0, int pf_dist) {
  extern __shared__ __align__(16) double smd[];
  constexpr int NSTREAM = 18 + NS_T;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smd + (21 + NSTREAM) * kCB);
  const int tid = threadIdx.x;
  const long long vl = (long long)blockIdx.x * kCB + tid;
  const long long v = vbase + vl;
  const long long N = f.N;
  const long long v0 = vbase + (long long)blockIdx.x * kCB;
  // pf_dist < 0 (EVP_K1_BULK=0, tests): per-thread loads for every block, the path partial / unaligned blocks take
  const bool bulk = pf_dist >= 0 && ((long long)(blockIdx.x + 1) * kCB <= count) && ((v0 & 1) == 0) && ((N & 1) == 0);
  double *st = smd + 21 * kCB;   // stream s of the block at st + s*kCB
  const bool active = vl < count;
  long long oid = 0;
  if (active) oid = f.orient[v];   // head of the only dependent load chain (class id -> class tables): issue it first
  if (bulk) {
    if (tid == 0) {
      tma::mbar_init(bar, 1);
      tma::fence_mbar_init();
    }
    __syncthreads();   // barrier initialised before any copy can complete on it / any thread waits on it
    // armed by lane 0 of warp 0 ahead of that warp's copies (program order); the single arrival keeps the phase open
    if (tid == 0) tma::mbar_expect_tx(bar, NSTREAM * kCB * 8);
    if (tid < NSTREAM) {   // one 1 KB stream per thread
      const int s = tid;
      const double *src = (s < 6) ? f.sig + (long long)s * N : (s < 12) ? f.e + (long long)(s - 6) * N
                          : (s < 18) ? f.epsp + (long long)(s - 12) * N : f.itc + (long long)(s - 18) * N;
      tma::load1(st + s * kCB, src + v0, kCB * 8, bar);
    }
  } else if (vl < count) {
#pragma unroll
    for (int c = 0; c < 6; ++c) st[c * kCB + tid] = f.sig[c * N + v];
#pragma unroll
    for (int c = 0; c < 6; ++c) st[(6 + c) * kCB + tid] = f.e[c * N + v];
#pragma unroll
    for (int c = 0; c < 6; ++c) st[(12 + c) * kCB + tid] = f.epsp[c * N + v];
#pragma unroll
    for (int s = 0; s < NS_T; ++s) st[(18 + s) * kCB + tid] = f.itc[(long long)s * N + v];
  }
  // L2 prefetch of the streams of the block that runs one residency wave later
  {
    const long long vp = vbase + ((long long)blockIdx.x + pf_dist) * kCB;
    if (pf_dist >= 0 && vp + kCB <= N) {
      if (tid >= 32 && tid < 32 + NSTREAM) {
        const int s = tid - 32;
        const double *base = (s < 6) ? f.sig + (long long)s * N : (s < 12) ? f.e + (long long)(s - 6) * N
                             : (s < 18) ? f.epsp + (long long)(s - 12) * N : f.itc + (long long)(s - 18) * N;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + vp), "r"(kCB * 8) : "memory");
      } else if (tid == 32 + NSTREAM) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(f.orient + vp), "r"(kCB * 4) : "memory");
      }
    }
  }
  double vals[kNSum] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // ds, de, sig[6], nit, nonfinite, unconverged
  int nit = 0;
  const PhaseDev &P = c_phase[0];
  const SmAcc jb{smd + tid}, gv{st + tid}, so{st + 6 * kCB + tid}, itc{st + 18 * kCB + tid};
  const long long NO = f.norient;
  double sc[6];
  double M[25];
  if (active) {
    double jbv[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) jbv[k] = __ldg(f.jb + k * NO + oid);
#pragma unroll
    for (int k = 0; k < 25; ++k) M[k] = __ldg(f.mrot + k * NO + oid);
#pragma unroll
    for (int k = 0; k < 21; ++k) jb(k, jbv[k]);
  }
  if (bulk) tma::mbar_wait(bar, 0);
  if (active) {
    double sig[6], em[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) sig[c] = st[c * kCB + tid];
#pragma unroll
    for (int c = 0; c < 6; ++c) em[c] = st[(6 + c) * kCB + tid] - st[(12 + c) * kCB + tid];
    constitutive_prep(c_cp, RegAcc25{M}, sig, em, gv, so, sc);   // writes g / s_old over this thread's sig / e column
    int bad = 0;
    nit = newton_crystal_p<NS_T, NPOW_T, TWIN, G, TAB>(P, c_cp, jb, gv, sc, itc, &bad);
    double ds, de;
#pragma unroll
    for (int k = 0; k < 25; ++k) M[k] = __ldg(f.mrot + k * NO + oid);   // second touch: L1/L2 hit
    constitutive_finish_p(P, RegAcc25{M}, sc, jb, so, sig, &ds, &de);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      f.sig[c * N + v] = sig[c];
      vals[2 + c] = sig[c];
    }
    vals[0] = ds;
    vals[1] = de;
    vals[8] = (double)nit;
    vals[9] = (double)(bad & 1);
    vals[10] = (double)(bad >> 1);
  }
  warp_partials_store<kNSum>(vals, nit, partials, nw, gw0);
}

// second stage of the reductions: fixed-order two-level sum of the warp partials (deterministic)
constexpr int kRedBlocks = 296;
__global__ void __launch_bounds__(256) k_reduce1(const double *__restrict__ partials, long long nw, double *__restrict__ scratch) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const long long chunk = (nw + gridDim.x - 1) / gridDim.x;
  const long long w0 = (long long)blockIdx.x * chunk, w1 = (w0 + chunk < nw) ? w0 + chunk : nw;
  for (int k = 0; k <= kNSum; ++k) {
    double acc = 0.0;
    for (long long w = w0 + tid; w < w1; w += 256) acc = (k < kNSum) ? acc + partials[(long long)k * nw + w] : fmax(acc, partials[(long long)k * nw + w]);
    red[tid] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (tid < s) red[tid] = (k < kNSum) ? red[tid] + red[tid + s] : fmax(red[tid], red[tid + s]);
      __syncthreads();
    }
    if (tid == 0) scratch[(long long)blockIdx.x * 16 + k] = red[0];
    __syncthreads();
  }
}
// one warp per quantity: lanes stride over the stage-1 partials, fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(32 * (kNSum + 1)) k_reduce2(const double *__restrict__ scratch, int nb, double *__restrict__ totals) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int b = lane; b < nb; b += 32) acc = (k < kNSum) ? acc + scratch[(long long)b * 16 + k] : fmax(acc, scratch[(long long)b * 16 + k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, acc, o);
    acc = (k < kNSum) ? acc + other : fmax(acc, other);
  }
  if (lane == 0) totals[k] = acc;
}

// rows a6 (normalisation) + a7 (macro strain correction) on the device.
// totals: [0] sum|dsig| [1] sum|S0 dsig| [2..7] sum sig [8] sum nit [9] non-finite [10] unconverged [11] max nit
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
  m->newton_max = (int)totals[kNSum];
  m->nonfinite = (int)totals[9];
  m->unconverged = (long long)totals[10];
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

struct CommitParams {
  double dt;
  double wapp[3];       // applied (macroscopic) spin, axial (w32,w13,w21)
  int texture, twinning;
};

// per-increment commit (§8(f).1): eps_p += dt*edp(sig); extended Voce; twin fractions; lattice rotation
//   dR = exp(dt*W_app + (w_local_new - w_local_old) - dt*W_plastic) with w_local from the FFT of the compatible strain
__global__ void __launch_bounds__(kCB) k_commit(Fields f, CommitParams cp, double *__restrict__ partials, long long nw) {
  extern __shared__ double dg_sm[];  // [nsmax][kCB] |dgamma|
  const int tid = threadIdx.x;
  const long long v = (long long)blockIdx.x * kCB + tid;
  const long long N = f.N;
  const double dt = cp.dt;
  double sums[kNSum] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [0] sum of current twin fractions, [1] twin fraction added by this increment, [2..7] eps_p
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
    double edc[5] = {0, 0, 0, 0, 0}, dG = 0.0, wpc[3] = {0, 0, 0}, fsum = 0.0, dfsum = 0.0;
    const int ns = P.nsys;
    for (int s = 0; s < ns; ++s) {
      double tau = 0.0;
#pragma unroll
      for (int c = 0; c < 5; ++c) tau += P.m[s][c] * sc[c];
      double gd, dgd;
      slip_rate(P, s, tau, 1.0 / f.crss[(long long)s * N + v], gd, dgd);
#pragma unroll
      for (int c = 0; c < 5; ++c) edc[c] += gd * P.m[s][c];
#pragma unroll
      for (int k = 0; k < 3; ++k) wpc[k] += P.alpha[s][k] * gd;
      const double dg = fabs(gd) * dt;
      dg_sm[s * kCB + tid] = dg;
      dG += dg;
      if (cp.twinning && P.twin[s]) {
        const double df = gd * dt * P.itshear[s];   // also for voxels already reoriented: F_acc is a history sum
        const double fnew = f.twinf[(long long)s * N + v] + df;
        f.twinf[(long long)s * N + v] = fnew;
        fsum += fnew;
        dfsum += df;
      }
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
    sums[0] = fsum;
    sums[1] = dfsum;
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
    if (cp.texture) {
      double dw[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double wps = R[3 * k] * wpc[0] + R[3 * k + 1] * wpc[1] + R[3 * k + 2] * wpc[2];   // plastic spin, sample frame
        const double wn = f.de[k * N + v];                                                         // new local rotation (FFT)
        dw[k] = dt * cp.wapp[k] + (wn - f.wrot[k * N + v]) - dt * wps;
        f.wrot[k * N + v] = wn;
      }
      rotate_lattice(R, dw);
#pragma unroll
      for (int k = 0; k < 9; ++k) f.rot[k * N + v] = R[k];
    }
  }
  warp_partials_store<kNSum>(sums, 0, partials, nw);
}

// PTR (Tome, Lebensohn, Kocks 1991): a voxel whose predominant twin system exceeds thr1 + thr2*ratio (ratio = F_eff/F_acc)
// takes the twin orientation R (2 n n^T - I); counts go to partial slot 0
__global__ void __launch_bounds__(kCB) k_twin_reorient(Fields f, double ratio, double *__restrict__ partials, long long nw) {
  const long long v = (long long)blockIdx.x * kCB + threadIdx.x;
  const long long N = f.N;
  double sums[kNSum] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (v < N && !f.twinned[v]) {
    const PhaseDev &P = c_phase[f.phase[v]];
    const double thr = P.twin_thr1 + P.twin_thr2 * ratio;
    int best = -1;
    double fb = 0.0;
    for (int s = 0; s < P.nsys; ++s) {
      if (!P.twin[s]) continue;
      const double fs = f.twinf[(long long)s * N + v];
      if (fs > fb) { fb = fs; best = s; }
    }
    if (best >= 0 && fb > thr) {
      double R[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = f.rot[k * N + v];
      const double n0 = P.nrm[best][0], n1 = P.nrm[best][1], n2 = P.nrm[best][2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double rn = R[3 * i] * n0 + R[3 * i + 1] * n1 + R[3 * i + 2] * n2;
        f.rot[(3 * i) * N + v] = 2.0 * rn * n0 - R[3 * i];
        f.rot[(3 * i + 1) * N + v] = 2.0 * rn * n1 - R[3 * i + 1];
        f.rot[(3 * i + 2) * N + v] = 2.0 * rn * n2 - R[3 * i + 2];
      }
      for (int s = 0; s < P.nsys; ++s) f.twinf[(long long)s * N + v] = 0.0;
      f.twinned[v] = 1;
      sums[0] = 1.0;
    }
  }
  warp_partials_store<kNSum>(sums, 0, partials, nw);
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
// every __global__ launch of the library goes through one of the launch_* functions below and is counted here
// (evp_launch_count: bench.py reports the launches inside its timed region from this counter)
static long long g_launches = 0;
long long launch_count() { return g_launches; }

// peer-memory stores: 1 = a block waits until its stores have been performed at the peer (cp.async.bulk.wait_group 0),
// 2 = only until shared memory has been read out (the kernel boundary and the cross-GPU barrier that follows order the writes)
int p2p_wait_mode() {
  static const int m = getenv("EVP_P2P_WAIT") ? (std::string(getenv("EVP_P2P_WAIT")) == "read" ? 2 : 1) : 1;
  return m;
}
bool fft_size_supported(int n) { return n >= 8 && n <= 1024 && (n & (n - 1)) == 0; }

// opt in to large dynamic shared memory (static + dynamic > 48 KB), once per kernel instantiation (the call is not free)
#define set_smem(bytes, ...)                                                                          \
  do {                                                                                                \
    static bool done_ = false;                                                                        \
    if (!done_ && (bytes) > 40 * 1024) cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
    done_ = true;                                                                                     \
  } while (0)

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

void launch_xfwd(int nx, const double *sig, double2 *W, long long N, int rowbase, int nrows, SpecLayout L, const double2 *tw,
                 cudaStream_t st) { g_launches += 1;
#define X_(NX)                                                                                        \
  {                                                                                                   \
    using C = XCfg<NX>;                                                                               \
    set_smem(C::smem, k_xfwd<NX>);                                                                    \
    dim3 grid((nrows + C::L - 1) / C::L, 3);                                                          \
    k_xfwd<NX><<<grid, C::T, C::smem, st>>>(sig, W, N, rowbase, nrows, L, tw);                         \
  }
  EVP_DISPATCH_N(nx, X_)
#undef X_
}

void launch_xinv(int nx, const double2 *W, double *e, double *de_dbg, const MacroDev *macro, long long N, int rowbase, int nrows,
                 SpecLayout L, const double2 *tw, cudaStream_t st) { g_launches += 1;
#define X_(NX)                                                                                        \
  {                                                                                                   \
    using C = XCfg<NX>;                                                                               \
    set_smem(C::smem, k_xinv<NX>);                                                                    \
    dim3 grid((nrows + C::L - 1) / C::L, 3);                                                          \
    k_xinv<NX><<<grid, C::T, C::smem, st>>>(W, e, de_dbg, macro, N, rowbase, nrows, L, tw);            \
  }
  EVP_DISPATCH_N(nx, X_)
#undef X_
}

void launch_ypass(int ny, bool inv, const PeerMaps &tin, bool pull, const PeerMaps &tout, bool p2p, TileInfo in, TileInfo out, int nxh, int nzc,
                  const double2 *tw, cudaStream_t st) { g_launches += 1;
#define Y_(NY)                                                                                        \
  {                                                                                                   \
    using C = YCfg<NY>;                                                                               \
    dim3 grid((nxh + C::TX - 1) / C::TX, (nzc + C::ZT - 1) / C::ZT, 6);                               \
    if (inv) {                                                                                        \
      set_smem(C::smem, k_ypass<NY, true>);                                                           \
      k_ypass<NY, true><<<grid, C::T, C::smem, st>>>(tin, tout, in.lg, in.chunk, out.lg, out.chunk, nzc, p2p ? p2p_wait_mode() : 0, pull ? 1 : 0, tw); \
    } else {                                                                                          \
      set_smem(C::smem, k_ypass<NY, false>);                                                          \
      k_ypass<NY, false><<<grid, C::T, C::smem, st>>>(tin, tout, in.lg, in.chunk, out.lg, out.chunk, nzc, p2p ? p2p_wait_mode() : 0, pull ? 1 : 0, tw); \
    }                                                                                                 \
  }
  EVP_DISPATCH_N(ny, Y_)
#undef Y_
}

void launch_zfused(int nz, int mode, bool one_shot, const ZMaps &tz, const ZOutMaps &tzo, bool p2p_, int lg_nzl, int lg_nzc, int zrun,
                   int nxh /* local kx columns */, int kx0, int nyl, int ky0, int nx, int ny, double dx, double dy, double dz, const double2 *tw, cudaStream_t st) { g_launches += 1;
  const int p2p = p2p_ ? p2p_wait_mode() : 0;
  const double rx = 1.0 / (nx * dx), ry = 1.0 / (ny * dy), rz = 1.0 / (nz * dz);
  const double scale = 1.0 / ((double)nx * ny * nz);
  static const int zver = getenv("EVP_ZKERNEL") ? atoi(getenv("EVP_ZKERNEL")) : 2;   // 1 = one-shot kernel, 2 = persistent radix-16
  static const int znb = getenv("EVP_ZNB") ? atoi(getenv("EVP_ZNB")) : 2;            // persistent kernel: resident blocks per SM (1 or 2)
  const bool fwd_only = mode == 1;
  if (mode == 0 && !one_shot && zver == 2 && nz == 512) {
    using C = Z3Cfg<512>;
    static int nsm3 = 0;
    if (!nsm3) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm3, cudaDevAttrMultiProcessorCount, dev); }
    const int nkx = (nxh + C::TX - 1) / C::TX, ntiles = nkx * nyl;
    set_smem(C::smem, k_zfused3<512>);
    k_zfused3<512><<<ntiles < nsm3 ? ntiles : nsm3, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, nkx, ntiles, tw);
    return;
  }
  // nz = 256: EVP_Z4=1 selects k_zfused4 (in place, per-component TMA pipeline).  Measured 0.520 ms against 0.505 ms for k_zfused2 at
  // 256^3 (same box, `profiles/r02_z4_{on,off}.json`): the second resident block already hides the tile load, so k_zfused2 stays the default
  const int z4 = getenv("EVP_Z4") ? atoi(getenv("EVP_Z4")) : 0;
  if (mode == 0 && !one_shot && zver == 2 && nz == 256 && z4) {
    using C = Z4Cfg<256>;
    static int nsm4 = 0;
    if (!nsm4) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm4, cudaDevAttrMultiProcessorCount, dev); }
    const int nkx = (nxh + C::TX - 1) / C::TX, ntiles = nkx * nyl;
    set_smem(C::smem, k_zfused4<256>);
    k_zfused4<256><<<ntiles < 2 * nsm4 ? ntiles : 2 * nsm4, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, nkx, ntiles, tw);
    return;
  }
  if (mode == 0 && !one_shot && zver == 2 && (nz == 128 || nz == 256)) {
    static int nsm = 0;
    if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
    if (nz == 256) {
      using C = Z2Cfg<256>;
      const int nkx = (nxh + C::TX - 1) / C::TX, ntiles = nkx * nyl;
      if (znb == 2) {
        set_smem(C::smem2, k_zfused2<256, 2>);
        k_zfused2<256, 2><<<ntiles < 2 * nsm ? ntiles : 2 * nsm, C::T2, C::smem2, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, nkx, ntiles, tw);
      } else {
        set_smem(C::smem, k_zfused2<256, 1>);
        k_zfused2<256, 1><<<ntiles < nsm ? ntiles : nsm, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, nkx, ntiles, tw);
      }
    } else {
      using C = Z2Cfg<128>;
      const int nkx = (nxh + C::TX - 1) / C::TX, ntiles = nkx * nyl;
      if (znb == 2) {
        set_smem(C::smem2, k_zfused2<128, 2>);
        k_zfused2<128, 2><<<ntiles < 2 * nsm ? ntiles : 2 * nsm, C::T2, C::smem2, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, nkx, ntiles, tw);
      } else {
        set_smem(C::smem, k_zfused2<128, 1>);
        k_zfused2<128, 1><<<ntiles < nsm ? ntiles : nsm, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, nkx, ntiles, tw);
      }
    }
    return;
  }
#define Z_(NZ)                                                                                        \
  {                                                                                                   \
    using C = ZCfg<NZ>;                                                                               \
    dim3 grid((nxh + C::TX - 1) / C::TX, nyl);                                                        \
    if (fwd_only) {                                                                                   \
      set_smem(C::smem, k_zfused<NZ, 1>);                                                             \
      k_zfused<NZ, 1><<<grid, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, tw); \
    } else if (mode == 2) {                                                                           \
      set_smem(C::smem, k_zfused<NZ, 2>);                                                             \
      k_zfused<NZ, 2><<<grid, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, tw); \
    } else {                                                                                          \
      set_smem(C::smem, k_zfused<NZ, 0>);                                                             \
      k_zfused<NZ, 0><<<grid, C::T, C::smem, st>>>(tz, tzo, p2p, lg_nzl, lg_nzc, zrun, ky0, kx0, nx, ny, rx, ry, rz, scale, tw); \
    }                                                                                                 \
  }
  EVP_DISPATCH_N(nz, Z_)
#undef Z_
}

int ypass_tx() { return 8; }
int zpass_tx(int nz) { return (nz >= 1024) ? 2 : ((nz >= 256) ? 4 : 8); }

static long long num_warps(long long N) { return ((N + kCB - 1) / kCB) * (kCB / 32); }
long long partial_doubles(long long N) { return (kNSum + 1) * num_warps(N); }
int reduce_scratch_doubles() { return kRedBlocks * 16; }

template <int NS_T, int NPOW_T, bool ONEPH, int MINB>
static void launch_const_t(const Fields &f, long long vbase, long long count, int nsmax, double *partials, cudaStream_t st) { g_launches += 1;
  const int nb = (int)((count + kCB - 1) / kCB);
  const size_t smem = (size_t)(33 + (nsmax > 0 ? nsmax : 1)) * kCB * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_constitutive_t<NS_T, NPOW_T, ONEPH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem > 48 * 1024 ? (int)smem : 48 * 1024);
    cudaFuncSetAttribute(k_constitutive_t<NS_T, NPOW_T, ONEPH, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_done = true;
  }
  static int pf = -1;
  if (pf < 0) pf = getenv("EVP_K1_PF") ? atoi(getenv("EVP_K1_PF")) : 148 * MINB;   // prefetch distance in blocks (0 = off)
  // partial slots: chunks are multiples of kCB voxels, so warp index = voxel / 32
  k_constitutive_t<NS_T, NPOW_T, ONEPH, MINB><<<nb, kCB, smem, st>>>(f, vbase, count, partials, num_warps(f.N), vbase / 32, pf > 0 ? pf : (1 << 30));
}


template <int NS_T, int NPOW_T, bool TWIN, int MINB, int G, int TAB = 0>
static void launch_const_p(const Fields &f, long long vbase, long long count, double *partials, cudaStream_t st) { g_launches += 1;
  const int nb = (int)((count + kCB - 1) / kCB);
  const size_t smem = (size_t)(21 + 18 + NS_T) * kCB * sizeof(double) + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_constitutive_p<NS_T, NPOW_T, TWIN, MINB, G, TAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_constitutive_p<NS_T, NPOW_T, TWIN, MINB, G, TAB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_done = true;
  }
  static int pf = -1;
  if (pf < 0) pf = getenv("EVP_K1_PF") ? atoi(getenv("EVP_K1_PF")) : 148 * MINB;
  const bool plain = getenv("EVP_K1_BULK") && atoi(getenv("EVP_K1_BULK")) == 0;
  k_constitutive_p<NS_T, NPOW_T, TWIN, MINB, G, TAB><<<nb, kCB, smem, st>>>(f, vbase, count, partials, num_warps(f.N), vbase / 32,
                                                                          plain ? -1 : (pf > 0 ? pf : (1 << 30)));
}

// the uniform-exponent fast path (k_constitutive_p) applies to: one phase, every system with the same integer exponent
// n in {10, 20}, and 12 systems without twins or 24 systems; n = 10 also with 30 systems.  Returns n-1, or -1 when the generic
// kernels are used.
int constitutive_fast_npow(int nphases, int uniform_ns, int uniform_npow, int any_twin) {
  const bool legacy = getenv("EVP_K1_LEGACY") && atoi(getenv("EVP_K1_LEGACY")) != 0;   // thread-loads kernel (A/B timing, tests)
  if (legacy || nphases != 1 || (uniform_npow != 9 && uniform_npow != 19)) return -1;
  if ((uniform_ns == 12 && !any_twin) || uniform_ns == 24) return uniform_npow;
  if (uniform_ns == 30 && uniform_npow == 9) return uniform_npow;   // HCP with tensile + compressive twins (evp_phase_hcp, with_twin = 2)
  return -1;
}

// fast_npow: the decision taken at evp_begin_increment (constitutive_fast_npow), which also fixed the form of the Jb tables and of
// f.itc for this increment; it is not re-derived here, so the kernel always matches the tables it reads
void launch_constitutive(const Fields &f, long long vbase, long long count, int nsmax, int nphases, int uniform_ns, int uniform_npow,
                         int any_twin, int fast_npow, double *partials, cudaStream_t st) {
  const bool one = nphases == 1;
  static const int minb = getenv("EVP_K1_MINB") ? atoi(getenv("EVP_K1_MINB")) : 0;   // tuning knobs
  if (fast_npow >= 0) {
    if (uniform_ns == 12 && !any_twin) {
      if (uniform_npow == 9) {
        if (minb == 3) return launch_const_p<12, 9, false, 3, 12>(f, vbase, count, partials, st);
        if (g_table == 1) return launch_const_p<12, 9, false, 4, 12, 1>(f, vbase, count, partials, st);
        return launch_const_p<12, 9, false, 4, 12>(f, vbase, count, partials, st);
      }
      if (g_table == 1) return launch_const_p<12, 19, false, 4, 12, 1>(f, vbase, count, partials, st);
      return launch_const_p<12, 19, false, 4, 12>(f, vbase, count, partials, st);
    }
    if (uniform_ns == 24) {
      if (uniform_npow == 9) {
        if (g_table == 2) return launch_const_p<24, 9, true, 3, 12, 2>(f, vbase, count, partials, st);
        return launch_const_p<24, 9, true, 3, 12>(f, vbase, count, partials, st);
      }
      if (g_table == 2) return launch_const_p<24, 19, true, 3, 12, 2>(f, vbase, count, partials, st);
      return launch_const_p<24, 19, true, 3, 12>(f, vbase, count, partials, st);
    }
    if (uniform_ns == 30) return launch_const_p<30, 9, true, 3, 10>(f, vbase, count, partials, st);   // 69 staged columns: 70 KB, 3 blocks per SM
  }
  if (one && uniform_ns == 12 && uniform_npow == 9) return launch_const_t<12, 9, true, 3>(f, vbase, count, nsmax, partials, st);
  if (one && uniform_ns == 12 && uniform_npow == 19) return launch_const_t<12, 19, true, 3>(f, vbase, count, nsmax, partials, st);
  if (one && uniform_ns == 12) return launch_const_t<12, -2, true, 3>(f, vbase, count, nsmax, partials, st);
  if (one && uniform_ns == 24 && uniform_npow == 9) return launch_const_t<24, 9, true, 3>(f, vbase, count, nsmax, partials, st);
  if (one && uniform_ns == 24 && uniform_npow == 19) return launch_const_t<24, 19, true, 3>(f, vbase, count, nsmax, partials, st);
  if (one && uniform_ns == 24) return launch_const_t<24, -2, true, 3>(f, vbase, count, nsmax, partials, st);
  return launch_const_t<0, -2, false, 3>(f, vbase, count, nsmax, partials, st);
}

__global__ void k_voxel_classes(Fields f) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < f.N; v += (long long)gridDim.x * blockDim.x) {
    f.orient[v] = (int32_t)v;
    f.orient_rep[v] = v;
  }
}
void launch_voxel_classes(const Fields &f, cudaStream_t st) { g_launches += 1; k_voxel_classes<<<592, 256, 0, st>>>(f); }

void launch_prep_increment(const Fields &f, int nsmax, int fast_npow, cudaStream_t st) { g_launches += 2;
  const int nb = (int)((f.norient + kCB - 1) / kCB);
  k_prep_orient<<<nb, kCB, 0, st>>>(f, fast_npow >= 0 ? 1 : 0);
  k_prep_itc<<<1184, 256, 0, st>>>(f, nsmax, fast_npow);
}

void launch_commit(const Fields &f, int nsmax, double dt, const double wapp[3], int texture, int twinning, double *partials, cudaStream_t st) { g_launches += 1;
  const int nb = (int)((f.N + kCB - 1) / kCB);
  const size_t smem = (size_t)(nsmax > 0 ? nsmax : 1) * kCB * sizeof(double);
  CommitParams cp{dt, {wapp[0], wapp[1], wapp[2]}, texture, twinning};
  k_commit<<<nb, kCB, smem, st>>>(f, cp, partials, num_warps(f.N));
}
void launch_twin_reorient(const Fields &f, double ratio, double *partials, cudaStream_t st) { g_launches += 1;
  const int nb = (int)((f.N + kCB - 1) / kCB);
  k_twin_reorient<<<nb, kCB, 0, st>>>(f, ratio, partials, num_warps(f.N));
}

void launch_reduce(const double *partials, long long N, double *scratch, double *totals, cudaStream_t st) { g_launches += 2;
  const long long nw = num_warps(N);
  const int nb = (int)((nw < kRedBlocks) ? nw : kRedBlocks);
  k_reduce1<<<nb, 256, 0, st>>>(partials, nw, scratch);
  k_reduce2<<<1, 32 * (kNSum + 1), 0, st>>>(scratch, nb, totals);
}
void launch_macro(const double *totals, MacroDev *macro, double ntot, cudaStream_t st) { g_launches += 1; k_macro<<<1, 32, 0, st>>>(totals, macro, ntot); }
void launch_fill(double *p, long long n, double v, cudaStream_t st) { g_launches += 1; k_fill<<<592, 256, 0, st>>>(p, n, v); }
void launch_init_crss(const Fields &f, int nsmax, cudaStream_t st) { g_launches += 1; k_init_crss<<<592, 256, 0, st>>>(f, nsmax); }


// ---------------------------------------------------------------------------------------------
// fp64 peak of this device, measured (evp_debug_fp64_peak): 8 dependent DFMA chains per thread, 32 warps per SM,
// every SM busy for about 1 ms; CUDA events on the caller's stream, best of `reps`.  The denominator of the
// constitutive kernel's roofline (MEASURED_PEAKS.json holds HBM and bf16 figures only).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, double a, double b, int iters) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 1.2345e300) out[0] = s;   // never true: keeps the chains alive without a store per thread
}
double measure_fp64_peak(int reps, cudaStream_t st) {
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  double *out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return 0.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1000, nb = nsm * 4, nt = 256;
  g_launches += 1 + (reps > 0 ? reps : 1);
  k_fp64_peak<<<nb, nt, 0, st>>>(out, 0.999999, 1e-7, 50);
  double best = 0.0;
  for (int r = 0; r < (reps > 0 ? reps : 1); ++r) {
    cudaEventRecord(e0, st);
    k_fp64_peak<<<nb, nt, 0, st>>>(out, 0.999999, 1e-7, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * (double)iters * 16 * 8 * nt * nb / (ms * 1e-3) * 1e-12;
    if (ms > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // namespace evp
