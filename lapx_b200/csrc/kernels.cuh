// kernels.cuh — hand-written sm_100a kernels of the EVPFFT equilibrium iteration.
//   K2 k_xfwd        a1  x pass: two real fields -> two half spectra (two-for-one complex FFT)
//   K3 k_ypass<fwd>  a1  y pass on tiles of 8 kx-columns
//   K4 k_zfused      a1+a2+a3  z forward FFT, Green operator per frequency, z inverse FFT
//   K5 k_ypass<inv>  a3
//   K6 k_xinv        a3  x inverse pass fused with  e <- e - de + dE
//   K1 k_constitutive a4+a5+a6  crystal-frame Newton, norms by warp shuffles
//      k_reduce / k_macro   a6 second stage + a7 on the device
//      k_commit      §8(f).1 per-increment state update (plastic strain, Voce hardening)
// Reference counterpart: absent (mount holds only LICENSE); units per SURVEY.md §8(a).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "evp_core.h"

namespace evp {

// ---------------------------------------------------------------------------------------------
// spectral buffer addressing.  One formula serves the plain local layout, the all-to-all "send"
// layout (rows grouped by destination rank) and the "recv" layout (planes grouped by source).
//   y-split view : addr(c, zl, y)   = (y / nyl) * dstride + c * cstride + zl * zstride + (y % nyl) * nxp
//   z-split view : addr(c, z, yl)   = (z / nzl) * dstride + c * cstride + (z % nzl) * zstride + yl * nxp
// with cstride = nzl*nyl*nxp, zstride = nyl*nxp, dstride = 6*cstride.  Single GPU: nyl = ny, nzl = nz.
// ---------------------------------------------------------------------------------------------
// x-stage view (k_xfwd output / k_xinv input; pencil decomposition: kx pieces grouped by the rank of the row that owns them):
//   addr(c, row, k)   = (k / kxl) * dstride + c * cstride + row * kxl + (k % kxl),  row = zl * nyl + y_local.
// Slab / single GPU: kxl = nxp, i.e. the plain layout.
struct SpecLayout {
  int nyl, nzl, nxp, nxh;      // nxp: row pitch (= kxl), nxh: nx/2 + 1 (global)
  int lg_nyl, lg_nzl;          // all extents are powers of two: divisions become shifts
  long long cstride, zstride, dstride;
  __host__ __device__ long long row_x(int c, int row, int k) const {
    int j = 0;
    while (k >= nxp) { k -= nxp; ++j; }   // at most py - 1 rounds; never taken for slabs
    return (long long)j * dstride + (long long)c * cstride + (long long)row * nxp + k;
  }
  __host__ __device__ long long row_ysplit(int c, int zl, int y) const {
    return (long long)(y >> lg_nyl) * dstride + (long long)c * cstride + (long long)zl * zstride + (long long)(y & (nyl - 1)) * nxp;
  }
  __host__ __device__ long long row_zsplit(int c, int z, int yl) const {
    return (long long)(z >> lg_nzl) * dstride + (long long)c * cstride + (long long)(z & (nzl - 1)) * zstride + (long long)yl * nxp;
  }
};

// device-resident macroscopic state (rows a6/a7); one copy per handle
struct MacroDev {
  double E[6], Et[6], dEpend[6], savg[6], scau[6];
  double Mmac[36];          // dE = Mmac (scau - savg): (C0_TT)^-1 embedded, zero rows for strain control
  double err_s, err_e, newton_mean;
  double epavg[6];
  int newton_max, nonfinite, iter, pad;
  long long unconverged;
};

struct Fields {
  double *sig, *e, *epsp, *edotp, *crss, *rot, *gacc, *twinf, *de;
  double *wrot;             // local rotation fluctuation (3) of the committed strain field
  int32_t *twinned;         // PTR flag
  double *mrot, *jb, *itc;  // per-increment invariants: deviatoric rotation M (25), Jb = S0_c + S_c (21), 1/tau_c
  int32_t *grain, *phase;
  int32_t *orient;          // orientation class of every voxel (index into mrot/jb tables)
  long long *orient_rep;    // per class: a local voxel carrying that orientation (or -1)
  long long norient;        // table length: #grains (texture per grain) or N (texture per voxel)
  long long N;              // local voxels; component stride of every SoA field
};


void upload_phase_tables(const PhaseDev *ph, int nph);
void upload_green(const GreenConst &g);
void upload_const_params(const ConstParams &p);

// launches (all on `st`)
void launch_xfwd(int nx, const double *sig, double2 *W, long long N, int rowbase, int nrows, SpecLayout L, const double2 *tw,
                 cudaStream_t st);
// TMA tiling of one spectral-buffer view: rows (y or z) per op and log2 of the rows per rank chunk
struct TileInfo { int lg, chunk; };
constexpr int kMaxChunks = 8;
constexpr int kMaxRanks = 8;
constexpr int kMaxChunksP2P = 4;   // pipeline chunks when the transposes are peer-memory stores (maps are kernel parameters)
struct ZMaps { CUtensorMap m[kMaxChunks]; };   // one tensor map per pipeline chunk of the z-split buffer
struct PeerMaps { CUtensorMap m[kMaxRanks]; }; // y pass input / output: m[0] = local layout, or one map per source / destination rank (peer memory)
struct ZOutMaps { CUtensorMap m[kMaxRanks * kMaxChunksP2P]; };   // z pass output in p2p mode: [destination rank][chunk]
// pull: input rows of source rank d are TMA-loaded from tin.m[d] (peer mapping);  p2p: output rows of destination rank d are
// TMA-stored through tout.m[d]
void launch_ypass(int ny, bool inv, const PeerMaps &tin, bool pull, const PeerMaps &tout, bool p2p, TileInfo in, TileInfo out, int nxh, int nzc,
                  const double2 *tw, cudaStream_t st);
void launch_zfused(int nz, int mode /* 0 Green, 1 forward only, 2 local rotation */, bool one_shot, const ZMaps &tz, const ZOutMaps &tzo, bool p2p, int lg_nzl, int lg_nzc, int zrun,
                   int nxv /* local kx columns */, int kx0 /* first global kx */, int nyl, int ky0, int nx, int ny, double dx, double dy, double dz, const double2 *tw,
                   cudaStream_t st);
int ypass_tx();
int zpass_tx(int nz);
void launch_xinv(int nx, const double2 *W, double *e, double *de_dbg, const MacroDev *macro, long long N, int rowbase, int nrows,
                 SpecLayout L, const double2 *tw, cudaStream_t st);
void launch_constitutive(const Fields &f, long long vbase, long long count, int nsmax, int nphases, int uniform_ns, int uniform_npow, int any_twin,
                         int fast_npow /* decision of constitutive_fast_npow at evp_begin_increment */, double *partials, cudaStream_t st);
void launch_reduce(const double *partials, long long N, double *scratch, double *totals, cudaStream_t st);
long long partial_doubles(long long N);
int reduce_scratch_doubles();
void launch_macro(const double *totals, MacroDev *macro, double ntot_global, cudaStream_t st);
void launch_voxel_classes(const Fields &f, cudaStream_t st);
void launch_prep_increment(const Fields &f, int nsmax, int fast_npow, cudaStream_t st);
int constitutive_fast_npow(int nphases, int uniform_ns, int uniform_npow, int any_twin);
void launch_commit(const Fields &f, int nsmax, double dt, const double wapp[3], int texture, int twinning, double *partials, cudaStream_t st);
void launch_twin_reorient(const Fields &f, double ratio, double *partials, cudaStream_t st);
void launch_fill(double *p, long long n, double v, cudaStream_t st);
void launch_init_crss(const Fields &f, int nsmax, cudaStream_t st);
bool fft_size_supported(int n);
long long launch_count();
double measure_fp64_peak(int reps, cudaStream_t st);   // TFLOP/s, dependent-chain DFMA microbenchmark
int constitutive_block();

}  // namespace evp
