/*
 * evpfft.h — C ABI of the B200-native EVPFFT equilibrium loop (lanl/LApx hot path).
 *
 * Boundary provenance.  /root/reference holds only LICENSE (LICENSE:1, LICENSE:3), so no
 * reference FFI/plugin surface can be cited file:line.  Every entry point below implements
 * the proposal of SURVEY.md §8(b), which in turn restates BASELINE.json:5
 * ("Host code ... calls CUDA through a thin C-ABI shim").  The unit each call drives is
 * named by its SURVEY.md §8(a) row (a1..a7) or §8(f) row.
 *
 * The same ABI is implemented twice:
 *   lapx_b200/csrc  -> lapx_b200/libevpfft_b200.so   (CUDA sm_100a product path; no CPU fallback)
 *   oracle/         -> oracle/libevp_oracle.so       (CPU restatement; test infrastructure only)
 *
 * Conventions
 *   - plain C types only; the caller owns every host pointer; the library copies in/out and
 *     owns all device memory.  No C++ exception crosses this boundary.
 *   - return value: 0 = EVP_OK, negative = evp_status; evp_last_error() gives text.
 *   - fields are structure-of-arrays [component][z][y][x], x fastest, fp64;
 *     ids are int32 [z][y][x].  In a distributed solver every rank passes / receives only
 *     its own z-slab [component][z_local][y][x].
 *   - symmetric tensors use 6 components in the order 11,22,33,23,13,12 (tensor, not
 *     engineering, shear components).
 *   - a handle is bound to one GPU and one stream; calls on a handle are stream ordered;
 *     reports are written after the stream has been synchronised.  Not thread-safe.
 *   - ONE GPU PER PROCESS: kernel attributes and the __constant__ tables of the product library are
 *     per-device state; evp_create on a second device of the same process fails with
 *     EVP_ERR_UNSUPPORTED (multi-GPU runs use one process per GPU, see evp_dist).
 */
#ifndef EVPFFT_H
#define EVPFFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVP_ABI_VERSION 2
#define EVP_MAX_SYS    32   /* slip + twin systems per phase (FCC 12, HCP 24..30) */
#define EVP_MAX_MODES   8   /* deformation modes per phase                          */
#define EVP_MAX_PHASES  4

typedef struct evp_solver *evp_handle;

typedef enum {
  EVP_OK = 0,
  EVP_ERR_ARG = -1,        /* bad argument / size mismatch                    */
  EVP_ERR_STATE = -2,      /* call out of order (e.g. iter before loading)    */
  EVP_ERR_DEVICE = -3,     /* CUDA / NCCL failure, or no sm_100 device        */
  EVP_ERR_UNSUPPORTED = -4,/* grid size / option not supported by this build  */
  EVP_ERR_NUMERIC = -5     /* Newton failed / non-finite value detected       */
} evp_status;

typedef struct {
  int32_t nx, ny, nz;      /* GLOBAL grid                                    */
  double  dx, dy, dz;      /* voxel edge lengths (only their ratios matter)  */
} evp_grid;

/* One crystalline phase.  Constitutive law: SURVEY.md §8(a) row a4. */
typedef struct {
  int32_t nsys;                         /* total slip+twin systems                */
  int32_t nmodes;
  double  c_voigt[36];                  /* crystal-frame stiffness, 6x6 row major,
                                           standard Voigt (C44 = C_2323)          */
  double  b[EVP_MAX_SYS][3];            /* slip / twin shear direction (crystal Cartesian) */
  double  n[EVP_MAX_SYS][3];            /* plane normal (crystal Cartesian)       */
  int32_t mode[EVP_MAX_SYS];            /* mode index of each system              */
  int32_t twin[EVP_MAX_MODES];          /* 1: unidirectional (twin) mode          */
  double  gamma0[EVP_MAX_MODES];        /* reference shear rate                   */
  double  nrate[EVP_MAX_MODES];         /* rate exponent n                        */
  double  tau0[EVP_MAX_MODES];          /* extended Voce: initial CRSS            */
  double  tau1[EVP_MAX_MODES];
  double  theta0[EVP_MAX_MODES];
  double  theta1[EVP_MAX_MODES];
  double  hlat[EVP_MAX_MODES][EVP_MAX_MODES]; /* latent hardening, mode x mode   */
  double  twin_shear[EVP_MAX_MODES];    /* characteristic twin shear              */
  double  twin_thr1, twin_thr2;         /* PTR thresholds                         */
} evp_phase;

/* Transport of the FFT transposes between ranks (evp_dist.transport). */
typedef enum {
  EVP_TRANSPORT_AUTO = 0,  /* peer-memory TMA stores when CUDA IPC mapping works on every rank, else NCCL */
  EVP_TRANSPORT_NCCL = 1,  /* grouped ncclSend/ncclRecv all-to-all                                          */
  EVP_TRANSPORT_P2P  = 2   /* peer-memory stores required: evp_create fails if IPC is unavailable          */
} evp_transport_kind;

/* Domain decomposition.  nranks = 1 for a single GPU.
 *   py <= 1 : z-SLABS.  Rank r owns the planes z in [r*nz/nranks, (r+1)*nz/nranks); two transposes per iteration.
 *   py >= 2 : PENCILS on a py x pz process grid (pz = nranks/py, rank = iy*pz + iz).  Rank (iy,iz) owns the block
 *             y in [iy*ny/py, ...), z in [iz*nz/pz, ...), all x; four transposes per iteration (x<->y inside a row of
 *             py ranks, y<->z inside a column of pz ranks).  NCCL transport only.                                   */
typedef struct {
  int32_t nranks;
  int32_t rank;
  int32_t device;                       /* CUDA device ordinal for this rank      */
  int32_t transport;                    /* evp_transport_kind                     */
  uint8_t nccl_id[128];                 /* ncclUniqueId from evp_nccl_unique_id() */
  int32_t py;                           /* 0/1 = slab, >= 2 = pencil rows         */
  int32_t reserved[3];
} evp_dist;

typedef struct {
  double  tol_stress;                   /* stop when err_stress <= tol_stress ... */
  double  tol_strain;                   /* ... and err_strain <= tol_strain       */
  int32_t itmax;
  int32_t itmin;
  double  tol_newton;                   /* per-voxel Newton: |dsig| <= tol*|sig|  */
  int32_t newton_itmax;
  int32_t update_texture;               /* commit: lattice rotation (macro spin + local FFT spin - plastic spin) */
  int32_t update_twinning;              /* commit: twin fractions and PTR reorientation (Tome, Lebensohn, Kocks 1991) */
} evp_ctrl;

typedef struct {
  int32_t iter;                         /* iteration index inside the increment (1-based) */
  int32_t newton_max;                   /* max Newton iterations over voxels      */
  double  newton_mean;
  double  err_stress;                   /* <|sig_new - sig_old|> / |<sig>|   (a6) */
  double  err_strain;                   /* <|eps(sig) - e|> / |E|            (a6) */
  double  savg[6];                      /* <sig>                                   */
  double  emacro[6];                    /* macroscopic strain E after a7           */
  int32_t converged;                    /* tolerances met AND unconverged == 0    */
  int32_t nonfinite;                    /* voxels whose Newton produced a non-finite value */
  int64_t unconverged;                  /* voxels whose Newton hit newton_itmax without meeting tol_newton (all ranks) */
} evp_iter_report;

typedef struct {
  int32_t iters;
  int32_t converged;
  double  err_stress, err_strain;
  double  savg[6];
  double  emacro[6];
  double  epavg[6];                     /* <eps_plastic> after commit             */
  double  seconds;                      /* wall time of the increment             */
  double  twin_acc;                     /* F_acc: twin volume fraction accumulated over the whole history (never decreases:
                                           reorientation zeroes a voxel's fractions but not its contribution to F_acc) */
  double  twin_eff;                     /* reoriented volume fraction F_eff       */
  int64_t reoriented;                   /* voxels reoriented by this commit (all ranks) */
} evp_step_report;

typedef enum {
  EVP_FIELD_STRESS = 0,        /* 6  fp64  sigma (== Lagrange multiplier lambda) */
  EVP_FIELD_STRAIN = 1,        /* 6  fp64  compatible strain e (total, t+dt)                    */
  EVP_FIELD_PLASTIC_STRAIN = 2,/* 6  fp64  eps_p committed at t                                 */
  EVP_FIELD_PLASTIC_RATE = 3,  /* 6  fp64  eps_p rate at the last evaluated stress              */
  EVP_FIELD_CRSS = 4,          /* nsys_max fp64  critical resolved shear stresses               */
  EVP_FIELD_ROTATION = 5,      /* 9  fp64  crystal->sample rotation, row major                  */
  EVP_FIELD_GRAIN = 6,         /* 1  int32                                                      */
  EVP_FIELD_PHASE = 7,         /* 1  int32                                                      */
  EVP_FIELD_GAMMA_ACC = 8,     /* 1  fp64  accumulated shear                                    */
  EVP_FIELD_TWIN_FRACTION = 9, /* nsys_max fp64 accumulated twin volume fraction per system     */
  EVP_FIELD_STRAIN_INCR = 10,  /* 6  fp64  last Gamma*sigma correction (debug / tests)          */
  EVP_FIELD_LOCAL_ROTATION = 11,/* 3 fp64  local rotation fluctuation (w32,w13,w21) of the compatible strain field   */
  EVP_FIELD_TWINNED = 12       /* 1  int32 1 = voxel has been reoriented by PTR                 */
} evp_field;

/* ---- life cycle ----------------------------------------------------------------------- */
int  evp_abi_version(void);
/* "cuda-sm100a" for the product library, "cpu-oracle" for the oracle. */
const char *evp_backend(void);

int  evp_create(const evp_grid *grid, const evp_phase *phases, int32_t nphases,
                const evp_dist *dist /* NULL = single GPU, device 0 */, evp_handle *out);
int  evp_destroy(evp_handle h);
const char *evp_last_error(evp_handle h /* may be NULL: creation errors */);

/* Local slab owned by this handle: z in [z0, z0+nzl). */
int  evp_local_slab(evp_handle h, int32_t *z0, int32_t *nzl);
/* Local block (pencil decomposition; for slabs y0 = 0, nyl = ny): fields passed to / returned by this handle are
 * [component][z_local][y_local][x].                                                                             */
int  evp_local_block(evp_handle h, int32_t *y0, int32_t *nyl, int32_t *z0, int32_t *nzl);
/* Number of CRSS / twin-fraction components stored per voxel (max nsys over phases). */
int  evp_nsys_max(evp_handle h);

/* ---- set-up --------------------------------------------------------------------------- */
/* grain/phase ids and crystal->sample rotations of the local slab; resets all state fields
 * (sigma = e = eps_p = 0, CRSS = tau0 of the voxel's phase).                              */
int  evp_set_microstructure(evp_handle h, const int32_t *grain, const int32_t *phase,
                            const double *rot9 /* [9][z][y][x] */);
/* Reference medium C0, 6x6 Voigt.  c0 == NULL: Voigt average of the rotated crystal
 * stiffnesses over all voxels (all ranks).                                               */
int  evp_set_reference_medium(evp_handle h, const double *c0_voigt36);
int  evp_get_reference_medium(evp_handle h, double *c0_voigt36);
int  evp_set_control(evp_handle h, const evp_ctrl *ctrl);

/* Mixed boundary conditions (row a7).  iudot[3][3]/udot: imposed velocity-gradient
 * components; iscau[6]/scau: imposed Cauchy stress components (value reached at the end of
 * the increment).  For every symmetric component exactly one of the two must be imposed. */
int  evp_set_loading(evp_handle h, const int32_t iudot[9], const double udot[9],
                     const int32_t iscau[6], const double scau[6]);

/* ---- the hot path --------------------------------------------------------------------- */
int  evp_begin_increment(evp_handle h, double dt);
/* ONE fixed-point iteration = rows a1..a7; the timed unit of BASELINE.json "metric".     */
int  evp_equilibrium_iter(evp_handle h, evp_iter_report *rep);
/* Same iteration split in its two halves, for unit parity tests:                          */
int  evp_op_green(evp_handle h);                       /* a1+a2+a3: e <- e - G0*sigma + dE  */
int  evp_op_constitutive(evp_handle h, evp_iter_report *rep); /* a4+a5+a6 (+a7 on the host) */
/* Commit the increment (§8(f).1): eps_p, Voce hardening, lattice rotation, twinning.      */
int  evp_end_increment(evp_handle h, evp_step_report *rep);
/* begin + iterate to tolerance/itmax + end.                                               */
int  evp_step(evp_handle h, double dt, evp_step_report *rep);

/* Enqueue `n` iterations back to back without host round trips (macro update a7 for
 * stress-controlled components is then applied on the device); reports only the last.    */
int  evp_equilibrium_iters(evp_handle h, int32_t n, evp_iter_report *last);

/* ---- field access --------------------------------------------------------------------- */
int  evp_field_components(evp_handle h, evp_field f);
int  evp_get_field(evp_handle h, evp_field f, void *host, size_t bytes);
int  evp_set_field(evp_handle h, evp_field f, const void *host, size_t bytes);
int  evp_get_macro(evp_handle h, double emacro[6], double savg[6]);

/* ---- checkpoint / restart ------------------------------------------- */
/* Restart file format "EVPCKPT2" — one raw little-endian binary file per rank, identical in both back ends:
 *   header   char magic[8] = "EVPCKPT2"; int32 nx, ny, nz, y0, nyl, z0, nzl (local block), nsys_max, nranks, rank;
 *            double Et[6] (macro strain at t), Edot_prev[6]; int64 ntwinned; double facc (accumulated twin fraction)
 *   payload  the state fields of the local block in the ABI layout [component][z_local][y_local][x], in this order:
 *            stress (6 fp64), strain (6), plastic strain (6), CRSS (ns), rotation (9), accumulated shear (1),
 *            twin fractions (ns), local rotation (3), grain (int32), phase (int32), twin flags (int32);  ns = max(nsys_max, 1)
 * The file size is checked against the header BEFORE any field is overwritten (a truncated file leaves the state untouched).
 * evp_load_state needs a handle created with the same grid / phases / decomposition and with the reference medium and
 * loading already set.                                                                                               */
int  evp_save_state(evp_handle h, const char *path);
int  evp_load_state(evp_handle h, const char *path);

/* ---- test / measurement hooks (same in both back ends) ------------------------------- */
/* Half spectrum of the forward 3-D r2c FFT (row a1) of stress component `comp`:
 * out[(z*ny + y)*(nx/2+1) + kx] as interleaved (re,im).  Single-rank handles only.        */
int  evp_debug_spectrum(evp_handle h, int32_t comp, double *out_reim);
/* Stream the handle runs on (cudaStream_t as void*), NULL in the oracle.                  */
void *evp_stream(evp_handle h);
/* Device time in ms of each kernel of the last evp_equilibrium_iter (CUDA events):
 * [0]=x fwd [1]=y fwd [2]=z fused [3]=y inv [4]=x inv+update [5]=constitutive [6]=exchange
 * [7]=whole iteration.  Only filled when evp_set_profiling(h,1).  Flag bits: 1 = kernel timers,
 * 2 = keep the strain increment field (EVP_FIELD_STRAIN_INCR), 4 = use the one-shot z kernel.      */
int  evp_set_profiling(evp_handle h, int32_t on);
int  evp_last_kernel_ms(evp_handle h, double ms[8]);

/* ---- host-side helpers (no GPU needed; product library only) ------------------------- */
/* Fill `out` with the FCC {111}<110> (12 systems) / HCP tables.                           */
int  evp_phase_fcc(evp_phase *out, double c11, double c12, double c44,
                   double gamma0, double nrate, double tau0, double tau1,
                   double theta0, double theta1);
int  evp_phase_hcp(evp_phase *out, double covera, const double c5[5] /* C11 C12 C13 C33 C44 */,
                   int32_t with_twin, double gamma0, double nrate,
                   const double tau0_mode[4], const double voce_mode[4][3]);
/* Periodic Voronoi tessellation, integer exact; writes grain ids of the
 * z-range [z0,z0+nzl) and per-grain rotations (ngrains*9).                                */
int  evp_voronoi(const evp_grid *g, int32_t ngrains, uint64_t seed, int32_t z0, int32_t nzl,
                 int32_t *grain_out, double *grain_rot9_out);
int  evp_nccl_unique_id(uint8_t id[128]);
/* Transport in use: EVP_TRANSPORT_NCCL (also for a single rank) or EVP_TRANSPORT_P2P. */
int  evp_transport(evp_handle h);
/* Hex digest of the sources this library was compiled from (lapx_b200/build.py embeds it; build() compares it
 * with the digest of the sources in the tree, so that a stale binary is never what gets benchmarked).          */
const char *evp_build_id(void);
/* Kernel launches enqueued by this handle since creation (every __global__ launch of the library is counted). */
int64_t evp_launch_count(evp_handle h);
/* fp64 peak of the device the handle runs on, measured now: dependent-chain DFMA microbenchmark (8 chains per
 * thread, 32 warps per SM, every SM), CUDA-event timed; best of `reps`.  The roofline denominator of the
 * constitutive kernel ("fp64 peak to be measured").                                            */
int  evp_debug_fp64_peak(evp_handle h, int32_t reps, double *tflops);
/* Per-voxel text microstructure: one line per voxel "phi1 Phi phi2 i j k grain phase" with Bunge
 * Euler angles in degrees, 1-based voxel indices (any order) and 1-based phase ids.  read: fills grain/phase
 * [z][y][x] and rot9 [9][z][y][x] (crystal->sample); write: the inverse (angles recovered from rot9).            */
int  evp_read_microstructure_txt(const char *path, const evp_grid *g, int32_t *grain, int32_t *phase, double *rot9);
int  evp_write_microstructure_txt(const char *path, const evp_grid *g, const int32_t *grain, const int32_t *phase,
                                  const double *rot9);

#ifdef __cplusplus
}
#endif
#endif /* EVPFFT_H */
