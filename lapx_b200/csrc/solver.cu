// solver.cu — host side of the product library: the C ABI of include/evpfft.h on top of the
// sm_100a kernels (kernels.cu).  There is NO CPU fallback here: without an sm_100 device
// evp_create fails with EVP_ERR_DEVICE.
// Reference counterpart: absent (/root/reference holds only LICENSE); ABI per SURVEY.md §8(b).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "host_math.h"

using namespace evp;
using namespace evp::host;

#ifndef EVP_SRC_HASH
#define EVP_SRC_HASH "unknown"
#endif

namespace {

std::string g_create_error;
int g_device = -1;   // the one device this process drives (kernel attributes / __constant__ tables are per device)

// ---- NCCL through dlopen (torch's bundled libnccl.so.2 if the process already loaded it) ----
struct Id128 { char b[128]; };
struct Nccl {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommSplit)(void *, int, int, void **, void *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
bool load_nccl(std::string *err) {
  if (g_nccl.lib) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) { *err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
  auto sym = [&](const char *s) { return dlsym(lib, s); };
  g_nccl.GetUniqueId = (int (*)(void *))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void **, int, Id128, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))sym("ncclRecv");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclAllReduce");
  g_nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
  g_nccl.CommSplit = (int (*)(void *, int, int, void **, void *))sym("ncclCommSplit");
  g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))sym("ncclAllGather");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.AllReduce) {
    *err = "libnccl lacks required symbols";
    return false;
  }
  g_nccl.lib = lib;
  return true;
}
constexpr int kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2, kNcclChar = 0;  // ncclFloat64, ncclSum, ncclMax, ncclInt8

}  // namespace

struct evp_solver {
  evp_grid g{};
  int nx = 0, ny = 0, nz = 0, nxh = 0, nxp = 0;
  int nranks = 1, rank = 0, device = 0;
  // decomposition: process grid py x pz, rank = iy*pz + iz (slab: py = 1).  Real space: block [nzl][nyb][nx] at (z0, y0);
  // y stage: [nzl][ny][kxl] at kx0 (nxv valid columns);  z stage: [nz][nyl][kxl] at (ky0, kx0)
  int py = 1, pz = 1, iy = 0, iz = 0;
  int nzl = 0, z0 = 0, nyl = 0, ky0 = 0, nyb = 0, y0 = 0, kxl = 0, kx0 = 0, nxv = 0;
  long long N = 0;        // local voxels
  double Ntot = 0;        // global voxels
  int nphases = 0, nsmax = 0;
  std::vector<evp_phase> ph;
  std::vector<PhaseDev> phd;
  cudaStream_t st = nullptr;
  Fields f{};
  double2 *WA = nullptr, *WB = nullptr;
  double2 *twx = nullptr, *twy = nullptr, *twz = nullptr;
  MacroDev *d_macro = nullptr, *h_macro = nullptr;  // h_macro pinned
  double *d_partials = nullptr, *d_totals = nullptr, *d_scratch = nullptr;
  int uniform_ns = 0, uniform_npow = -2;
  int k1_fast = -1;       // >= 0: the uniform-exponent fast path runs this increment (set by evp_begin_increment)
  bool any_twin = false;
  // The local slab is processed in `nchunks` z-chunks of nzc planes, each with its own sub-buffers, so that with
  // ranks > 1 the all-to-alls of one chunk run (on the communication stream) under the kernels of the others.
  struct Chunk {
    int rowbase = 0;                 // first (z,y) row of the chunk
    long long vbase = 0, count = 0;  // voxel range
    double2 *WA = nullptr, *WB = nullptr;
    CUtensorMap tm_y_plain{}, tm_y_split{};
    PeerMaps in_fwd{};               // forward y pass input: m[0] = local plain layout
    PeerMaps out_fwd{};              // forward y pass output: m[0] = local send layout, or one map per destination rank (p2p)
    PeerMaps in_inv{};               // inverse y pass input: m[0] = local split layout
    PeerMaps in_pull{};              // inverse y pass input, p2p pull: one map per source rank (that rank's z-pass buffer)
    PeerMaps out_inv{};              // inverse y pass output: m[0] = local plain layout
    cudaEvent_t ev_pull = nullptr;   // p2p pull: this chunk's inverse y pass (way-back transpose) has finished
    cudaEvent_t ev_fwd = nullptr, ev_a1 = nullptr, ev_a2 = nullptr;
    cudaEvent_t ev_p[4] = {nullptr, nullptr, nullptr, nullptr};   // pencil: kernel -> exchange hand-offs
    cudaEvent_t ev_q[2] = {nullptr, nullptr};                     // pencil: row exchanges done (way back, forward)
  };
  int nchunks = 1, nzc = 0;
  Chunk ch[kMaxChunks];
  ZMaps zmaps{};
  SpecLayout Lplain{}, Lsplit{};     // per-chunk layouts (nzl := nzc)
  TileInfo ti_y_plain{}, ti_y_split{};
  int zrun = 0, lg_nzl = 0, lg_nzc = 0;
  int zrun_z = 0, lg_nzc_z = 0;      // what the z pass sees: per-chunk sub-buffers, or (pull) one receive buffer for the whole slab
  cudaStream_t stc = nullptr;        // communication stream
  cudaEvent_t ev_k4 = nullptr, ev_it0 = nullptr, ev_it1 = nullptr;
  bool green_inflight = false;       // forward FFT + Green + way-back exchange of the CURRENT stress already enqueued
  void *comm2 = nullptr;             // second communicator for the small all-reduces (compute stream)
  void *comm_row = nullptr, *comm_col = nullptr;   // pencil: x<->y exchange among the py ranks of a row, y<->z among the pz of a column
  // peer-memory transport: the transposes are TMA stores into the other ranks' buffers (CUDA IPC mappings)
  bool p2p = false;
  bool pull = false;                 // p2p: way back = TMA loads from the peers' z-pass buffers in the inverse y pass (default), instead
                                     // of TMA stores into the peers' way-back buffers in the z pass (EVP_WAYBACK=push)
  cudaStream_t stp = nullptr;        // pull stream
  cudaEvent_t ev_z = nullptr;
  double2 *WC = nullptr;             // p2p: receive buffer of the forward transpose (written by every rank's y pass)
  double2 *peerWA[kMaxRanks]{}, *peerWC[kMaxRanks]{};
  ZOutMaps zout{};
  double *d_bar = nullptr;           // barrier payload
  cudaEvent_t ev_b1 = nullptr;
  struct Timers { cudaEvent_t a[128], b[128]; int type[128]; cudaStream_t s[128]; int n = 0; bool made = false; } tm;
  double C0m[36]{}, S0m[36]{};
  ConstParams cp{};
  GreenConst green{};
  bool have_micro = false, have_c0 = false, have_loading = false, in_incr = false;
  evp_ctrl ctrl{1e-6, 1e-6, 100, 1, 1e-6, 100, 0, 0};
  long long ntwinned = 0;   // voxels reoriented by PTR so far (all ranks)
  double facc = 0.0;        // F_acc: twin volume fraction accumulated over the history (all ranks; never decreases)
  char *stage[2] = {nullptr, nullptr};   // pinned staging buffers of evp_get_field / evp_set_field / evp_set_microstructure
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  int iudot[9]{}, iscau[6]{};
  double udot[9]{}, scau[6]{};
  bool strain_ctl[6]{};
  double Et[6]{}, Edot_prev[6]{};
  double dt = 0;
  int flags = 0;  // bit0: kernel timing, bit1: keep strain increment, bit2: one-shot z kernel instead of the persistent one
  double last_ms[8]{};
  void *comm = nullptr;
  std::string err;
};

namespace {

int fail(evp_handle h, int code, const std::string &m) {
  if (h) h->err = m; else g_create_error = m;
  return code;
}
#define CUDA_OK(h, call)                                                                          \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(h, EVP_ERR_DEVICE, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int ilog2i(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

// 5-D tensor [2*nxp doubles][nyl][nzl][6][P] over a spectral buffer; box = [2*tx][box_y][box_z][1][1]
bool make_tmap(CUtensorMap *m, void *base, const SpecLayout &L, int P, int tx, int box_y, int box_z, std::string *err, bool swizzle128 = false) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      *err = "cuTensorMapEncodeTiled entry point not available";
      return false;
    }
    encode = (PFN_cuTensorMapEncodeTiled)fn;
  }
  const cuuint64_t dims[5] = {(cuuint64_t)2 * L.nxp, (cuuint64_t)L.nyl, (cuuint64_t)L.nzl, 6, (cuuint64_t)P};
  const cuuint64_t strides[4] = {(cuuint64_t)L.nxp * 16, (cuuint64_t)L.zstride * 16, (cuuint64_t)L.cstride * 16, (cuuint64_t)L.dstride * 16};
  const cuuint32_t box[5] = {(cuuint32_t)2 * tx, (cuuint32_t)box_y, (cuuint32_t)box_z, 1, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
    return false;
  }
  return true;
}

double2 *make_twiddles(int n) {
  std::vector<double2> t(n);
  for (int k = 0; k < n; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
    t[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  double2 *d = nullptr;
  if (cudaMalloc(&d, sizeof(double2) * n) != cudaSuccess) return nullptr;
  cudaMemcpy(d, t.data(), sizeof(double2) * n, cudaMemcpyHostToDevice);
  return d;
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
void *field_ptr(evp_solver *S, int f, size_t *el) {
  *el = sizeof(double);
  switch (f) {
    case EVP_FIELD_STRESS: return S->f.sig;
    case EVP_FIELD_STRAIN: return S->f.e;
    case EVP_FIELD_PLASTIC_STRAIN: return S->f.epsp;
    case EVP_FIELD_PLASTIC_RATE: return S->f.edotp;
    case EVP_FIELD_CRSS: return S->f.crss;
    case EVP_FIELD_ROTATION: return S->f.rot;
    case EVP_FIELD_GAMMA_ACC: return S->f.gacc;
    case EVP_FIELD_TWIN_FRACTION: return S->f.twinf;
    case EVP_FIELD_STRAIN_INCR: return S->f.de;
    case EVP_FIELD_LOCAL_ROTATION: return S->f.wrot;
    case EVP_FIELD_TWINNED: *el = sizeof(int32_t); return S->f.twinned;
    case EVP_FIELD_GRAIN: *el = sizeof(int32_t); return S->f.grain;
    case EVP_FIELD_PHASE: *el = sizeof(int32_t); return S->f.phase;
    default: return nullptr;
  }
}

int nccl_check(evp_handle h, int rc, const char *what) {
  if (rc == 0) return EVP_OK;
  return fail(h, EVP_ERR_DEVICE, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"));
}

// all-to-all of equal contiguous pieces (FFT transpose): grouped ncclSend/ncclRecv
int all_to_all_on(evp_handle h, void *comm, int np, size_t piece_elems, const double2 *send, double2 *recv, cudaStream_t st) {
  const size_t piece = piece_elems * sizeof(double2);
  int rc = g_nccl.GroupStart();
  for (int p = 0; p < np && rc == 0; ++p) {
    rc = g_nccl.Send((const char *)send + (size_t)p * piece, piece, kNcclChar, p, comm, st);
    if (rc == 0) rc = g_nccl.Recv((char *)recv + (size_t)p * piece, piece, kNcclChar, p, comm, st);
  }
  const int rc2 = g_nccl.GroupEnd();
  return nccl_check(h, rc ? rc : rc2, "nccl all-to-all");
}
// slab: y <-> z transpose among all ranks
int all_to_all(evp_handle h, const double2 *send, double2 *recv, cudaStream_t st) {
  return all_to_all_on(h, h->comm, h->nranks, (size_t)h->Lsplit.dstride, send, recv, st);
}

// kernel timers: event pairs around every launch, summed per kernel type in fetch_report
void tbeg(evp_handle h, int type, cudaStream_t st) {
  if (!(h->flags & 1) || !h->tm.made || h->tm.n >= 128) return;
  h->tm.type[h->tm.n] = type;
  h->tm.s[h->tm.n] = st;
  cudaEventRecord(h->tm.a[h->tm.n], st);
}
void tend(evp_handle h) {
  if (!(h->flags & 1) || !h->tm.made || h->tm.n >= 128) return;
  cudaEventRecord(h->tm.b[h->tm.n], h->tm.s[h->tm.n]);
  h->tm.n += 1;
}

// K2 + K3 of chunk i, then (ranks > 1) its forward all-to-all on the communication stream
int enqueue_forward_chunk(evp_handle h, int i, const double *field = nullptr) {
  evp_solver::Chunk &c = h->ch[i];
  const int nrows = h->nyb * h->nzc;
  tbeg(h, 0, h->st);
  launch_xfwd(h->nx, field ? field : h->f.sig, c.WB, h->N, c.rowbase, nrows, h->Lplain, h->twx, h->st);
  tend(h);
  if (h->p2p) {
    // y pass + forward transpose in one kernel (TMA stores into the peers' receive buffers), on the communication
    // stream: NVLink-bound, it runs under the constitutive kernel of the next chunk
    cudaEventRecord(c.ev_fwd, h->st);
    cudaStreamWaitEvent(h->stc, c.ev_fwd, 0);
    tbeg(h, 1, h->stc);
    launch_ypass(h->ny, false, c.in_fwd, false, c.out_fwd, true, h->ti_y_plain, h->ti_y_split, h->nxv, h->nzc, h->twy, h->stc);
    tend(h);
    return EVP_OK;
  }
  tbeg(h, 1, h->st);
  launch_ypass(h->ny, false, c.in_fwd, false, c.out_fwd, false, h->ti_y_plain, h->ti_y_split, h->nxv, h->nzc, h->twy, h->st);
  tend(h);
  if (h->nranks > 1) {
    cudaEventRecord(c.ev_fwd, h->st);
    cudaStreamWaitEvent(h->stc, c.ev_fwd, 0);
    tbeg(h, 6, h->stc);
    int rc = all_to_all(h, c.WA, c.WB, h->stc);
    tend(h);
    if (rc) return rc;
    cudaEventRecord(c.ev_a1, h->stc);
  }
  return EVP_OK;
}

// cross-GPU barrier on a stream: everything the ranks enqueued before it (peer stores included) has completed when it returns
int enqueue_barrier(evp_handle h, void *comm, cudaStream_t st) {
  const int rc = g_nccl.AllReduce(h->d_bar, h->d_bar, 1, kNcclDouble, kNcclSum, comm, st);
  return rc ? nccl_check(h, rc, "nccl barrier") : EVP_OK;
}

// K4 over all chunks (needs every forward exchange), then the way-back all-to-all of every chunk
int enqueue_z_and_back(evp_handle h, int zmode = 0) {
  if (h->p2p) {
    // all forward transposes (every rank's y-pass stores) done -> z pass
    tbeg(h, 6, h->stc);
    int rc = enqueue_barrier(h, h->comm, h->stc);
    tend(h);
    if (rc) return rc;
    cudaEventRecord(h->ev_b1, h->stc);
    cudaStreamWaitEvent(h->st, h->ev_b1, 0);
    tbeg(h, 2, h->st);
    // pull: the z pass works in place on the receive buffer; push: its TMA stores ARE the way-back transpose
    launch_zfused(h->nz, zmode, (h->flags & 4) != 0, h->zmaps, h->zout, !h->pull, h->lg_nzl, h->lg_nzc_z, h->zrun_z, h->nxv, h->kx0, h->nyl, h->ky0, h->nx,
                  h->ny, h->g.dx, h->g.dy, h->g.dz, h->twz, h->st);
    tend(h);
    tbeg(h, 6, h->st);
    rc = enqueue_barrier(h, h->comm2 ? h->comm2 : h->comm, h->st);
    tend(h);
    if (rc == 0 && h->pull) {
      // way back: the inverse y pass of every chunk TMA-loads its rows out of the peers' buffers, on the pull stream, so that
      // the transpose of chunk i+1 runs under the x pass / Newton kernel of chunk i
      cudaEventRecord(h->ev_z, h->st);
      cudaStreamWaitEvent(h->stp, h->ev_z, 0);
      for (int i = 0; i < h->nchunks; ++i) {
        evp_solver::Chunk &c = h->ch[i];
        tbeg(h, 3, h->stp);
        launch_ypass(h->ny, true, c.in_pull, true, c.out_inv, false, h->ti_y_split, h->ti_y_plain, h->nxv, h->nzc, h->twy, h->stp);
        tend(h);
        cudaEventRecord(c.ev_pull, h->stp);
      }
    }
    h->green_inflight = true;
    return rc;
  }
  if (h->nranks > 1)
    for (int i = 0; i < h->nchunks; ++i) cudaStreamWaitEvent(h->st, h->ch[i].ev_a1, 0);
  tbeg(h, 2, h->st);
  launch_zfused(h->nz, zmode, (h->flags & 4) != 0, h->zmaps, h->zout, false, h->lg_nzl, h->lg_nzc, h->zrun, h->nxv, h->kx0, h->nyl, h->ky0, h->nx,
                h->ny, h->g.dx, h->g.dy, h->g.dz, h->twz, h->st);
  tend(h);
  if (h->nranks > 1) {
    cudaEventRecord(h->ev_k4, h->st);
    cudaStreamWaitEvent(h->stc, h->ev_k4, 0);
    for (int i = 0; i < h->nchunks; ++i) {
      tbeg(h, 6, h->stc);
      int rc = all_to_all(h, h->ch[i].WB, h->ch[i].WA, h->stc);
      tend(h);
      if (rc) return rc;
      cudaEventRecord(h->ch[i].ev_a2, h->stc);
    }
  }
  h->green_inflight = true;
  return EVP_OK;
}

// K5 + K6 of chunk i (after its way-back exchange has landed)
int enqueue_back_chunk(evp_handle h, int i, bool plain = false) {
  evp_solver::Chunk &c = h->ch[i];
  if (h->nranks > 1 && !h->p2p) cudaStreamWaitEvent(h->st, c.ev_a2, 0);
  if (h->pull) {
    cudaStreamWaitEvent(h->st, c.ev_pull, 0);     // K5 of this chunk ran on the pull stream (enqueue_z_and_back)
  } else {
    tbeg(h, 3, h->st);
    launch_ypass(h->ny, true, c.in_inv, false, c.out_inv, false, h->ti_y_split, h->ti_y_plain, h->nxv, h->nzc, h->twy, h->st);
    tend(h);
  }
  tbeg(h, 4, h->st);
  launch_xinv(h->nx, c.WB, plain ? nullptr : h->f.e, (plain || (h->flags & 2)) ? h->f.de : nullptr, h->d_macro, h->N, c.rowbase,
              h->nyb * h->nzc, h->Lplain, h->twx, h->st);
  tend(h);
  return EVP_OK;
}

// ---- pencil decomposition (py x pz process grid, NCCL transport) ---------------------------------------------------------
// Per z-chunk i the spectrum moves   K2 -> WA_i (x layout) -> [row a2a] -> WB_i -> K3 -> WA_i -> [column a2a] -> WB_i -> K4 in place
//   -> [column a2a] -> WA_i -> K5 -> WB_i -> [row a2a] -> WA_i (x layout) -> K6.   Exchanges run on the communication stream.
int pencil_row_a2a(evp_handle h, const double2 *send, double2 *recv) {
  return all_to_all_on(h, h->comm_row, h->py, (size_t)h->Lplain.dstride, send, recv, h->stc);
}
int pencil_col_a2a(evp_handle h, const double2 *send, double2 *recv) {
  return all_to_all_on(h, h->comm_col, h->pz, (size_t)h->Lsplit.dstride, send, recv, h->stc);
}
// K2 of chunk i + its row exchange
int pencil_x_forward(evp_handle h, int i, const double *field) {
  evp_solver::Chunk &c = h->ch[i];
  tbeg(h, 0, h->st);
  launch_xfwd(h->nx, field ? field : h->f.sig, c.WA, h->N, c.rowbase, h->nyb * h->nzc, h->Lplain, h->twx, h->st);
  tend(h);
  cudaEventRecord(c.ev_p[0], h->st);
  cudaStreamWaitEvent(h->stc, c.ev_p[0], 0);
  tbeg(h, 6, h->stc);
  const int rc = pencil_row_a2a(h, c.WA, c.WB);
  tend(h);
  cudaEventRecord(c.ev_q[1], h->stc);
  return rc;
}
// K3 of chunk i + its column exchange
int pencil_y_forward(evp_handle h, int i) {
  evp_solver::Chunk &c = h->ch[i];
  cudaStreamWaitEvent(h->st, c.ev_q[1], 0);
  tbeg(h, 1, h->st);
  launch_ypass(h->ny, false, c.in_fwd, false, c.out_fwd, false, h->ti_y_plain, h->ti_y_split, h->nxv, h->nzc, h->twy, h->st);
  tend(h);
  cudaEventRecord(c.ev_p[1], h->st);
  cudaStreamWaitEvent(h->stc, c.ev_p[1], 0);
  tbeg(h, 6, h->stc);
  const int rc = pencil_col_a2a(h, c.WA, c.WB);
  tend(h);
  cudaEventRecord(c.ev_a1, h->stc);
  return rc;
}
// K4 over all chunks, then the column exchange back per chunk
int pencil_z_and_back(evp_handle h, int zmode) {
  for (int i = 0; i < h->nchunks; ++i) cudaStreamWaitEvent(h->st, h->ch[i].ev_a1, 0);
  tbeg(h, 2, h->st);
  launch_zfused(h->nz, zmode, (h->flags & 4) != 0, h->zmaps, h->zout, false, h->lg_nzl, h->lg_nzc, h->zrun, h->nxv, h->kx0, h->nyl, h->ky0, h->nx,
                h->ny, h->g.dx, h->g.dy, h->g.dz, h->twz, h->st);
  tend(h);
  cudaEventRecord(h->ev_k4, h->st);
  cudaStreamWaitEvent(h->stc, h->ev_k4, 0);
  for (int i = 0; i < h->nchunks; ++i) {
    tbeg(h, 6, h->stc);
    const int rc = pencil_col_a2a(h, h->ch[i].WB, h->ch[i].WA);
    tend(h);
    if (rc) return rc;
    cudaEventRecord(h->ch[i].ev_a2, h->stc);
  }
  h->green_inflight = true;
  return EVP_OK;
}
// K5 of chunk i + its row exchange back
int pencil_y_back(evp_handle h, int i) {
  evp_solver::Chunk &c = h->ch[i];
  cudaStreamWaitEvent(h->st, c.ev_a2, 0);
  tbeg(h, 3, h->st);
  launch_ypass(h->ny, true, c.in_inv, false, c.out_inv, false, h->ti_y_split, h->ti_y_plain, h->nxv, h->nzc, h->twy, h->st);
  tend(h);
  cudaEventRecord(c.ev_p[2], h->st);
  cudaStreamWaitEvent(h->stc, c.ev_p[2], 0);
  tbeg(h, 6, h->stc);
  const int rc = pencil_row_a2a(h, c.WB, c.WA);
  tend(h);
  cudaEventRecord(c.ev_q[0], h->stc);
  return rc;
}
// K6 of chunk i
void pencil_x_back(evp_handle h, int i, bool plain) {
  evp_solver::Chunk &c = h->ch[i];
  cudaStreamWaitEvent(h->st, c.ev_q[0], 0);
  tbeg(h, 4, h->st);
  launch_xinv(h->nx, c.WA, plain ? nullptr : h->f.e, (plain || (h->flags & 2)) ? h->f.de : nullptr, h->d_macro, h->N, c.rowbase,
              h->nyb * h->nzc, h->Lplain, h->twx, h->st);
  tend(h);
}
int pencil_forward_all(evp_handle h, const double *field, int zmode) {
  int rc = EVP_OK;
  for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = pencil_x_forward(h, i, field);
  for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = pencil_y_forward(h, i);
  return rc ? rc : pencil_z_and_back(h, zmode);
}
int pencil_back_all(evp_handle h, bool plain) {
  int rc = EVP_OK;
  for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = pencil_y_back(h, i);
  for (int i = 0; i < h->nchunks && rc == 0; ++i) pencil_x_back(h, i, plain);
  return rc;
}

void enqueue_const_chunk(evp_handle h, int i) {
  tbeg(h, 5, h->st);
  launch_constitutive(h->f, h->ch[i].vbase, h->ch[i].count, h->nsmax, h->nphases, h->uniform_ns, h->uniform_npow, h->any_twin ? 1 : 0, h->k1_fast,
                      h->d_partials, h->st);
  tend(h);
}

// rows a6 (second stage) + a7
int enqueue_reduce_macro(evp_handle h) {
  launch_reduce(h->d_partials, h->N, h->d_scratch, h->d_totals, h->st);
  if (h->nranks > 1) {
    void *cm = h->comm2 ? h->comm2 : h->comm;
    int rc = g_nccl.AllReduce(h->d_totals, h->d_totals, 11, kNcclDouble, kNcclSum, cm, h->st);
    if (rc == 0) rc = g_nccl.AllReduce(h->d_totals + 11, h->d_totals + 11, 1, kNcclDouble, kNcclMax, cm, h->st);
    if (rc) return nccl_check(h, rc, "nccl allreduce");
  }
  launch_macro(h->d_totals, h->d_macro, h->Ntot, h->st);
  return EVP_OK;
}

// the stress was changed from outside / buffers get reused: drop whatever transform is in flight
void invalidate_green(evp_handle h) {
  if (h->green_inflight && h->stc) cudaStreamSynchronize(h->stc);
  if (h->green_inflight && h->stp) cudaStreamSynchronize(h->stp);
  h->green_inflight = false;
}

int enqueue_forward_all(evp_handle h, const double *field = nullptr, int zmode = 0) {
  if (h->py > 1) return pencil_forward_all(h, field, zmode);
  for (int i = 0; i < h->nchunks; ++i) {
    int rc = enqueue_forward_chunk(h, i, field);
    if (rc) return rc;
  }
  return enqueue_z_and_back(h, zmode);
}

// commit step: local rotation field (w32,w13,w21) of the compatible strain e into f.de[0..2] (same FFT chain, z mode 2)
int enqueue_rotation_field(evp_handle h) {
  invalidate_green(h);   // the chain reuses the spectral buffers
  int rc = enqueue_forward_all(h, h->f.e, 2);
  if (h->py > 1) { if (rc == 0) rc = pencil_back_all(h, true); }
  else for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = enqueue_back_chunk(h, i, true);
  h->green_inflight = false;
  return rc;
}

// rows a1+a2+a3 (unit-test entry): finish the Green step for the current stress
int enqueue_green(evp_handle h) {
  int rc = EVP_OK;
  if (!h->green_inflight) rc = enqueue_forward_all(h);
  if (h->py > 1) { if (rc == 0) rc = pencil_back_all(h, false); }
  else for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = enqueue_back_chunk(h, i);
  h->green_inflight = false;
  return rc;
}

// rows a4+a5+a6+a7 (unit-test entry)
int enqueue_constitutive(evp_handle h) {
  for (int i = 0; i < h->nchunks; ++i) enqueue_const_chunk(h, i);
  return enqueue_reduce_macro(h);
}

// One iteration as the solver runs it.  Per chunk: way back (K5, K6), Newton (K1) and at once the forward
// transform of the NEW stress (K2, K3, forward exchange), so the exchanges of chunk i overlap the kernels of
// chunk i+1; then the reductions, and the z pass + way-back exchange that the next iteration will consume.
int enqueue_iteration(evp_handle h) {
  int rc = EVP_OK;
  if ((h->flags & 1) && h->tm.made) { h->tm.n = 0; cudaEventRecord(h->ev_it0, h->st); }
  if (!h->green_inflight) rc = enqueue_forward_all(h);
  if (h->py > 1) {
    // pencils: the row exchange of chunk i runs under K5 / K6 / K1 / K2 of its neighbours, the column exchange under K3
    for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = pencil_y_back(h, i);
    for (int i = 0; i < h->nchunks && rc == 0; ++i) {
      pencil_x_back(h, i, false);
      enqueue_const_chunk(h, i);
      rc = pencil_x_forward(h, i, nullptr);
    }
    if (rc == 0) rc = enqueue_reduce_macro(h);
    for (int i = 0; i < h->nchunks && rc == 0; ++i) rc = pencil_y_forward(h, i);
    if (rc == 0) rc = pencil_z_and_back(h, 0);
    if ((h->flags & 1) && h->tm.made) cudaEventRecord(h->ev_it1, h->st);
    return rc;
  }
  for (int i = 0; i < h->nchunks && rc == 0; ++i) {
    rc = enqueue_back_chunk(h, i);
    if (rc) break;
    enqueue_const_chunk(h, i);
    rc = enqueue_forward_chunk(h, i);
  }
  if (rc == 0) rc = enqueue_reduce_macro(h);
  if (rc == 0) rc = enqueue_z_and_back(h);
  if ((h->flags & 1) && h->tm.made) cudaEventRecord(h->ev_it1, h->st);
  return rc;
}

int fetch_report(evp_handle h, evp_iter_report *rep) {
  CUDA_OK(h, cudaMemcpyAsync(h->h_macro, h->d_macro, sizeof(MacroDev), cudaMemcpyDeviceToHost, h->st));
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  const MacroDev &m = *h->h_macro;
  if (rep) {
    rep->iter = m.iter;
    rep->newton_max = m.newton_max;
    rep->newton_mean = m.newton_mean;
    rep->err_stress = m.err_s;
    rep->err_strain = m.err_e;
    for (int c = 0; c < 6; ++c) { rep->savg[c] = m.savg[c]; rep->emacro[c] = m.E[c]; }
    rep->converged = (m.iter >= h->ctrl.itmin && m.err_s <= h->ctrl.tol_stress && m.err_e <= h->ctrl.tol_strain && m.unconverged == 0) ? 1 : 0;
    rep->nonfinite = m.nonfinite;
    rep->unconverged = m.unconverged;
  }
  if ((h->flags & 1) && h->tm.made) {
    if (h->stc) cudaStreamSynchronize(h->stc);
    if (h->stp) cudaStreamSynchronize(h->stp);
    for (int i = 0; i < 8; ++i) h->last_ms[i] = 0;
    float ms;
    for (int i = 0; i < h->tm.n; ++i)
      if (cudaEventElapsedTime(&ms, h->tm.a[i], h->tm.b[i]) == cudaSuccess) h->last_ms[h->tm.type[i]] += ms;
    if (cudaEventElapsedTime(&ms, h->ev_it0, h->ev_it1) == cudaSuccess) h->last_ms[7] = ms;
  }
  return m.nonfinite ? fail(h, EVP_ERR_NUMERIC, "Newton produced a non-finite value / bad pivot") : EVP_OK;
}

// ---- host <-> device field transfers through pinned staging --------------------------------------------------------------
// Caller buffers are pageable; a plain cudaMemcpy of pageable memory runs at ~5 GB/s on these hosts (single-threaded driver
// staging).  Here: two pinned buffers, host-side copies split over a few threads, DMA of one buffer under the host copy of the
// other.  (std::thread, not OpenMP: torchrun exports OMP_NUM_THREADS=1.)
constexpr size_t kStageBytes = (size_t)32 << 20;
int stage_threads() {
  static int n = 0;
  if (!n) {
    const unsigned hc = std::thread::hardware_concurrency();
    n = (int)std::max(1u, std::min(8u, hc ? hc : 4u));
    if (const char *e = getenv("EVP_STAGE_THREADS")) n = std::max(1, atoi(e));
  }
  return n;
}
void par_copy(char *dst, const char *src, size_t n) {
  const int nt = (n < ((size_t)4 << 20)) ? 1 : stage_threads();
  if (nt == 1) { std::memcpy(dst, src, n); return; }
  std::vector<std::thread> th;
  const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
  for (int t = 0; t < nt; ++t) {
    const size_t a = (size_t)t * per, b = std::min(n, a + per);
    if (a >= b) break;
    th.emplace_back([=]() { std::memcpy(dst + a, src + a, b - a); });
  }
  for (auto &t : th) t.join();
}
int stage_init(evp_handle h) {
  for (int b = 0; b < 2; ++b) {
    if (!h->stage[b]) CUDA_OK(h, cudaMallocHost(&h->stage[b], kStageBytes));
    if (!h->stage_ev[b]) CUDA_OK(h, cudaEventCreateWithFlags(&h->stage_ev[b], cudaEventDisableTiming));
  }
  return EVP_OK;
}
// enqueues on h->st; the caller synchronises the stream
int copy_h2d(evp_handle h, void *dst, const void *src, size_t bytes) {
  if (bytes < ((size_t)1 << 20)) { CUDA_OK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->st)); return EVP_OK; }
  int rc = stage_init(h);
  if (rc) return rc;
  int k = 0;
  for (size_t off = 0; off < bytes; off += kStageBytes, ++k) {
    const int b = k & 1;
    const size_t n = std::min(kStageBytes, bytes - off);
    CUDA_OK(h, cudaEventSynchronize(h->stage_ev[b]));   // the DMA that last read this buffer is done
    par_copy(h->stage[b], (const char *)src + off, n);
    CUDA_OK(h, cudaMemcpyAsync((char *)dst + off, h->stage[b], n, cudaMemcpyHostToDevice, h->st));
    CUDA_OK(h, cudaEventRecord(h->stage_ev[b], h->st));
  }
  return EVP_OK;
}
// complete on return
int copy_d2h(evp_handle h, void *dst, const void *src, size_t bytes) {
  if (bytes < ((size_t)1 << 20)) {
    CUDA_OK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st));
    CUDA_OK(h, cudaStreamSynchronize(h->st));
    return EVP_OK;
  }
  int rc = stage_init(h);
  if (rc) return rc;
  const size_t nchunk = (bytes + kStageBytes - 1) / kStageBytes;
  auto issue = [&](size_t k) -> cudaError_t {
    const size_t off = k * kStageBytes, n = std::min(kStageBytes, bytes - off);
    cudaError_t e = cudaMemcpyAsync(h->stage[k & 1], (const char *)src + off, n, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaEventRecord(h->stage_ev[k & 1], h->st);
    return e;
  };
  CUDA_OK(h, issue(0));
  for (size_t k = 0; k < nchunk; ++k) {
    if (k + 1 < nchunk) CUDA_OK(h, issue(k + 1));     // the other buffer was copied out in the previous round
    CUDA_OK(h, cudaEventSynchronize(h->stage_ev[k & 1]));
    const size_t off = k * kStageBytes, n = std::min(kStageBytes, bytes - off);
    par_copy((char *)dst + off, h->stage[k & 1], n);
  }
  return EVP_OK;
}

evp_handle g_active = nullptr;  // handle whose tables currently sit in __constant__ memory

// every voxel becomes its own orientation class (after the texture has been changed per voxel)
int switch_to_voxel_classes(evp_handle h) {
  const long long N = h->N;
  if (h->f.norient != N) {
    cudaFree(h->f.mrot); cudaFree(h->f.jb); cudaFree(h->f.orient_rep);
    h->f.mrot = h->f.jb = nullptr; h->f.orient_rep = nullptr;
    CUDA_OK(h, cudaMalloc(&h->f.mrot, sizeof(double) * 25 * N));
    CUDA_OK(h, cudaMalloc(&h->f.jb, sizeof(double) * 21 * N));
    CUDA_OK(h, cudaMalloc(&h->f.orient_rep, sizeof(long long) * N));
    h->f.norient = N;
  }
  launch_voxel_classes(h->f, h->st);
  return EVP_OK;
}

int upload_const(evp_handle h) {
  h->cp.dt = h->dt;
  h->cp.tol_newton = h->ctrl.tol_newton;
  h->cp.newton_itmax = h->ctrl.newton_itmax;
  fill_uniform_rate(h->phd[0], h->cp);
  upload_const_params(h->cp);
  return EVP_OK;
}

// __constant__ tables are per process: re-upload when another handle ran last
void activate(evp_handle h) {
  cudaSetDevice(h->device);
  if (g_active == h) return;
  upload_phase_tables(h->phd.data(), h->nphases);
  if (h->have_c0) upload_green(h->green);
  upload_const(h);
  g_active = h;
}

}  // namespace

extern "C" {

int evp_abi_version(void) { return EVP_ABI_VERSION; }
const char *evp_backend(void) { return "cuda-sm100a"; }
const char *evp_last_error(evp_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int evp_nccl_unique_id(uint8_t id[128]) {
  std::string e;
  if (!load_nccl(&e)) return fail(nullptr, EVP_ERR_DEVICE, e);
  return g_nccl.GetUniqueId(id) == 0 ? EVP_OK : EVP_ERR_DEVICE;
}

int evp_create(const evp_grid *grid, const evp_phase *phases, int32_t nphases, const evp_dist *dist, evp_handle *out) {
  if (!grid || !phases || !out || nphases < 1 || nphases > EVP_MAX_PHASES) return fail(nullptr, EVP_ERR_ARG, "evp_create: bad argument");
  const int nranks = dist ? dist->nranks : 1, rank = dist ? dist->rank : 0, device = dist ? dist->device : 0;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(nullptr, EVP_ERR_ARG, "evp_create: bad rank/nranks");
  if (!fft_size_supported(grid->nx) || !fft_size_supported(grid->ny) || !fft_size_supported(grid->nz))
    return fail(nullptr, EVP_ERR_UNSUPPORTED, "grid sizes must be powers of two in [8, 1024]");
  const int py = (dist && dist->py > 1) ? dist->py : 1;
  if (nranks % py) return fail(nullptr, EVP_ERR_ARG, "evp_create: py must divide nranks");
  const int pz = nranks / py;
  if (grid->nz % pz || grid->ny % pz || grid->ny % py) return fail(nullptr, EVP_ERR_UNSUPPORTED, "ny must be divisible by py and pz, nz by pz");
  if (g_device >= 0 && g_device != device)
    return fail(nullptr, EVP_ERR_UNSUPPORTED, "this process already drives CUDA device " + std::to_string(g_device) + ": one GPU per process (see evpfft.h)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device)
    return fail(nullptr, EVP_ERR_DEVICE, "no CUDA device: this library has no CPU fallback");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, EVP_ERR_DEVICE, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, EVP_ERR_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", library is built for sm_100a only");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, EVP_ERR_DEVICE, "cudaSetDevice failed");
  g_device = device;

  evp_solver *S = new evp_solver();
  evp_handle h = S;
  S->g = *grid;
  if (!(S->g.dx > 0)) S->g.dx = 1.0;
  if (!(S->g.dy > 0)) S->g.dy = 1.0;
  if (!(S->g.dz > 0)) S->g.dz = 1.0;
  S->nx = grid->nx; S->ny = grid->ny; S->nz = grid->nz;
  S->nxh = S->nx / 2 + 1;
  S->nranks = nranks; S->rank = rank; S->device = device;
  S->py = py; S->pz = pz; S->iy = rank / pz; S->iz = rank % pz;
  S->kxl = ((S->nxh + py - 1) / py + 7) / 8 * 8;          // kx columns per row rank (pitch of every spectral row)
  S->nxp = S->kxl;
  S->kx0 = S->iy * S->kxl;
  S->nxv = std::max(0, std::min(S->kxl, S->nxh - S->kx0));
  S->nzl = S->nz / pz; S->z0 = S->iz * S->nzl;
  S->nyl = S->ny / pz; S->ky0 = S->iz * S->nyl;
  S->nyb = S->ny / py; S->y0 = S->iy * S->nyb;
  S->N = (long long)S->nx * S->nyb * S->nzl;
  S->Ntot = (double)S->nx * S->ny * S->nz;
  S->nphases = nphases;
  S->ph.assign(phases, phases + nphases);
  S->phd.resize(nphases);
  bool any_twin = false;
  for (int p = 0; p < nphases; ++p) {
    if (phases[p].nsys < 0 || phases[p].nsys > EVP_MAX_SYS || phases[p].nmodes < 0 || phases[p].nmodes > EVP_MAX_MODES) {
      delete S;
      return fail(nullptr, EVP_ERR_ARG, "phase: nsys/nmodes out of range");
    }
    build_phase_dev(phases[p], S->phd[p]);
    S->nsmax = std::max(S->nsmax, (int)phases[p].nsys);
    for (int q = 0; q < phases[p].nsys; ++q) {
      const int np = S->phd[p].npow[q];
      if (p == 0 && q == 0) S->uniform_npow = np;
      else if (np != S->uniform_npow) S->uniform_npow = -2;
    }
    if (p == 0) S->uniform_ns = phases[p].nsys;
    else if (phases[p].nsys != S->uniform_ns) S->uniform_ns = 0;
    for (int m = 0; m < phases[p].nmodes; ++m) any_twin |= phases[p].twin[m] != 0;
  }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
      evp_destroy(S);                                                                              \
      return fail(nullptr, EVP_ERR_DEVICE, m_);                                                    \
    }                                                                                              \
  } while (0)
  CK(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
  const long long N = S->N;
  const int ns = std::max(S->nsmax, 1);
  S->f.N = N;
  CK(cudaMalloc(&S->f.sig, sizeof(double) * 6 * N));
  CK(cudaMalloc(&S->f.e, sizeof(double) * 6 * N));
  CK(cudaMalloc(&S->f.epsp, sizeof(double) * 6 * N));
  CK(cudaMalloc(&S->f.edotp, sizeof(double) * 6 * N));
  CK(cudaMalloc(&S->f.crss, sizeof(double) * ns * N));
  CK(cudaMalloc(&S->f.rot, sizeof(double) * 9 * N));
  CK(cudaMalloc(&S->f.gacc, sizeof(double) * N));
  CK(cudaMalloc(&S->f.orient, sizeof(int32_t) * N));
  CK(cudaMalloc(&S->f.wrot, sizeof(double) * 3 * N));
  CK(cudaMalloc(&S->f.twinned, sizeof(int32_t) * N));
  CK(cudaMalloc(&S->f.itc, sizeof(double) * ns * N));
  S->any_twin = any_twin;
  if (any_twin) CK(cudaMalloc(&S->f.twinf, sizeof(double) * ns * N));
  CK(cudaMalloc(&S->f.grain, sizeof(int32_t) * N));
  CK(cudaMalloc(&S->f.phase, sizeof(int32_t) * N));
  // spectral work buffers (half spectrum, 6 components)
  const size_t wbytes = sizeof(double2) * 6 * (size_t)S->nzl * S->ny * S->nxp;
  CK(cudaMalloc(&S->WA, wbytes));
  CK(cudaMemsetAsync(S->WA, 0, wbytes, S->st));
  if (nranks > 1) {
    CK(cudaMalloc(&S->WB, wbytes));
    CK(cudaMemsetAsync(S->WB, 0, wbytes, S->st));
  } else {
    S->WB = S->WA;
  }
  if (nranks > 1) {
    std::string e;
    if (!load_nccl(&e)) { evp_destroy(S); return fail(nullptr, EVP_ERR_DEVICE, e); }
    Id128 id;
    std::memcpy(id.b, dist->nccl_id, 128);
    const int rc = g_nccl.CommInitRank(&S->comm, nranks, id, rank);
    if (rc) { evp_destroy(S); return fail(nullptr, EVP_ERR_DEVICE, "ncclCommInitRank failed"); }
    // separate communicator for the tiny norm all-reduces so that they do not queue behind the transposes
    if (g_nccl.CommSplit && g_nccl.CommSplit(S->comm, 0, rank, &S->comm2, nullptr) != 0) S->comm2 = nullptr;
    if (py > 1) {
      if (!g_nccl.CommSplit || g_nccl.CommSplit(S->comm, S->iz, S->iy, &S->comm_row, nullptr) != 0 ||
          g_nccl.CommSplit(S->comm, S->iy, S->iz, &S->comm_col, nullptr) != 0) {
        evp_destroy(S);
        return fail(nullptr, EVP_ERR_DEVICE, "ncclCommSplit (pencil row / column communicators) failed");
      }
    }
  }
  // transport of the FFT transposes (evp_transport_kind); pencils exchange through NCCL only
  if (nranks > 1) {
    int tr = dist->transport;
    if (const char *e = getenv("EVP_TRANSPORT")) tr = (std::string(e) == "nccl") ? EVP_TRANSPORT_NCCL : (std::string(e) == "p2p" ? EVP_TRANSPORT_P2P : tr);
    if (py > 1) {
      if (tr == EVP_TRANSPORT_P2P) { evp_destroy(S); return fail(nullptr, EVP_ERR_UNSUPPORTED, "pencil decomposition supports the NCCL transport only"); }
      tr = EVP_TRANSPORT_NCCL;
    }
    if (tr != EVP_TRANSPORT_NCCL) {
      std::string why;
      if (nranks > kMaxRanks) why = "too many ranks";
      if (why.empty() && cudaMalloc(&S->WC, wbytes) != cudaSuccess) why = "cudaMalloc(WC) failed";
      if (why.empty() && !g_nccl.AllGather) why = "ncclAllGather missing";
      if (why.empty()) {
        cudaMemsetAsync(S->WC, 0, wbytes, S->st);
        cudaIpcMemHandle_t mine[2];
        std::vector<cudaIpcMemHandle_t> all((size_t)2 * nranks);
        char *d_h = nullptr;
        const size_t hb = sizeof(mine);
        if (cudaIpcGetMemHandle(&mine[0], S->WA) != cudaSuccess || cudaIpcGetMemHandle(&mine[1], S->WC) != cudaSuccess) why = "cudaIpcGetMemHandle failed";
        if (why.empty() && cudaMalloc(&d_h, hb * nranks) != cudaSuccess) why = "cudaMalloc failed";
        if (why.empty()) {
          cudaMemcpyAsync(d_h + hb * rank, mine, hb, cudaMemcpyHostToDevice, S->st);
          if (g_nccl.AllGather(d_h + hb * rank, d_h, hb, kNcclChar, S->comm, S->st) != 0) why = "ncclAllGather failed";
          cudaMemcpyAsync(all.data(), d_h, hb * nranks, cudaMemcpyDeviceToHost, S->st);
          cudaStreamSynchronize(S->st);
          cudaFree(d_h);
        }
        int ok = why.empty() ? 1 : 0;
        for (int p = 0; p < nranks && ok; ++p) {
          if (p == rank) { S->peerWA[p] = S->WA; S->peerWC[p] = S->WC; continue; }
          if (cudaIpcOpenMemHandle((void **)&S->peerWA[p], all[2 * p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
              cudaIpcOpenMemHandle((void **)&S->peerWC[p], all[2 * p + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            ok = 0;
            why = "cudaIpcOpenMemHandle failed";
            cudaGetLastError();
          }
        }
        // every rank must take the same decision
        double flag = ok ? 0.0 : 1.0, *d_f = nullptr;
        cudaMalloc(&d_f, sizeof(double));
        cudaMemcpy(d_f, &flag, sizeof(double), cudaMemcpyHostToDevice);
        g_nccl.AllReduce(d_f, d_f, 1, kNcclDouble, kNcclSum, S->comm, S->st);
        cudaMemcpyAsync(&flag, d_f, sizeof(double), cudaMemcpyDeviceToHost, S->st);
        cudaStreamSynchronize(S->st);
        cudaFree(d_f);
        S->p2p = (flag == 0.0);
        if (!S->p2p && why.empty()) why = "a peer could not map the buffers";
      }
      if (!S->p2p && tr == EVP_TRANSPORT_P2P) { evp_destroy(S); return fail(nullptr, EVP_ERR_DEVICE, "peer-memory transport requested but unavailable: " + why); }
    }
    CK(cudaMalloc(&S->d_bar, sizeof(double)));
    CK(cudaMemset(S->d_bar, 0, sizeof(double)));
    CK(cudaEventCreateWithFlags(&S->ev_b1, cudaEventDisableTiming));
  }
  // pipeline chunks: only with ranks > 1 (nothing to overlap otherwise); chunk voxel counts must be multiples of 128
  {
    S->pull = S->p2p && !(getenv("EVP_WAYBACK") && std::string(getenv("EVP_WAYBACK")) == "push");
    // pull mode: the z pass does not see the chunks.  More chunks shorten the exposed head (first pull) and tail (last push) of
    // the pipeline, but the transposes then time-share the SMs with the Newton kernel in smaller pieces; measured at 2 ranks
    // (256x256x512): 2 / 4 / 8 chunks = 4.27 / 4.33 / 4.38 ms per iteration (profiles/r02_multigpu.md).  With more ranks the
    // transposes are longer (7/8 of the spectrum crosses NVLink at 8 ranks) and the head / tail weigh more: 4 chunks.
    int want = getenv("EVP_CHUNKS") ? atoi(getenv("EVP_CHUNKS")) : ((nranks > 1) ? ((S->pull && nranks == 2) ? 2 : 4) : 1);   // env: also on one rank (tests)
    // push mode: the z pass carries one output tensor map per (destination, chunk) as kernel parameters
    want = std::max(1, std::min(want, (int)((S->p2p && !S->pull) ? kMaxChunksP2P : kMaxChunks)));
    while (want > 1 && (S->nzl % want != 0 || ((long long)(S->nzl / want) * S->nyb * S->nx) % 128 != 0)) want /= 2;
    S->nchunks = want;
    S->nzc = S->nzl / want;
  }
  // "plain" = the x-side view of the y stage: rows grouped by the row rank that owns them in real space (py groups of nyb
  // rows; one group of ny rows for slabs);  "split" = its z-side view: rows grouped by the column rank that transforms them
  S->Lplain.nyl = S->nyb; S->Lplain.nzl = S->nzc; S->Lplain.nxp = S->nxp; S->Lplain.nxh = S->nxh;
  S->Lplain.zstride = (long long)S->nyb * S->nxp;
  S->Lplain.cstride = (long long)S->nzc * S->Lplain.zstride;
  S->Lplain.dstride = 6 * S->Lplain.cstride;
  S->Lsplit.nyl = S->nyl; S->Lsplit.nzl = S->nzc; S->Lsplit.nxp = S->nxp; S->Lsplit.nxh = S->nxh;
  S->Lsplit.zstride = (long long)S->nyl * S->nxp;
  S->Lsplit.cstride = (long long)S->nzc * S->Lsplit.zstride;
  S->Lsplit.dstride = 6 * S->Lsplit.cstride;
  S->Lplain.lg_nyl = ilog2i(S->nyb); S->Lplain.lg_nzl = ilog2i(S->nzc);
  S->Lsplit.lg_nyl = ilog2i(S->nyl); S->Lsplit.lg_nzl = ilog2i(S->nzc);
  S->lg_nzl = ilog2i(S->nzl); S->lg_nzc = ilog2i(S->nzc);
  {
    std::string e;
    const int ycp = std::min(S->nyb, 256), ycs = std::min(S->nyl, 256);
    S->zrun = std::min(S->nzc, 256);
    S->ti_y_plain = {S->Lplain.lg_nyl, ycp};
    S->ti_y_split = {S->Lsplit.lg_nyl, ycs};
    const size_t csize = (size_t)6 * S->nzc * S->ny * S->nxp;   // complex elements per chunk sub-buffer (same in the x, y and z stages)
    for (int i = 0; i < S->nchunks; ++i) {
      evp_solver::Chunk &c = S->ch[i];
      c.rowbase = i * S->nzc * S->nyb;
      c.vbase = (long long)c.rowbase * S->nx;
      c.count = (long long)S->nzc * S->nyb * S->nx;
      c.WA = S->WA + (size_t)i * csize;
      c.WB = S->WB + (size_t)i * csize;
      // K2 -> WB (plain) -> K3 -> WA (split) -> [all-to-all -> WB] -> K4 in place -> [all-to-all -> WA] -> K5 -> WB (plain) -> K6
      // p2p:  K2 -> WB (plain) -> K3 stores into every peer's WC -> K4 reads WC, stores into every peer's WA -> K5 -> WB -> K6
      double2 *Wz = S->p2p ? S->WC + (size_t)i * csize : ((nranks > 1) ? c.WB : c.WA);
      // pencil (NCCL): K2 -> WA (x layout) -> [row a2a -> WB] -> K3 -> WA (split) -> [column a2a -> WB] -> K4 in place -> [column a2a -> WA]
      //                -> K5 -> WB (plain) -> [row a2a -> WA (x layout)] -> K6
      bool ok = make_tmap(&c.tm_y_plain, c.WB, S->Lplain, py, ypass_tx(), ycp, 1, &e) &&
                make_tmap(&c.tm_y_split, c.WA, S->Lsplit, pz, ypass_tx(), ycs, 1, &e);
      // pull mode: the receive buffer WC is laid out for the WHOLE slab ([source rank][c][z_local][ky_local][kx]); the chunks push
      // into / pull from their own z range of it, and the z pass reads runs of up to 256 planes whatever the chunk count
      SpecLayout Lwc = S->Lsplit;
      Lwc.nzl = S->nzl; Lwc.lg_nzl = S->lg_nzl;
      Lwc.cstride = (long long)S->nzl * Lwc.zstride;
      Lwc.dstride = 6 * Lwc.cstride;
      SpecLayout Lwc_chunk = Lwc;          // the same strides, the z extent of one chunk
      Lwc_chunk.nzl = S->nzc;
      const size_t zoff = (size_t)i * S->nzc * Lwc.zstride;
      if (ok && S->pull) {
        S->zrun_z = std::min(S->nzl, 256); S->lg_nzc_z = S->lg_nzl;
        if (i == 0) ok = make_tmap(&S->zmaps.m[0], S->WC, Lwc, pz, zpass_tx(S->nz), 1, S->zrun_z, &e);
        else S->zmaps.m[i] = S->zmaps.m[0];
      } else if (ok) {
        S->zrun_z = S->zrun; S->lg_nzc_z = S->lg_nzc;
        ok = make_tmap(&S->zmaps.m[i], Wz, S->Lsplit, pz, zpass_tx(S->nz), 1, S->zrun, &e);
      }
      for (int p = 0; p < kMaxRanks && ok; ++p) {
        c.out_inv.m[p] = c.tm_y_plain;
        c.in_fwd.m[p] = c.tm_y_plain;
        c.out_fwd.m[p] = c.tm_y_split;
        c.in_inv.m[p] = c.tm_y_split;
        c.in_pull.m[p] = c.tm_y_split;
        if (S->p2p && p < nranks) {
          // pull: the rows rank p transformed along z for my planes sit in p's receive buffer at source slot `rank`
          if (S->pull) {
            ok = make_tmap(&c.in_pull.m[p], S->peerWC[p] + (size_t)rank * Lwc.dstride + zoff, Lwc_chunk, 1, ypass_tx(), ycs, 1, &e) &&
                 make_tmap(&c.out_fwd.m[p], S->peerWC[p] + (size_t)rank * Lwc.dstride + zoff, Lwc_chunk, 1, ypass_tx(), ycs, 1, &e);
            continue;
          }
          // my rows for destination p land in p's receive buffer at source slot `rank`
          ok = make_tmap(&c.out_fwd.m[p], S->peerWC[p] + (size_t)i * csize + (size_t)rank * S->Lsplit.dstride, S->Lsplit, 1, ypass_tx(), ycs, 1, &e);
          // my ky rows of p's planes land in p's way-back buffer at slot `rank`
          if (ok && !S->pull) ok = make_tmap(&S->zout.m[p * kMaxChunksP2P + i], S->peerWA[p] + (size_t)i * csize + (size_t)rank * S->Lsplit.dstride, S->Lsplit,
                                 1, zpass_tx(S->nz), 1, S->zrun, &e);
        }
      }
      if (!ok) {
        evp_destroy(S);
        return fail(nullptr, EVP_ERR_DEVICE, e);
      }
      if (nranks > 1) {
        CK(cudaEventCreateWithFlags(&c.ev_fwd, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.ev_a1, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.ev_a2, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.ev_pull, cudaEventDisableTiming));
        for (cudaEvent_t &ev : c.ev_p) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (cudaEvent_t &ev : c.ev_q) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      }
    }
    for (int i = S->nchunks; i < kMaxChunks; ++i) S->zmaps.m[i] = S->zmaps.m[0];
    CK(cudaEventCreate(&S->ev_it0));
    CK(cudaEventCreate(&S->ev_it1));
    if (nranks > 1) {
      {
        // EVP_COMM_PRIO: 0 = same priority as the compute stream (default), 1 = lower, -1 = higher
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const int mode = getenv("EVP_COMM_PRIO") ? atoi(getenv("EVP_COMM_PRIO")) : 0;
        const int pr = (mode > 0) ? lo : (mode < 0 ? hi : 0);
        CK(cudaStreamCreateWithPriority(&S->stc, cudaStreamNonBlocking, pr));
      }
      CK(cudaEventCreateWithFlags(&S->ev_k4, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&S->ev_z, cudaEventDisableTiming));
      CK(cudaStreamCreateWithFlags(&S->stp, cudaStreamNonBlocking));
    }
  }
  S->twx = make_twiddles(S->nx); S->twy = make_twiddles(S->ny); S->twz = make_twiddles(S->nz);
  if (!S->twx || !S->twy || !S->twz) { evp_destroy(S); return fail(nullptr, EVP_ERR_DEVICE, "twiddle allocation failed"); }
  CK(cudaMalloc(&S->d_macro, sizeof(MacroDev)));
  CK(cudaMemset(S->d_macro, 0, sizeof(MacroDev)));
  CK(cudaMallocHost(&S->h_macro, sizeof(MacroDev)));
  std::memset(S->h_macro, 0, sizeof(MacroDev));
  CK(cudaMalloc(&S->d_partials, sizeof(double) * (size_t)partial_doubles(N)));
  CK(cudaMalloc(&S->d_totals, sizeof(double) * 64));
  CK(cudaMalloc(&S->d_scratch, sizeof(double) * reduce_scratch_doubles()));
  CK(cudaStreamSynchronize(S->st));
#undef CK
  (void)h;
  *out = S;
  return EVP_OK;
}

int evp_destroy(evp_handle h) {
  if (!h) return EVP_OK;
  if (g_active == h) g_active = nullptr;
  cudaSetDevice(h->device);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->stc) cudaStreamSynchronize(h->stc);
  if (h->stp) cudaStreamSynchronize(h->stp);
  if (h->p2p) {
    // the host synchronises the ranks before destroying handles (no collective here: a lone destroy must not hang);
    // every iteration ends with a cross-GPU barrier, so no peer store targets this rank once its own stream is idle
    for (int p = 0; p < h->nranks; ++p)
      if (p != h->rank) { if (h->peerWA[p]) cudaIpcCloseMemHandle(h->peerWA[p]); if (h->peerWC[p]) cudaIpcCloseMemHandle(h->peerWC[p]); }
  }
  cudaFree(h->WC); cudaFree(h->d_bar);
  if (h->ev_b1) cudaEventDestroy(h->ev_b1);
  if (h->comm_row && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm_row);
  if (h->comm_col && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm_col);
  if (h->comm2 && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm2);
  for (int b = 0; b < 2; ++b) {
    if (h->stage[b]) cudaFreeHost(h->stage[b]);
    if (h->stage_ev[b]) cudaEventDestroy(h->stage_ev[b]);
  }
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  cudaFree(h->f.sig); cudaFree(h->f.e); cudaFree(h->f.epsp); cudaFree(h->f.edotp); cudaFree(h->f.crss);
  cudaFree(h->f.mrot); cudaFree(h->f.jb); cudaFree(h->f.itc); cudaFree(h->f.orient); cudaFree(h->f.orient_rep);
  cudaFree(h->f.wrot); cudaFree(h->f.twinned);
  cudaFree(h->f.rot); cudaFree(h->f.gacc); cudaFree(h->f.twinf); cudaFree(h->f.de); cudaFree(h->f.grain); cudaFree(h->f.phase);
  if (h->WB && h->WB != h->WA) cudaFree(h->WB);
  cudaFree(h->WA);
  cudaFree(h->twx); cudaFree(h->twy); cudaFree(h->twz);
  cudaFree(h->d_macro); cudaFree(h->d_partials); cudaFree(h->d_totals); cudaFree(h->d_scratch);
  if (h->h_macro) cudaFreeHost(h->h_macro);
  if (h->tm.made) for (int i = 0; i < 128; ++i) { cudaEventDestroy(h->tm.a[i]); cudaEventDestroy(h->tm.b[i]); }
  for (int i = 0; i < kMaxChunks; ++i) {
    if (h->ch[i].ev_fwd) cudaEventDestroy(h->ch[i].ev_fwd);
    if (h->ch[i].ev_a1) cudaEventDestroy(h->ch[i].ev_a1);
    if (h->ch[i].ev_a2) cudaEventDestroy(h->ch[i].ev_a2);
    if (h->ch[i].ev_pull) cudaEventDestroy(h->ch[i].ev_pull);
    for (cudaEvent_t ev : h->ch[i].ev_p) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->ch[i].ev_q) if (ev) cudaEventDestroy(ev);
  }
  if (h->ev_k4) cudaEventDestroy(h->ev_k4);
  if (h->ev_it0) cudaEventDestroy(h->ev_it0);
  if (h->ev_it1) cudaEventDestroy(h->ev_it1);
  if (h->ev_z) cudaEventDestroy(h->ev_z);
  if (h->stp) cudaStreamDestroy(h->stp);
  if (h->stc) cudaStreamDestroy(h->stc);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return EVP_OK;
}

int evp_local_slab(evp_handle h, int32_t *z0, int32_t *nzl) {
  if (!h) return EVP_ERR_ARG;
  if (z0) *z0 = h->z0;
  if (nzl) *nzl = h->nzl;
  return EVP_OK;
}
int evp_local_block(evp_handle h, int32_t *y0, int32_t *nyl, int32_t *z0, int32_t *nzl) {
  if (!h) return EVP_ERR_ARG;
  if (y0) *y0 = h->y0;
  if (nyl) *nyl = h->nyb;
  if (z0) *z0 = h->z0;
  if (nzl) *nzl = h->nzl;
  return EVP_OK;
}
int evp_nsys_max(evp_handle h) { return h ? h->nsmax : EVP_ERR_ARG; }
int evp_transport(evp_handle h) { return (h && h->p2p) ? EVP_TRANSPORT_P2P : EVP_TRANSPORT_NCCL; }
const char *evp_build_id(void) { return EVP_SRC_HASH; }
int64_t evp_launch_count(evp_handle) { return (int64_t)launch_count(); }
int evp_debug_fp64_peak(evp_handle h, int32_t reps, double *tflops) {
  if (!h || !tflops) return EVP_ERR_ARG;
  cudaSetDevice(h->device);
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  *tflops = measure_fp64_peak(reps, h->st);
  CUDA_OK(h, cudaGetLastError());
  return (*tflops > 0.0) ? EVP_OK : fail(h, EVP_ERR_DEVICE, "fp64 peak measurement failed");
}
void *evp_stream(evp_handle h) { return h ? (void *)h->st : nullptr; }

int evp_set_microstructure(evp_handle h, const int32_t *grain, const int32_t *phase, const double *rot9) {
  if (!h || !grain || !rot9) return fail(h, EVP_ERR_ARG, "set_microstructure: null pointer");
  cudaSetDevice(h->device);
  const long long N = h->N;
  if (phase)
    for (long long v = 0; v < N; ++v)
      if (phase[v] < 0 || phase[v] >= h->nphases) return fail(h, EVP_ERR_ARG, "set_microstructure: phase id out of range");
  // orientation classes: one per grain if every voxel of a grain carries the same rotation (the usual
  // starting texture), else one per voxel.  The per-class invariants (M, Jb) are rebuilt every increment.
  {
    int32_t gmax = -1;
    bool ok = true;
    for (long long v = 0; v < N; ++v) {
      if (grain[v] < 0) { ok = false; break; }
      gmax = std::max(gmax, grain[v]);
    }
    std::vector<long long> rep;
    if (ok && getenv("EVP_ORIENT_PER_VOXEL") == nullptr && (long long)gmax + 1 <= N / 4) {
      rep.assign((size_t)gmax + 1, -1);
      for (long long v = 0; v < N && ok; ++v) {
        long long &r = rep[grain[v]];
        if (r < 0) { r = v; continue; }
        // the class tables (Jb = S0_c + S_c, M) are built from the representative voxel: every voxel of the class must
        // share its rotation AND its phase (a grain id that spans two phases falls back to per-voxel classes)
        if (phase && phase[v] != phase[r]) { ok = false; break; }
        for (int k = 0; k < 9; ++k)
          if (rot9[(size_t)k * N + v] != rot9[(size_t)k * N + r]) { ok = false; break; }
      }
    } else {
      ok = false;
    }
    std::vector<int32_t> oid;
    if (!ok) {  // per-voxel classes
      rep.resize((size_t)N);
      oid.resize((size_t)N);
      for (long long v = 0; v < N; ++v) { rep[v] = v; oid[v] = (int32_t)v; }
    }
    const long long NO = (long long)rep.size();
    if (NO != h->f.norient) {
      cudaFree(h->f.mrot); cudaFree(h->f.jb); cudaFree(h->f.orient_rep);
      h->f.mrot = h->f.jb = nullptr; h->f.orient_rep = nullptr;
      CUDA_OK(h, cudaMalloc(&h->f.mrot, sizeof(double) * 25 * NO));
      CUDA_OK(h, cudaMalloc(&h->f.jb, sizeof(double) * 21 * NO));
      CUDA_OK(h, cudaMalloc(&h->f.orient_rep, sizeof(long long) * NO));
      h->f.norient = NO;
    }
    CUDA_OK(h, cudaMemcpy(h->f.orient_rep, rep.data(), sizeof(long long) * NO, cudaMemcpyHostToDevice));
    { int rc = copy_h2d(h, h->f.orient, ok ? grain : oid.data(), sizeof(int32_t) * N); if (rc) return rc; }
    CUDA_OK(h, cudaStreamSynchronize(h->st));   // oid is a local
  }
  { int rc = copy_h2d(h, h->f.grain, grain, sizeof(int32_t) * N); if (rc) return rc; }
  if (phase) { int rc = copy_h2d(h, h->f.phase, phase, sizeof(int32_t) * N); if (rc) return rc; }
  else CUDA_OK(h, cudaMemsetAsync(h->f.phase, 0, sizeof(int32_t) * N, h->st));
  { int rc = copy_h2d(h, h->f.rot, rot9, sizeof(double) * 9 * N); if (rc) return rc; }
  CUDA_OK(h, cudaMemsetAsync(h->f.sig, 0, sizeof(double) * 6 * N, h->st));
  CUDA_OK(h, cudaMemsetAsync(h->f.e, 0, sizeof(double) * 6 * N, h->st));
  CUDA_OK(h, cudaMemsetAsync(h->f.epsp, 0, sizeof(double) * 6 * N, h->st));
  CUDA_OK(h, cudaMemsetAsync(h->f.edotp, 0, sizeof(double) * 6 * N, h->st));
  CUDA_OK(h, cudaMemsetAsync(h->f.gacc, 0, sizeof(double) * N, h->st));
  CUDA_OK(h, cudaMemsetAsync(h->f.wrot, 0, sizeof(double) * 3 * N, h->st));
  CUDA_OK(h, cudaMemsetAsync(h->f.twinned, 0, sizeof(int32_t) * N, h->st));
  h->ntwinned = 0;
  h->facc = 0.0;
  if (h->f.twinf) CUDA_OK(h, cudaMemsetAsync(h->f.twinf, 0, sizeof(double) * std::max(h->nsmax, 1) * N, h->st));
  if (h->f.de) CUDA_OK(h, cudaMemsetAsync(h->f.de, 0, sizeof(double) * 6 * N, h->st));
  activate(h);
  launch_init_crss(h->f, h->nsmax, h->st);
  {
    MacroDev &m = *h->h_macro;
    for (int c = 0; c < 6; ++c) { m.E[c] = m.Et[c] = m.dEpend[c] = m.savg[c] = m.epavg[c] = 0.0; h->Et[c] = 0; h->Edot_prev[c] = 0; }
    m.err_s = m.err_e = m.newton_mean = 0.0;
    m.newton_max = m.nonfinite = m.iter = 0;
    m.unconverged = 0;
    CUDA_OK(h, cudaMemcpyAsync(h->d_macro, &m, sizeof(MacroDev), cudaMemcpyHostToDevice, h->st));
  }
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  CUDA_OK(h, cudaGetLastError());
  h->have_micro = true; h->in_incr = false;
  invalidate_green(h);
  return EVP_OK;
}

int evp_set_reference_medium(evp_handle h, const double *c0) {
  if (!h) return EVP_ERR_ARG;
  if (c0) {
    voigt_to_mandel(c0, h->C0m);
  } else {
    if (!h->have_micro) return fail(h, EVP_ERR_STATE, "reference medium average needs the microstructure");
    // Voigt average of the rotated crystal stiffnesses: one-off set-up reduction, done on the host
    // from the device-resident orientation field (not on the hot path)
    const long long N = h->N;
    std::vector<double> rot((size_t)9 * N);
    std::vector<int32_t> phs((size_t)N);
    CUDA_OK(h, cudaMemcpy(rot.data(), h->f.rot, sizeof(double) * 9 * N, cudaMemcpyDeviceToHost));
    CUDA_OK(h, cudaMemcpy(phs.data(), h->f.phase, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
    std::vector<double> Cm((size_t)36 * h->nphases);
    for (int p = 0; p < h->nphases; ++p) voigt_to_mandel(h->ph[p].c_voigt, &Cm[36 * p]);
    // fixed blocks summed in block order: the result does not depend on the OpenMP thread count or schedule
    double acc[36] = {0};
    const long long kBlk = 4096, nblk = (N + kBlk - 1) / kBlk;
    std::vector<double> part((size_t)36 * nblk, 0.0);
#pragma omp parallel for schedule(dynamic, 4)
    for (long long b = 0; b < nblk; ++b) {
      double *pa = &part[(size_t)36 * b];
      const long long v1 = std::min(N, (b + 1) * kBlk);
      for (long long v = b * kBlk; v < v1; ++v) {
        double R[9], Q[36], T[36];
        for (int k = 0; k < 9; ++k) R[k] = rot[(size_t)k * N + v];
        mandel_rotation(R, Q);
        const double *C = &Cm[36 * phs[v]];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            double x = 0;
            for (int k = 0; k < 6; ++k) x += Q[6 * i + k] * C[6 * k + j];
            T[6 * i + j] = x;
          }
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            double x = 0;
            for (int k = 0; k < 6; ++k) x += T[6 * i + k] * Q[6 * j + k];
            pa[6 * i + j] += x;
          }
      }
    }
    for (long long b = 0; b < nblk; ++b)
      for (int k = 0; k < 36; ++k) acc[k] += part[(size_t)36 * b + k];
    if (h->nranks > 1) {
      CUDA_OK(h, cudaMemcpyAsync(h->d_totals + 16, acc, sizeof(acc), cudaMemcpyHostToDevice, h->st));
      int rc = g_nccl.AllReduce(h->d_totals + 16, h->d_totals + 16, 36, kNcclDouble, kNcclSum, h->comm, h->st);
      if (rc) return nccl_check(h, rc, "nccl allreduce (C0)");
      CUDA_OK(h, cudaMemcpyAsync(acc, h->d_totals + 16, sizeof(acc), cudaMemcpyDeviceToHost, h->st));
      CUDA_OK(h, cudaStreamSynchronize(h->st));
    }
    for (int k = 0; k < 36; ++k) h->C0m[k] = acc[k] / h->Ntot;
    for (int i = 0; i < 6; ++i)
      for (int j = i + 1; j < 6; ++j) h->C0m[6 * i + j] = h->C0m[6 * j + i] = 0.5 * (h->C0m[6 * i + j] + h->C0m[6 * j + i]);
  }
  if (!inv6(h->C0m, h->S0m)) return fail(h, EVP_ERR_NUMERIC, "reference medium is singular");
  build_s0b(h->S0m, h->cp.S0b, &h->cp.iso_c0);
  GreenConst G;
  build_green_const(h->C0m, h->S0m, G);
  h->green = G;
  h->have_c0 = true;
  invalidate_green(h);   // the Green operator changed
  g_active = nullptr;
  activate(h);
  if (h->have_loading) {  // Mmac depends on C0
    int32_t iu[9], is[6];
    double u[9], s[6];
    std::memcpy(iu, h->iudot, sizeof(iu)); std::memcpy(is, h->iscau, sizeof(is));
    std::memcpy(u, h->udot, sizeof(u)); std::memcpy(s, h->scau, sizeof(s));
    return evp_set_loading(h, iu, u, is, s);
  }
  return EVP_OK;
}

int evp_get_reference_medium(evp_handle h, double *c0) {
  if (!h || !c0 || !h->have_c0) return EVP_ERR_STATE;
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) c0[6 * a + b] = h->C0m[6 * a + b] / (kW[a] * kW[b]);
  return EVP_OK;
}

int evp_set_control(evp_handle h, const evp_ctrl *c) {
  if (!h || !c) return EVP_ERR_ARG;
  h->ctrl = *c;
  g_active = nullptr;
  activate(h);
  return EVP_OK;
}

int evp_set_loading(evp_handle h, const int32_t iudot[9], const double udot[9], const int32_t iscau[6], const double scau[6]) {
  if (!h || !iudot || !udot || !iscau || !scau) return fail(h, EVP_ERR_ARG, "set_loading: null pointer");
  bool sc[6];
  for (int c = 0; c < 6; ++c) {
    const int i = kI[c], j = kJ[c];
    sc[c] = iudot[3 * i + j] && iudot[3 * j + i];
    if (sc[c] == (iscau[c] != 0)) return fail(h, EVP_ERR_ARG, "set_loading: each symmetric component needs exactly one of strain-rate / stress imposed");
  }
  for (int c = 0; c < 6; ++c) h->strain_ctl[c] = sc[c];
  for (int k = 0; k < 9; ++k) { h->iudot[k] = iudot[k]; h->udot[k] = udot[k]; }
  for (int k = 0; k < 6; ++k) { h->iscau[k] = iscau[k]; h->scau[k] = scau[k]; }
  h->have_loading = true;
  if (h->have_c0) {
    // dE_T = (C0_TT)^-1 W (scau - savg)_T / W, embedded into a 6x6 acting on Cartesian components
    MacroDev &m = *h->h_macro;
    std::memset(m.Mmac, 0, sizeof(m.Mmac));
    int idx[6], nt = 0;
    for (int c = 0; c < 6; ++c)
      if (!sc[c]) idx[nt++] = c;
    if (nt) {
      for (int col = 0; col < nt; ++col) {
        double A[36], r[6] = {0, 0, 0, 0, 0, 0};
        for (int a = 0; a < nt; ++a)
          for (int b = 0; b < nt; ++b) A[nt * a + b] = h->C0m[6 * idx[a] + idx[b]];
        r[col] = 1.0;
        solve_n(nt, A, r);
        for (int a = 0; a < nt; ++a) m.Mmac[6 * idx[a] + idx[col]] = r[a] * kW[idx[col]] / kW[idx[a]];
      }
    }
    for (int c = 0; c < 6; ++c) m.scau[c] = scau[c];
    CUDA_OK(h, cudaMemcpyAsync(h->d_macro->Mmac, m.Mmac, sizeof(m.Mmac), cudaMemcpyHostToDevice, h->st));
    CUDA_OK(h, cudaMemcpyAsync(h->d_macro->scau, m.scau, sizeof(m.scau), cudaMemcpyHostToDevice, h->st));
    CUDA_OK(h, cudaStreamSynchronize(h->st));
  }
  return EVP_OK;
}

int evp_begin_increment(evp_handle h, double dt) {
  if (!h) return EVP_ERR_ARG;
  if (!h->have_micro || !h->have_c0 || !h->have_loading) return fail(h, EVP_ERR_STATE, "begin_increment: microstructure, reference medium and loading must be set");
  if (!(dt > 0)) return fail(h, EVP_ERR_ARG, "dt must be positive");
  h->dt = dt;
  g_active = nullptr;
  activate(h);
  MacroDev &m = *h->h_macro;
  for (int c = 0; c < 6; ++c) {
    const int i = kI[c], j = kJ[c];
    const double rate = h->strain_ctl[c] ? 0.5 * (h->udot[3 * i + j] + h->udot[3 * j + i]) : h->Edot_prev[c];
    m.dEpend[c] = dt * rate;
    m.Et[c] = h->Et[c];
    m.E[c] = h->Et[c] + m.dEpend[c];
  }
  m.iter = 0;
  // which Newton kernel runs this increment (fast path or generic) fixes the form of the class tables and of f.itc: decided once, here
  h->k1_fast = constitutive_fast_npow(h->nphases, h->uniform_ns, h->uniform_npow, h->any_twin ? 1 : 0);
  launch_prep_increment(h->f, h->nsmax, h->k1_fast, h->st);   // orientation / CRSS invariants of this increment
  CUDA_OK(h, cudaMemcpyAsync(h->d_macro->E, m.E, sizeof(double) * 18, cudaMemcpyHostToDevice, h->st));  // E, Et, dEpend
  CUDA_OK(h, cudaMemcpyAsync(&h->d_macro->iter, &m.iter, sizeof(int), cudaMemcpyHostToDevice, h->st));
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  h->in_incr = true;
  return EVP_OK;
}

int evp_op_green(evp_handle h) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "op_green outside an increment");
  activate(h);
  h->tm.n = 0;
  int rc = enqueue_green(h);
  if (rc) return rc;
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  CUDA_OK(h, cudaGetLastError());
  return EVP_OK;
}

int evp_op_constitutive(evp_handle h, evp_iter_report *rep) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "op_constitutive outside an increment");
  activate(h);
  invalidate_green(h);   // the stress is about to change outside the pipelined iteration
  int rc = enqueue_constitutive(h);
  if (rc) return rc;
  rc = fetch_report(h, rep);
  CUDA_OK(h, cudaGetLastError());
  return rc;
}

int evp_equilibrium_iter(evp_handle h, evp_iter_report *rep) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "equilibrium_iter outside an increment");
  activate(h);
  int rc = enqueue_iteration(h);
  if (rc) return rc;
  rc = fetch_report(h, rep);
  CUDA_OK(h, cudaGetLastError());
  return rc;
}

int evp_equilibrium_iters(evp_handle h, int32_t n, evp_iter_report *last) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "equilibrium_iters outside an increment");
  activate(h);
  for (int i = 0; i < n; ++i) {
    int rc = enqueue_iteration(h);
    if (rc) return rc;
  }
  int rc = fetch_report(h, last);
  CUDA_OK(h, cudaGetLastError());
  return rc;
}

int evp_end_increment(evp_handle h, evp_step_report *rep) {
  if (!h || !h->in_incr) return fail(h, EVP_ERR_STATE, "end_increment outside an increment");
  activate(h);
  const int tex = h->ctrl.update_texture != 0, twn = (h->ctrl.update_twinning != 0 && h->f.twinf != nullptr);
  if (tex) {
    if (!h->f.de) {
      CUDA_OK(h, cudaMalloc(&h->f.de, sizeof(double) * 6 * h->N));
      CUDA_OK(h, cudaMemsetAsync(h->f.de, 0, sizeof(double) * 6 * h->N, h->st));
    }
    int rc = enqueue_rotation_field(h);
    if (rc) return rc;
  }
  const double wapp[3] = {0.5 * (h->udot[7] - h->udot[5]), 0.5 * (h->udot[2] - h->udot[6]), 0.5 * (h->udot[3] - h->udot[1])};
  launch_commit(h->f, h->nsmax, h->dt, wapp, tex, twn, h->d_partials, h->st);
  launch_reduce(h->d_partials, h->N, h->d_scratch, h->d_totals + 16, h->st);
  if (h->nranks > 1) {
    int rc = g_nccl.AllReduce(h->d_totals + 16, h->d_totals + 16, 11, kNcclDouble, kNcclSum, h->comm2 ? h->comm2 : h->comm, h->st);
    if (rc) return nccl_check(h, rc, "nccl allreduce (commit)");
  }
  double tot[12];
  CUDA_OK(h, cudaMemcpyAsync(tot, h->d_totals + 16, sizeof(tot), cudaMemcpyDeviceToHost, h->st));
  CUDA_OK(h, cudaMemcpyAsync(h->h_macro, h->d_macro, sizeof(MacroDev), cudaMemcpyDeviceToHost, h->st));
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  CUDA_OK(h, cudaGetLastError());
  // F_acc is a history sum (Tome, Lebensohn, Kocks 1991): this increment's twin fraction is added, nothing is ever
  // subtracted when a voxel reorients and its per-system fractions are reset
  if (twn) h->facc += tot[1] / h->Ntot;
  const double Facc = h->facc;
  long long nre = 0;
  if (twn) {
    const double Feff = (double)h->ntwinned / h->Ntot;
    launch_twin_reorient(h->f, (Facc > 0.0) ? Feff / Facc : 0.0, h->d_partials, h->st);
    launch_reduce(h->d_partials, h->N, h->d_scratch, h->d_totals + 32, h->st);
    if (h->nranks > 1) {
      int rc = g_nccl.AllReduce(h->d_totals + 32, h->d_totals + 32, 1, kNcclDouble, kNcclSum, h->comm2 ? h->comm2 : h->comm, h->st);
      if (rc) return nccl_check(h, rc, "nccl allreduce (PTR)");
    }
    double cnt = 0;
    CUDA_OK(h, cudaMemcpyAsync(&cnt, h->d_totals + 32, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CUDA_OK(h, cudaStreamSynchronize(h->st));
    nre = (long long)(cnt + 0.5);
    h->ntwinned += nre;
  }
  if (tex || nre > 0) {   // orientations are now per voxel: the invariant tables go per voxel too (rebuilt at begin_increment)
    int rc = switch_to_voxel_classes(h);
    if (rc) return rc;
    CUDA_OK(h, cudaStreamSynchronize(h->st));
  }
  MacroDev &m = *h->h_macro;
  // the pending macro correction belongs to an iteration that will not run
  for (int c = 0; c < 6; ++c) {
    m.E[c] -= m.dEpend[c];
    m.dEpend[c] = 0.0;
    h->Edot_prev[c] = (m.E[c] - h->Et[c]) / h->dt;
    h->Et[c] = m.E[c];
    m.Et[c] = m.E[c];
  }
  CUDA_OK(h, cudaMemcpyAsync(h->d_macro->E, m.E, sizeof(double) * 18, cudaMemcpyHostToDevice, h->st));
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  h->in_incr = false;
  if (rep) {
    rep->iters = m.iter;
    rep->err_stress = m.err_s; rep->err_strain = m.err_e;
    rep->converged = (m.err_s <= h->ctrl.tol_stress && m.err_e <= h->ctrl.tol_strain && m.unconverged == 0) ? 1 : 0;
    for (int c = 0; c < 6; ++c) { rep->savg[c] = m.savg[c]; rep->emacro[c] = m.E[c]; rep->epavg[c] = tot[2 + c] / h->Ntot; }
    rep->twin_acc = Facc; rep->twin_eff = (double)h->ntwinned / h->Ntot; rep->reoriented = nre;
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

int evp_get_field(evp_handle h, evp_field f, void *host, size_t bytes) {
  if (!h || !host) return EVP_ERR_ARG;
  cudaSetDevice(h->device);
  size_t el;
  void *p = field_ptr(h, f, &el);
  const size_t need = field_comps(h, f) * (size_t)h->N * el;
  if (field_comps(h, f) == 0) return fail(h, EVP_ERR_ARG, "unknown field");
  if (bytes != need) return fail(h, EVP_ERR_ARG, "get_field: size mismatch");
  if (!p) {  // optional field that was never allocated (twin fractions without twins, strain increment)
    std::memset(host, 0, need);
    return EVP_OK;
  }
  return copy_d2h(h, host, p, need);
}

int evp_set_field(evp_handle h, evp_field f, const void *host, size_t bytes) {
  if (!h || !host) return EVP_ERR_ARG;
  cudaSetDevice(h->device);
  size_t el;
  void *p = field_ptr(h, f, &el);
  const size_t need = field_comps(h, f) * (size_t)h->N * el;
  if (field_comps(h, f) == 0) return fail(h, EVP_ERR_ARG, "unknown field");
  if (bytes != need) return fail(h, EVP_ERR_ARG, "set_field: size mismatch");
  if (!p) return fail(h, EVP_ERR_STATE, "field is not allocated in this configuration");
  { int rc = copy_h2d(h, p, host, need); if (rc) return rc; }
  if (f == EVP_FIELD_ROTATION) {
    int rc = switch_to_voxel_classes(h);
    if (rc) return rc;
  }
  if (f == EVP_FIELD_STRESS) invalidate_green(h);
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  return EVP_OK;
}

int evp_get_macro(evp_handle h, double emacro[6], double savg[6]) {
  if (!h) return EVP_ERR_ARG;
  cudaSetDevice(h->device);
  CUDA_OK(h, cudaMemcpyAsync(h->h_macro, h->d_macro, sizeof(MacroDev), cudaMemcpyDeviceToHost, h->st));
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  for (int c = 0; c < 6; ++c) {
    if (emacro) emacro[c] = h->h_macro->E[c];
    if (savg) savg[c] = h->h_macro->savg[c];
  }
  return EVP_OK;
}


// Restart file format "EVPCKPT2" (shared with oracle/evp_oracle.cpp; documented in include/evpfft.h)
struct CkptHeader {
  char magic[8];
  int32_t nx, ny, nz, y0, nyl, z0, nzl, nsmax, nranks, rank;
  double Et[6], Edot_prev[6];
  int64_t ntwinned;
  double facc;
};

static int ckpt_io(evp_handle h, std::FILE *f, bool write) {
  const size_t N = (size_t)h->N, ns = (size_t)std::max(h->nsmax, 1);
  struct Item { void *p; size_t bytes; };
  const Item items[] = {{h->f.sig, 6 * N * 8}, {h->f.e, 6 * N * 8}, {h->f.epsp, 6 * N * 8}, {h->f.crss, ns * N * 8}, {h->f.rot, 9 * N * 8},
                        {h->f.gacc, N * 8}, {h->f.twinf, ns * N * 8}, {h->f.wrot, 3 * N * 8}, {h->f.grain, N * 4}, {h->f.phase, N * 4},
                        {h->f.twinned, N * 4}};
  const size_t kBuf = (size_t)64 << 20;
  std::vector<char> buf(kBuf);
  for (const Item &it : items) {
    for (size_t off = 0; off < it.bytes; off += kBuf) {
      const size_t nb = std::min(kBuf, it.bytes - off);
      if (write) {
        if (it.p) { CUDA_OK(h, cudaMemcpy(buf.data(), (char *)it.p + off, nb, cudaMemcpyDeviceToHost)); }
        else std::memset(buf.data(), 0, nb);   // optional field not allocated (twin fractions without twin modes)
        if (std::fwrite(buf.data(), 1, nb, f) != nb) return fail(h, EVP_ERR_ARG, "short write");
      } else {
        if (std::fread(buf.data(), 1, nb, f) != nb) return fail(h, EVP_ERR_ARG, "short read");
        if (it.p) CUDA_OK(h, cudaMemcpy((char *)it.p + off, buf.data(), nb, cudaMemcpyHostToDevice));
      }
    }
  }
  return EVP_OK;
}

int evp_save_state(evp_handle h, const char *path) {
  if (!h || !path) return EVP_ERR_ARG;
  if (h->in_incr) return fail(h, EVP_ERR_STATE, "save_state inside an increment");
  cudaSetDevice(h->device);
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  std::FILE *f = std::fopen(path, "wb");
  if (!f) return fail(h, EVP_ERR_ARG, std::string("cannot write ") + path);
  CkptHeader hd{};
  std::memcpy(hd.magic, "EVPCKPT2", 8);
  hd.nx = h->nx; hd.ny = h->ny; hd.nz = h->nz; hd.y0 = h->y0; hd.nyl = h->nyb; hd.z0 = h->z0; hd.nzl = h->nzl; hd.nsmax = h->nsmax;
  hd.nranks = h->nranks; hd.rank = h->rank;
  for (int c = 0; c < 6; ++c) { hd.Et[c] = h->Et[c]; hd.Edot_prev[c] = h->Edot_prev[c]; }
  hd.ntwinned = h->ntwinned;
  hd.facc = h->facc;
  int rc = (std::fwrite(&hd, sizeof(hd), 1, f) == 1) ? EVP_OK : fail(h, EVP_ERR_ARG, "short write");
  if (rc == EVP_OK) rc = ckpt_io(h, f, true);
  std::fclose(f);
  return rc;
}

int evp_load_state(evp_handle h, const char *path) {
  if (!h || !path) return EVP_ERR_ARG;
  if (h->in_incr) return fail(h, EVP_ERR_STATE, "load_state inside an increment");
  activate(h);
  invalidate_green(h);
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  std::FILE *f = std::fopen(path, "rb");
  if (!f) return fail(h, EVP_ERR_ARG, std::string("cannot read ") + path);
  CkptHeader hd{};
  bool ok = std::fread(&hd, sizeof(hd), 1, f) == 1 && std::memcmp(hd.magic, "EVPCKPT2", 8) == 0;
  if (ok && (hd.nx != h->nx || hd.ny != h->ny || hd.nz != h->nz || hd.nzl != h->nzl || hd.z0 != h->z0 || hd.nyl != h->nyb || hd.y0 != h->y0 ||
             hd.nsmax != h->nsmax)) ok = false;
  if (!ok) { std::fclose(f); return fail(h, EVP_ERR_ARG, "checkpoint does not match this handle"); }
  {
    // the whole payload must be there before any device field is overwritten (a truncated file leaves the state untouched)
    const size_t N = (size_t)h->N, ns = (size_t)std::max(h->nsmax, 1);
    const size_t need = sizeof(hd) + (size_t)8 * N * (6 + 6 + 6 + ns + 9 + 1 + ns + 3) + (size_t)4 * N * 3;
    std::fseek(f, 0, SEEK_END);
    const long have = std::ftell(f);
    std::fseek(f, (long)sizeof(hd), SEEK_SET);
    if (have < 0 || (size_t)have != need) { std::fclose(f); return fail(h, EVP_ERR_ARG, "checkpoint file is truncated or has trailing bytes"); }
  }
  int rc = ckpt_io(h, f, false);
  std::fclose(f);
  if (rc) return rc;
  MacroDev &m = *h->h_macro;
  for (int c = 0; c < 6; ++c) {
    h->Et[c] = hd.Et[c]; h->Edot_prev[c] = hd.Edot_prev[c];
    m.E[c] = m.Et[c] = hd.Et[c]; m.dEpend[c] = 0.0;
  }
  h->ntwinned = hd.ntwinned;
  h->facc = hd.facc;
  CUDA_OK(h, cudaMemcpyAsync(h->d_macro->E, m.E, sizeof(double) * 18, cudaMemcpyHostToDevice, h->st));
  rc = switch_to_voxel_classes(h);   // orientations may be anything now
  if (rc) return rc;
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  h->have_micro = true;
  return EVP_OK;
}

int evp_debug_spectrum(evp_handle h, int32_t comp, double *out) {
  if (!h || !out || comp < 0 || comp > 5) return EVP_ERR_ARG;
  if (h->nranks != 1) return fail(h, EVP_ERR_UNSUPPORTED, "debug_spectrum: single-rank handles only");
  activate(h);
  invalidate_green(h);
  launch_xfwd(h->nx, h->f.sig, h->WA, h->N, 0, h->ny * h->nzl, h->Lplain, h->twx, h->st);
  launch_ypass(h->ny, false, h->ch[0].in_fwd, false, h->ch[0].out_inv, false, h->ti_y_plain, h->ti_y_plain, h->nxh, h->nzl, h->twy, h->st);
  launch_zfused(h->nz, 1, true, h->zmaps, h->zout, false, h->lg_nzl, h->lg_nzc, h->zrun, h->nxh, 0, h->ny, 0, h->nx, h->ny, h->g.dx, h->g.dy, h->g.dz, h->twz, h->st);
  CUDA_OK(h, cudaMemcpy2DAsync(out, sizeof(double2) * h->nxh, h->WA + (size_t)comp * h->Lplain.cstride, sizeof(double2) * h->nxp,
                               sizeof(double2) * h->nxh, (size_t)h->nz * h->ny, cudaMemcpyDeviceToHost, h->st));
  CUDA_OK(h, cudaStreamSynchronize(h->st));
  CUDA_OK(h, cudaGetLastError());
  return EVP_OK;
}

int evp_set_profiling(evp_handle h, int32_t on) {
  if (!h) return EVP_ERR_ARG;
  cudaSetDevice(h->device);
  h->flags = on;
  if ((on & 1) && !h->tm.made) {
    for (int i = 0; i < 128; ++i) { CUDA_OK(h, cudaEventCreate(&h->tm.a[i])); CUDA_OK(h, cudaEventCreate(&h->tm.b[i])); }
    h->tm.made = true;
  }
  if (on & 2) invalidate_green(h);
  if ((on & 2) && !h->f.de) {
    CUDA_OK(h, cudaMalloc(&h->f.de, sizeof(double) * 6 * h->N));
    CUDA_OK(h, cudaMemset(h->f.de, 0, sizeof(double) * 6 * h->N));
  }
  return EVP_OK;
}

int evp_last_kernel_ms(evp_handle h, double ms[8]) {
  if (!h || !ms) return EVP_ERR_ARG;
  for (int i = 0; i < 8; ++i) ms[i] = h->last_ms[i];
  return EVP_OK;
}

}  // extern "C"
