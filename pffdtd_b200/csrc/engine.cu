// engine.cu -- libpffdtd_b200.so: the C ABI of include/pffdtd_b200.h over the sm_100a kernels.
//
// Replaces run_sim of the reference (c_cuda/gpu_engine.h:665-1255): allocation + upload
// (:739-974), the host-driven time loop (:993-1170) and the halo exchange (:1086-1126).  Step order
// and arithmetic follow the reference CPU engine (cpu_engine.h:129-325), see kernels.cuh.
//
// One engine = one slab of the grid on one device; one host thread per engine.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <dlfcn.h>
#include <cuda_runtime.h>

#include "pffdtd_b200.h"
#include "kernels.cuh"
#include "air_tma.cuh"
#include "energy.cuh"
#include "vox.cuh"

using pf::i64;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
   char buf[1024];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   g_err = buf;
   return code;
}
#define CU(call)                                                                                           \
   do {                                                                                                    \
      cudaError_t err__ = (call);                                                                          \
      if (err__ != cudaSuccess)                                                                            \
         return fail(PFFDTD_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
   } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time so that the library loads on hosts without it
// ------------------------------------------------------------------------------------------------
struct Nccl {
   void *lib = nullptr;
   int (*GetUniqueId)(void *) = nullptr;
   int (*CommInitRank)(void **, int, /*ncclUniqueId by value*/ struct Id128, int) = nullptr;
   int (*CommDestroy)(void *) = nullptr;
   int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
   int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
};
struct Id128 {
   char b[128];
};
static Nccl g_nccl;
static int nccl_load() {
   if (g_nccl.lib) return 0;
   // Prefer an NCCL that is already in the process (the host's torch brings its own and must not get a second,
   // older one forced on it); never export its symbols (RTLD_LOCAL).
   const char *names[] = {getenv("PFFDTD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
   void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
   for (const char *n : names) {
      if (h) break;
      if (n && *n) h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
   }
   if (!h) return fail(PFFDTD_ENCCL, "cannot load libnccl.so.2 (set PFFDTD_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                                      \
   *(void **)(&g_nccl.field) = dlsym(h, name);                                                \
   if (!g_nccl.field) return fail(PFFDTD_ENCCL, "libnccl lacks %s", name);
   SYM(GetUniqueId, "ncclGetUniqueId")
   SYM(CommInitRank, "ncclCommInitRank")
   SYM(CommDestroy, "ncclCommDestroy")
   SYM(Send, "ncclSend")
   SYM(Recv, "ncclRecv")
   SYM(GroupStart, "ncclGroupStart")
   SYM(GroupEnd, "ncclGroupEnd")
   SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
   g_nccl.lib = h;
   return 0;
}
#define NC(call)                                                                                       \
   do {                                                                                                \
      int err__ = (call);                                                                              \
      if (err__ != 0)                                                                                  \
         return fail(PFFDTD_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(err__)); \
   } while (0)

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
struct EvPair {
   cudaEvent_t a, b;
};

struct pffdtd_engine {
   int device = 0, precision = 0, fcc = 0, Nm = 0, NN = 6;
   i64 Nx = 0, Ny = 0, Nz = 0, Nzp = 0, mwpr = 0, Nb = 0, Nbl = 0, Nba = 0, Ns = 0, Nr = 0, Nt = 0;
   i64 Nblp = 0;  // pitch of the branch-major boundary state: Nbl rounded up to 32 elements
   i64 ix0 = 0;
   int x_lo_edge = 1, x_hi_edge = 1;
   double l = 0, l2 = 0, a1 = 0, a2 = 0, sl2 = 0, lo2 = 0;
   size_t rs = 4;  // sizeof(Real)
   pf::Offsets off{};
   // device memory
   void *u[2] = {nullptr, nullptr};  // u[cur] = u1 (state n), u[cur^1] = u0
   int cur = 0;
   uint32_t *mask = nullptr;
   i64 *bn = nullptr, *bnl = nullptr, *bna = nullptr, *in = nullptr, *out = nullptr;
   uint16_t *adj = nullptr;
   int8_t *mat = nullptr, *Q = nullptr, *Mb = nullptr;
   void *ssaf = nullptr, *beta = nullptr, *quads = nullptr, *insig = nullptr, *uout = nullptr;
   void *hist[2] = {nullptr, nullptr}, *u2ba = nullptr, *vh1 = nullptr, *gh1 = nullptr;
   void *lo2Kbg = nullptr, *facb = nullptr;  // per lossy node constants (k_fd_prep)
   void *zold = nullptr, *yold = nullptr, *xold = nullptr;  // pre-update values of the shell nodes (fused step)
   uint16_t *matmb = nullptr;                // per lossy node: material | Mb << 8
   int serial_src = 0;
   int ticked = 0;       // this step's last k_io advanced the device step counter
   i64 *d_n = nullptr;   // device step counter read by k_io / k_fd
   i64 n_dev = -1;       // value the host knows it holds (-1 unknown)
   cudaGraphExec_t graph[2] = {nullptr, nullptr};  // two consecutive steps starting with cur = 0 / 1
   cudaGraphExec_t graphA[2] = {nullptr, nullptr}, graphB[2] = {nullptr, nullptr};  // NCCL slabs: one step's work before / after the exchange
   double graphA_launches[2] = {0, 0}, graphB_launches[2] = {0, 0};
   cudaGraphExec_t hgraph[2] = {nullptr, nullptr};  // one host-driven step (H2D samples, step, D2H samples) with cur = 0 / 1
   double hgraph_launches[2] = {0, 0};
   void *in_stage = nullptr, *out_stage = nullptr;  // device staging of one step's source / receiver samples
   int host_mode = 0;                               // the step being launched belongs to pffdtd_step_host
   double graph_launches[2] = {0, 0};
   int use_graph = 1;
   i64 steps_plain = 0;  // steps launched kernel by kernel so far
   // fused Cartesian step (tiled air kernel applies the ABC shell and mirrors the halos on write)
   int fuse_ok = 0;     // the ABC list is the canonical full shell, so the air kernel may apply it
   int halo_dirty = 0;  // u1's halos were not produced by mirror-on-write: run the mirror kernels first
   // unfused step: k_abc writes the z halos of the new state (kernels.cuh k_abc `zf`), the next mirror pass skips k_flip_z
   uint8_t *zf = nullptr;
   int zflip_want = 1, zflip_ok = 0;  // option; the lists allow it (canonical shell, no boundary / source node at z = 2 | Nz-3)
   int zhalo_ok = 0;                  // the current state's z halos were written by the previous step's k_abc
   i64 *pair_src = nullptr, *pair_dst = nullptr;  // late halo mirrors (boundary / source nodes at index 2 | N-3)
   i64 np = 0, np_lo = 0, np_hi = 0;
   // prefix/suffix sizes of the sorted node lists that lie in the first/last owned plane (edge work
   // that must finish before the halo exchange); only valid when `sorted`
   int sorted = 0;
   i64 nb_lo = 0, nb_hi = 0, nbl_lo = 0, nbl_hi = 0, nba_lo = 0, nba_hi = 0, ns_lo = 0, ns_hi = 0;
   // host staging (pinned)
   void *h_in = nullptr, *h_out = nullptr;
   // streams / events
   cudaStream_t s_main = nullptr, s_comm = nullptr;
   cudaEvent_t ev_edge = nullptr, ev_comm = nullptr, ev_step = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
   cudaEvent_t ev_push = nullptr;  // (single-process slabs) this slab's edge planes have landed in the neighbours' halos
   i64 first_step = 0;             // first step of the current pffdtd_multi_run_steps batch (no neighbour event to wait for before it)
   // comm
   void *comm = nullptr;
   int rank = 0, nranks = 1, comm_pending = 0;
   // multi-process slabs over peer memory (pffdtd_peer_export / _connect): IPC mappings of the neighbours' grids and flag words;
   // the halo planes are copied straight into them and a flag tells the neighbour (no NCCL in the step)
   int p2p = 0;
   long long *flags = nullptr;                 // [0] / [1]: steps whose plane the lower / upper neighbour has delivered, [2] this slab's step count, [3] error
   void *ipc_u_lo[2] = {nullptr, nullptr}, *ipc_u_hi[2] = {nullptr, nullptr};  // the neighbours' grids
   long long *ipc_flags_lo = nullptr, *ipc_flags_hi = nullptr;
   i64 nx_lo = 0;                              // planes of the lower neighbour's slab (its upper halo plane is nx_lo - 1)
   // single-process multi-GPU (pffdtd_multi_*): the engines of the neighbouring slabs; halo planes are pushed into their grids
   // with peer copies instead of NCCL send/recv
   pffdtd_engine *peer_lo = nullptr, *peer_hi = nullptr;
   // options
   int air_kernel = 1, overlap = 1, profile_air = 0, manual_halo = 0, fuse = -1;
   int abc_overlap = 1;  // the absorbing-shell kernel runs beside the boundary kernels when their node sets are disjoint
   int abc_disjoint = 0, abc_pending = 0;
   cudaStream_t s_abc = nullptr;
   cudaEvent_t ev_abc0 = nullptr, ev_abc1 = nullptr;
   int edge_overlap = 0;  // 1: slabs with two neighbours run the upper edge plane on s_edge beside the lower one (measured at 4 GPUs: no gain, two persistent air kernels do not share the SMs)
   cudaStream_t s_edge = nullptr;
   cudaEvent_t ev_e0 = nullptr, ev_e1 = nullptr;
   // in-kernel boundary work (air_tma.cuh AirSvc): per tile-plane lists of the sparse rigid nodes + the shell's z faces, and the
   // dense remainder of the boundary list that stays with k_rigid
   int mb_max = 0;  // largest branch count among the materials
   int fd_bulk = 0; // 1: k_fd_bulk (branch state through the TMA unit) instead of k_fd -- measured 53.6 vs 50.9 us on c2: not the default
   int svc_want = -1, svc_on = 0, svc_cap = 64;  // svc_want / fuse: -1 = the layout's default (7-point: on; 13-point: off, measured), 0, 1
   int svc_shell = 0;   // the lists hold the shell's z faces too (fused Cartesian step); otherwise rigid nodes only (any step)
   int bn_off_abc = 0;  // no boundary node is also an absorbing-shell node: the rigid update commutes with the shell update
   float negzero = -0.0f;  // travels as a kernel argument so that the compiler cannot fold it (air_tma.cuh "packed fp32 arithmetic")
   uint32_t *svc_list = nullptr, *svc_off = nullptr;
   i64 *bn_left = nullptr;
   uint16_t *adj_left = nullptr;
   i64 Nb_left = 0, nbL_lo = 0, nbL_hi = 0, svc_entries = 0;
   // stats
   i64 steps_done = 0;  // next time index expected by run_steps
   double launches = 0;
   std::vector<EvPair> air_ev;
   size_t air_ev_used = 0;
   double air_ms = 0;
   i64 air_timed = 0;
   pf::AirTma tma;
   // energy balance (energy.cuh): third grid Lu, copies of the pre-step branch / source-node values, partial sums
   int energy_on = 0;
   void *Lu = nullptr, *vold = nullptr, *u2in = nullptr;
   double *en_def = nullptr, *en_part = nullptr, *en_H = nullptr, *en_lost = nullptr, *en_in = nullptr;
   pf::EnergyCoef en_k{};
   double en_Ts = 0;
   std::vector<void *> allocs;
};

template <typename T>
static int dalloc(pffdtd_engine *e, T **p, size_t count, bool zero = true) {
   size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
   CU(cudaMalloc((void **)p, bytes));
   e->allocs.push_back((void *)*p);
   if (zero) CU(cudaMemset(*p, 0, bytes));
   return 0;
}
static int dalloc_bytes(pffdtd_engine *e, void **p, size_t bytes, bool zero = true) {
   bytes = std::max<size_t>(bytes, 16);
   CU(cudaMalloc(p, bytes));
   e->allocs.push_back(*p);
   if (zero) CU(cudaMemset(*p, 0, bytes));
   return 0;
}

static void dfree(pffdtd_engine *e, void *p) {
   if (!p) return;
   auto it = std::find(e->allocs.begin(), e->allocs.end(), p);
   if (it != e->allocs.end()) e->allocs.erase(it);
   cudaFree(p);
}

// doubles holding Real-rounded values -> Real array on the device
static int upload_real(pffdtd_engine *e, void **dst, const double *src, size_t n) {
   if (dalloc_bytes(e, dst, n * e->rs)) return PFFDTD_ECUDA;
   if (n == 0) return 0;
   if (e->precision == 2) {
      CU(cudaMemcpy(*dst, src, n * 8, cudaMemcpyHostToDevice));
   } else {
      std::vector<float> tmp(n);
      for (size_t i = 0; i < n; i++) tmp[i] = (float)src[i];
      CU(cudaMemcpy(*dst, tmp.data(), n * 4, cudaMemcpyHostToDevice));
   }
   return 0;
}

// reference-layout linear indices -> padded pitch; checks range
static int upload_idx(pffdtd_engine *e, i64 **dst, const int64_t *src, i64 n, const char *what, bool interior) {
   if (dalloc(e, dst, (size_t)n, false)) return PFFDTD_ECUDA;
   if (n == 0) return 0;
   if (!src) return fail(PFFDTD_EINVAL, "%s is NULL", what);
   std::vector<i64> tmp((size_t)n);
   const i64 Npts = e->Nx * e->Ny * e->Nz;
   for (i64 i = 0; i < n; i++) {
      const i64 v = src[i];
      if (v < 0 || v >= Npts) return fail(PFFDTD_EINVAL, "%s[%lld]=%lld outside the grid", what, (long long)i, (long long)v);
      const i64 row = v / e->Nz, iz = v - row * e->Nz;
      if (interior) {
         const i64 ix = row / e->Ny, iy = row - ix * e->Ny;
         if (ix < 1 || ix > e->Nx - 2 || iy < 1 || iy > e->Ny - 2 || iz < 1 || iz > e->Nz - 2)
            return fail(PFFDTD_EINVAL, "%s[%lld]=%lld lies on the halo layer", what, (long long)i, (long long)v);
      }
      tmp[(size_t)i] = row * e->Nzp + iz;
   }
   CU(cudaMemcpy(*dst, tmp.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
   return 0;
}

template <typename T>
static int upload_raw(pffdtd_engine *e, T **dst, const T *src, i64 n, const char *what) {
   if (dalloc(e, dst, (size_t)n, false)) return PFFDTD_ECUDA;
   if (n == 0) return 0;
   if (!src) return fail(PFFDTD_EINVAL, "%s is NULL", what);
   CU(cudaMemcpy(*dst, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice));
   return 0;
}

// count of leading entries of a sorted list below `limit`, and of trailing entries >= `from`
static void edge_counts(const int64_t *a, i64 n, i64 limit, i64 from, i64 *lo, i64 *hi) {
   *lo = std::lower_bound(a, a + n, limit) - a;
   *hi = (a + n) - std::lower_bound(a, a + n, from);
}
static bool ascending(const int64_t *a, i64 n, bool strict = true) {
   for (i64 i = 1; i < n; i++)
      if (strict ? a[i] <= a[i - 1] : a[i] < a[i - 1]) return false;
   return true;
}

extern "C" const char *pffdtd_last_error(void) { return g_err.c_str(); }
extern "C" const char *pffdtd_version(void) { return "pffdtd_b200 0.1 sm_100a"; }

extern "C" int pffdtd_destroy(pffdtd_engine *e) {
   if (!e) return PFFDTD_OK;
   cudaSetDevice(e->device);
   if (e->s_main) cudaStreamSynchronize(e->s_main);
   if (e->s_comm) cudaStreamSynchronize(e->s_comm);
   if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
   for (int k = 0; k < 2; k++) {
      if (e->ipc_u_lo[k]) cudaIpcCloseMemHandle(e->ipc_u_lo[k]);
      if (e->ipc_u_hi[k]) cudaIpcCloseMemHandle(e->ipc_u_hi[k]);
   }
   if (e->ipc_flags_lo) cudaIpcCloseMemHandle(e->ipc_flags_lo);
   if (e->ipc_flags_hi) cudaIpcCloseMemHandle(e->ipc_flags_hi);
   for (void *p : e->allocs) cudaFree(p);
   if (e->h_in) cudaFreeHost(e->h_in);
   if (e->h_out) cudaFreeHost(e->h_out);
   for (auto &p : e->air_ev) {
      cudaEventDestroy(p.a);
      cudaEventDestroy(p.b);
   }
   for (int c = 0; c < 2; c++) {
      if (e->graph[c]) cudaGraphExecDestroy(e->graph[c]);
      if (e->hgraph[c]) cudaGraphExecDestroy(e->hgraph[c]);
      if (e->graphA[c]) cudaGraphExecDestroy(e->graphA[c]);
      if (e->graphB[c]) cudaGraphExecDestroy(e->graphB[c]);
   }
   if (e->ev_abc0) cudaEventDestroy(e->ev_abc0);
   if (e->ev_abc1) cudaEventDestroy(e->ev_abc1);
   if (e->s_abc) cudaStreamDestroy(e->s_abc);
   if (e->ev_e0) cudaEventDestroy(e->ev_e0);
   if (e->ev_e1) cudaEventDestroy(e->ev_e1);
   if (e->s_edge) cudaStreamDestroy(e->s_edge);
   if (e->ev_edge) cudaEventDestroy(e->ev_edge);
   if (e->ev_comm) cudaEventDestroy(e->ev_comm);
   if (e->ev_step) cudaEventDestroy(e->ev_step);
   if (e->ev_push) cudaEventDestroy(e->ev_push);
   if (e->ev_t0) cudaEventDestroy(e->ev_t0);
   if (e->ev_t1) cudaEventDestroy(e->ev_t1);
   if (e->s_main) cudaStreamDestroy(e->s_main);
   if (e->s_comm) cudaStreamDestroy(e->s_comm);
   delete e;
   return PFFDTD_OK;
}

// Lists for the air kernel's service warp (air_tma.cuh AirSvc), for the tile shape of the current configuration.
// Per (tile, plane): the z faces of the absorbing shell that lie in it (rows 2..Ny-3 of planes off the x shell: Q = 1) and, when
// the tile-plane holds at most `svc_cap` boundary nodes, those nodes with their adjacency.  Boundary nodes of denser tile-planes
// (walls perpendicular to x or y: contiguous runs along z) go to the "left" list for k_rigid, where they coalesce.
// Only for the fused Cartesian step with the canonical shell and no boundary / source node on it (abc_disjoint): then the shell
// update, the rigid update and the air update touch disjoint nodes and commute.
// Rigid nodes may go to the service warp whenever no boundary node is a shell node (bn_off_abc); the shell's z faces go with them
// when the step is the fused Cartesian one (then the air kernel stashes nothing for them and k_abc_faces skips them).
// Defaults by layout.  7-point: fused step + service warp (c2: 0.251 -> 0.211 ms per step).  13-point: neither -- its air kernel is
// bound by its own consumers' issue / LSU slots, not by HBM, so every instruction moved into it costs what it saves elsewhere:
// measured on B200 (c3s), unfused 0.449 ms per step with or without the service warp, fused 0.560 ms (the fused epilogue spills at the
// 72 registers that keep 11 consumer warps).  Both stay available (options "fuse" / "svc" = 1) and are covered by the parity tests.
static bool zflip_on(const pffdtd_engine *e) { return e->zflip_want && e->zflip_ok && e->zf && !e->energy_on && !e->manual_halo; }
static bool want_fuse(const pffdtd_engine *e) { return e->fuse < 0 ? e->fcc == 0 : e->fuse != 0; }
static bool want_svc(const pffdtd_engine *e) { return e->svc_want < 0 ? e->fcc == 0 : e->svc_want != 0; }
static bool step_fused(const pffdtd_engine *e) {
   // (the 13-point kernel has no stash for the shell's z faces: its fused step needs the service warp to do them)
   const bool fcc_ok = e->fcc == 0 || (want_svc(e) && e->tma.svc && e->bn_off_abc && e->abc_disjoint);
   // (a grid whose shell node z = Nz-2 opens a z tile: the 7-point kernel handles it -- see `lone` / `tail2` in air_tma.cuh --, the
   // fused 13-point epilogue does not)
   return want_fuse(e) && e->fuse_ok && fcc_ok && e->air_kernel == 1 && e->tma.ok && !(e->tma.z_edge && e->fcc != 0) && !e->energy_on;
}
static bool svc_eligible(const pffdtd_engine *e) { return want_svc(e) && e->bn_off_abc && e->tma.ok && e->tma.svc && !e->energy_on; }
static int build_service(pffdtd_engine *e) {
   dfree(e, e->svc_list), dfree(e, e->svc_off), dfree(e, e->bn_left), dfree(e, e->adj_left);
   e->svc_list = e->svc_off = nullptr, e->bn_left = nullptr, e->adj_left = nullptr;
   e->svc_on = 0, e->Nb_left = e->nbL_lo = e->nbL_hi = e->svc_entries = 0;
   e->tma.sv = pf::AirSvc{nullptr, nullptr, 0};
   if (!svc_eligible(e)) return 0;
   const bool with_shell = step_fused(e) && e->abc_disjoint;
   const bool fold = e->fcc == 2, checker = e->fcc == 1;
   e->svc_shell = with_shell;
   const i64 Nx = e->Nx, Ny = e->Ny, Nz = e->Nz, Nzp = e->Nzp, TY = e->tma.ty, TZ = e->tma.tzn;
   const i64 tzc = (Nz - 1 + TZ - 1) / TZ, tyc = (Ny - 2 + TY - 1) / TY, ntile = tzc * tyc, pitch = Nx + 1;
   if (TY > 64 || TZ > 128 || Ny < 6) return 0;
   std::vector<i64> bn((size_t)e->Nb);
   std::vector<uint16_t> adj((size_t)e->Nb);
   if (e->Nb) {
      CU(cudaMemcpy(bn.data(), e->bn, (size_t)e->Nb * 8, cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(adj.data(), e->adj, (size_t)e->Nb * 2, cudaMemcpyDeviceToHost));
   }
   auto key_of = [&](i64 c, i64 *r, i64 *col) {
      const i64 row = c / Nzp, iz = c - row * Nzp, ix = row / Ny, iy = row - ix * Ny;
      const i64 ty = (iy - 1) / TY, tz = iz / TZ;
      *r = iy - 1 - ty * TY, *col = iz - tz * TZ;
      return (ty * tzc + tz) * pitch + ix;
   };
   std::vector<uint32_t> cnt((size_t)(ntile * pitch), 0u);
   for (i64 i = 0; i < e->Nb; i++) {
      i64 r, c;
      cnt[(size_t)key_of(bn[(size_t)i], &r, &c)]++;
   }
   // shell rows of a tile row: y in [2, Ny-3]
   auto shell_rows = [&](i64 ty, i64 *ya, i64 *yb) {
      *ya = std::max<i64>(2, 1 + ty * TY), *yb = std::min<i64>(fold ? Ny - 2 : Ny - 3, ty * TY + TY);
   };
   auto on_xshell = [&](i64 ix) { return (e->x_lo_edge && ix == 1) || (e->x_hi_edge && ix == Nx - 2); };
   const i64 tz_hi = (Nz - 2) / TZ;
   auto shell_node = [&](i64 x, i64 y, i64 z) { return !checker || ((e->ix0 + x + y + z) & 1) == 0; };  // (stored nodes of the layout)
   // a tile-plane's entries travel into its shared-memory stage (PF_SVC_CAP of them at most, segments padded to 16 bytes)
   const uint32_t cap = (uint32_t)std::max<i64>(0, std::min<i64>(e->svc_cap, PF_SVC_CAP - (with_shell ? 2 * TY : 0)));
   std::vector<uint32_t> off((size_t)(ntile * pitch), 0u);
   uint64_t total = 0;
   for (i64 t = 0; t < ntile; t++) {
      const i64 ty = t / tzc, tz = t - ty * tzc;
      i64 ya, yb;
      shell_rows(ty, &ya, &yb);
      for (i64 x = 0; x <= Nx; x++) {
         off[(size_t)(t * pitch + x)] = (uint32_t)total;
         if (x >= 1 && x <= Nx - 2) {
            uint64_t n = 0;
            if (with_shell && !on_xshell(x))
               for (i64 y = ya; y <= yb; y++)
                  n += ((tz == 0 && shell_node(x, y, 1)) ? 1 : 0) + ((tz == tz_hi && shell_node(x, y, Nz - 2)) ? 1 : 0);
            const uint32_t k = cnt[(size_t)(t * pitch + x)];
            if (k <= cap) n += k;
            total += (n + 3) / 4 * 4;
         }
      }
      if (total > 0xfffffff0ull) return 0;  // (never on one device's slab; keep the list kernels)
   }
   std::vector<uint32_t> list((size_t)total, PF_SVC_NONE);
   std::vector<uint32_t> fill(off);  // next free slot per (tile, plane)
   for (i64 t = 0; t < ntile; t++) {
      const i64 ty = t / tzc, tz = t - ty * tzc;
      i64 ya, yb;
      shell_rows(ty, &ya, &yb);
      if (!with_shell || yb < ya || (tz != 0 && tz != tz_hi)) continue;
      for (i64 x = 1; x <= Nx - 2; x++) {
         if (on_xshell(x)) continue;
         uint32_t &f = fill[(size_t)(t * pitch + x)];
         for (i64 y = ya; y <= yb; y++) {
            const uint32_t r = (uint32_t)(y - 1 - ty * TY);
            if (tz == 0 && shell_node(x, y, 1)) list[f++] = (1u << 13) | (r << 7) | 1u;
            if (tz == tz_hi && shell_node(x, y, Nz - 2)) list[f++] = (1u << 13) | (r << 7) | (uint32_t)(Nz - 2 - tz * TZ);
         }
      }
   }
   std::vector<i64> bl;
   std::vector<uint16_t> al;
   for (i64 i = 0; i < e->Nb; i++) {
      i64 r, c;
      const i64 key = key_of(bn[(size_t)i], &r, &c);
      if (cnt[(size_t)key] <= cap)
         list[fill[(size_t)key]++] = ((uint32_t)(adj[(size_t)i] & 0xfffu) << 16) | ((uint32_t)r << 7) | (uint32_t)c;
      else
         bl.push_back(bn[(size_t)i]), al.push_back(adj[(size_t)i]);
   }
   int rc;
   if ((rc = upload_raw(e, &e->svc_list, list.data(), (i64)list.size(), "svc_list"))) return rc;
   if ((rc = upload_raw(e, &e->svc_off, off.data(), (i64)off.size(), "svc_off"))) return rc;
   if ((rc = upload_raw(e, &e->bn_left, bl.data(), (i64)bl.size(), "bn_left"))) return rc;
   if ((rc = upload_raw(e, &e->adj_left, al.data(), (i64)al.size(), "adj_left"))) return rc;
   e->Nb_left = (i64)bl.size();
   e->svc_entries = (i64)total;
   if (e->sorted) {
      const i64 P = Ny * Nzp;
      e->nbL_lo = std::lower_bound(bl.begin(), bl.end(), 2 * P) - bl.begin();
      e->nbL_hi = bl.end() - std::lower_bound(bl.begin(), bl.end(), (Nx - 2) * P);
   }
   e->tma.sv = pf::AirSvc{e->svc_list, e->svc_off, (int)pitch};
   e->svc_on = 1;
   return 0;
}

static int create_impl(const pffdtd_desc *d, int device, pffdtd_engine *e) {
   if (d->precision != 1 && d->precision != 2) return fail(PFFDTD_EINVAL, "precision must be 1 or 2");
   if (d->fcc_flag < 0 || d->fcc_flag > 2) return fail(PFFDTD_EINVAL, "fcc_flag must be 0, 1 or 2");
   if (d->Nx < 3 || d->Ny < 3 || d->Nz < 3) return fail(PFFDTD_EINVAL, "grid dims must be >= 3");
   if (d->Nb < 0 || d->Nbl < 0 || d->Nba < 0 || d->Ns < 0 || d->Nr < 0 || d->Nt < 0 || d->Nm < 0 || d->Nm > PFFDTD_MNM)
      return fail(PFFDTD_EINVAL, "negative count or too many materials");
   if (d->Nbl > 0 && (!d->Mb || !d->mat_beta || !d->mat_quads || !d->mat_bnl || !d->ssaf_bnl))
      return fail(PFFDTD_EINVAL, "lossy nodes without material tables");
   int ndev = 0;
   CU(cudaGetDeviceCount(&ndev));
   if (device < 0 || device >= ndev) return fail(PFFDTD_ECUDA, "no CUDA device %d (%d visible)", device, ndev);
   CU(cudaSetDevice(device));
   e->device = device;
   e->precision = d->precision;
   e->rs = d->precision == 1 ? 4 : 8;
   e->fcc = d->fcc_flag;
   e->NN = d->fcc_flag ? 12 : 6;
   e->Nm = d->Nm;
   e->Nx = d->Nx, e->Ny = d->Ny, e->Nz = d->Nz;
   e->Nzp = (d->Nz + 31) / 32 * 32;
   e->mwpr = (e->Nzp / 32 + 3) / 4 * 4;  // mask words per row, a multiple of 16 bytes (TMA stride)
   e->Nb = d->Nb, e->Nbl = d->Nbl, e->Nba = d->Nba, e->Ns = d->Ns, e->Nr = d->Nr, e->Nt = d->Nt;
   e->Nblp = (d->Nbl + 31) / 32 * 32;
   e->ix0 = d->ix0, e->x_lo_edge = d->x_lo_edge, e->x_hi_edge = d->x_hi_edge;
   e->l = d->l, e->l2 = d->l2, e->a1 = d->a1, e->a2 = d->a2, e->sl2 = d->sl2, e->lo2 = d->lo2;
   for (i64 i = 0; i < d->Nbl; i++) {
      const int k = d->mat_bnl[i];
      if (k < 0 || k >= d->Nm) return fail(PFFDTD_EINVAL, "mat_bnl[%lld]=%d out of range", (long long)i, k);
   }
   for (int k = 0; k < d->Nm; k++) {
      if (d->Mb[k] < 0 || d->Mb[k] > PFFDTD_MMB) return fail(PFFDTD_EINVAL, "Mb[%d]=%d out of range", k, d->Mb[k]);
      e->mb_max = std::max<int>(e->mb_max, d->Mb[k]);
   }


   const i64 sx = e->Ny * e->Nzp, sy = e->Nzp;
   if (e->fcc == 0) {
      const i64 o[6] = {sx, -sx, sy, -sy, 1, -1};
      for (int j = 0; j < 6; j++) e->off.o[j] = o[j];
   } else {
      const i64 o[12] = {sx + sy, -sx - sy, sy + 1, -sy - 1, sx + 1, -sx - 1, sx - sy, -sx + sy, sy - 1, -sy + 1, sx - 1, -sx + 1};
      for (int j = 0; j < 12; j++) e->off.o[j] = o[j];
   }

   CU(cudaStreamCreateWithFlags(&e->s_main, cudaStreamNonBlocking));
   CU(cudaStreamCreateWithFlags(&e->s_comm, cudaStreamNonBlocking));
   CU(cudaStreamCreateWithFlags(&e->s_abc, cudaStreamNonBlocking));
   CU(cudaStreamCreateWithFlags(&e->s_edge, cudaStreamNonBlocking));
   CU(cudaEventCreateWithFlags(&e->ev_e0, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_e1, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_abc0, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_abc1, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_edge, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_step, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&e->ev_push, cudaEventDisableTiming));
   CU(cudaEventCreate(&e->ev_t0));
   CU(cudaEventCreate(&e->ev_t1));

   const size_t npad = (size_t)(e->Nx * e->Ny * e->Nzp);
   if (dalloc_bytes(e, &e->u[0], npad * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->u[1], npad * e->rs)) return PFFDTD_ECUDA;
   if (dalloc(e, &e->mask, (size_t)(e->Nx * e->Ny * e->mwpr))) return PFFDTD_ECUDA;

   int rc;
   if ((rc = upload_idx(e, &e->bn, d->bn_ixyz, e->Nb, "bn_ixyz", true))) return rc;
   if ((rc = upload_idx(e, &e->bnl, d->bnl_ixyz, e->Nbl, "bnl_ixyz", true))) return rc;
   if ((rc = upload_idx(e, &e->bna, d->bna_ixyz, e->Nba, "bna_ixyz", true))) return rc;
   if ((rc = upload_idx(e, &e->in, d->in_ixyz, e->Ns, "in_ixyz", true))) return rc;
   if ((rc = upload_idx(e, &e->out, d->out_ixyz, e->Nr, "out_ixyz", false))) return rc;
   if ((rc = upload_raw(e, &e->adj, d->adj_bn, e->Nb, "adj_bn"))) return rc;
   if ((rc = upload_raw(e, &e->mat, d->mat_bnl, e->Nbl, "mat_bnl"))) return rc;
   if ((rc = upload_raw(e, &e->Q, d->Q_bna, e->Nba, "Q_bna"))) return rc;
   if ((rc = upload_raw(e, &e->Mb, d->Mb, (i64)e->Nm, "Mb"))) return rc;
   if ((rc = upload_real(e, &e->ssaf, d->ssaf_bnl, (size_t)e->Nbl))) return rc;
   if ((rc = upload_real(e, &e->beta, d->mat_beta, (size_t)e->Nm))) return rc;
   if ((rc = upload_real(e, &e->quads, d->mat_quads, (size_t)e->Nm * PFFDTD_MMB * 4))) return rc;

   // source samples: (Real)in_sigs (cpu_engine.h:312), stored step-major [Nt][Ns]
   {
      const size_t n = (size_t)(e->Ns * e->Nt);
      std::vector<double> t(n);
      for (i64 s = 0; s < e->Ns; s++)
         for (i64 k = 0; k < e->Nt; k++) t[(size_t)(k * e->Ns + s)] = d->in_sigs ? d->in_sigs[s * e->Nt + k] : 0.0;
      if ((rc = upload_real(e, &e->insig, t.data(), n))) return rc;
   }
   if (dalloc_bytes(e, &e->uout, (size_t)(e->Nr * std::max<i64>(e->Nt, 1)) * e->rs)) return PFFDTD_ECUDA;
   for (int k = 0; k < 2; k++)
      if (dalloc_bytes(e, &e->hist[k], (size_t)e->Nbl * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->u2ba, (size_t)e->Nba * e->rs)) return PFFDTD_ECUDA;
   // branch state, v and g interleaved in groups of 32 nodes (kernels.cuh st_idx); gh1 = the same buffer, 32 elements on
   if (dalloc_bytes(e, &e->vh1, (size_t)e->Nblp * PFFDTD_MMB * 2 * e->rs)) return PFFDTD_ECUDA;
   e->gh1 = (char *)e->vh1 + 32 * e->rs;
   if (dalloc_bytes(e, &e->lo2Kbg, (size_t)e->Nbl * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->facb, (size_t)e->Nbl * e->rs)) return PFFDTD_ECUDA;
   if (dalloc(e, &e->matmb, (size_t)e->Nbl)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->zold, (size_t)(e->Nx * e->Ny * 2) * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->yold, (size_t)(e->Nx * 2 * e->Nzp) * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->xold, (size_t)(2 * e->Ny * e->Nzp) * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->in_stage, (size_t)e->Ns * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->out_stage, (size_t)e->Nr * e->rs)) return PFFDTD_ECUDA;
   CU(cudaMallocHost(&e->h_in, std::max<size_t>((size_t)e->Ns * 8, 64)));
   CU(cudaMallocHost(&e->h_out, std::max<size_t>((size_t)e->Nr * 8, 64)));

   // duplicate source nodes keep the reference's serial accumulation order
   {
      std::vector<int64_t> t(d->in_ixyz, d->in_ixyz + e->Ns);
      std::sort(t.begin(), t.end());
      e->serial_src = std::adjacent_find(t.begin(), t.end()) != t.end();
   }
   // edge/interior split for the overlapped halo exchange needs ascending lists (gpu_engine.h:497-513)
   // (duplicate source nodes are legal: non-decreasing is enough there, split_data only walks the list, and `serial_src`
   // keeps their accumulation order)
   if (ascending(d->bn_ixyz, e->Nb) && ascending(d->bnl_ixyz, e->Nbl) && ascending(d->bna_ixyz, e->Nba) &&
       ascending(d->in_ixyz, e->Ns, false)) {
      const i64 P = e->Ny * e->Nz;
      e->sorted = 1;
      edge_counts(d->bn_ixyz, e->Nb, 2 * P, (e->Nx - 2) * P, &e->nb_lo, &e->nb_hi);
      edge_counts(d->bnl_ixyz, e->Nbl, 2 * P, (e->Nx - 2) * P, &e->nbl_lo, &e->nbl_hi);
      edge_counts(d->bna_ixyz, e->Nba, 2 * P, (e->Nx - 2) * P, &e->nba_lo, &e->nba_hi);
      edge_counts(d->in_ixyz, e->Ns, 2 * P, (e->Nx - 2) * P, &e->ns_lo, &e->ns_hi);
   }

   // per-node constants of the frequency-dependent boundary
   if (e->Nbl) {
      pf::MatTable mt{e->quads, e->beta, e->Mb};
      const unsigned g = (unsigned)((e->Nbl + 127) / 128);
      if (e->precision == 1)
         pf::k_fd_prep<float><<<g, 128, 0, e->s_main>>>(e->mat, (const float *)e->ssaf, e->Nbl, (float)e->lo2, mt, (float *)e->lo2Kbg,
                                                        (float *)e->facb, e->matmb);
      else
         pf::k_fd_prep<double><<<g, 128, 0, e->s_main>>>(e->mat, (const double *)e->ssaf, e->Nbl, e->lo2, mt, (double *)e->lo2Kbg,
                                                         (double *)e->facb, e->matmb);
      CU(cudaGetLastError());
   }
   // can the tiled kernel apply the absorbing shell itself?  Only if the list handed in is exactly the
   // canonical shell of this slab (fdtd_data.h:620-675): every interior node with an index of 1 or N-2 -- on the checkerboard FCC grid
   // the even-parity ones, on the folded FCC grid every stored node, whose y shell is row 1 alone (both physical y faces fold onto
   // it; the high end of the folded grid is the seam).
   {
      const i64 Nx = e->Nx, Ny = e->Ny, Nz = e->Nz;
      const bool fold = e->fcc == 2, checker = e->fcc == 1;
      auto qx = [&](i64 ix) { return ((e->x_lo_edge && ix == 1) || (e->x_hi_edge && ix == Nx - 2)) ? 1 : 0; };
      auto qy = [&](i64 iy) { return (iy == 1 || (!fold && iy == Ny - 2)) ? 1 : 0; };
      auto qz = [&](i64 iz) { return (iz == 1 || iz == Nz - 2) ? 1 : 0; };
      auto even = [&](i64 ix, i64 iy, i64 iz) { return !checker || ((e->ix0 + ix + iy + iz) & 1) == 0; };
      i64 expect = 0;
      for (i64 ix = 1; ix <= Nx - 2; ix++)
         for (i64 iy = 1; iy <= Ny - 2; iy++) {
            if (qx(ix) || qy(iy)) {
               if (!checker) expect += Nz - 2;
               else
                  for (i64 iz = 1; iz <= Nz - 2; iz++) expect += even(ix, iy, iz) ? 1 : 0;
            } else {
               expect += (even(ix, iy, 1) ? 1 : 0) + ((Nz - 2 != 1 && even(ix, iy, Nz - 2)) ? 1 : 0);
            }
         }
      bool ok = (Ny >= 4 && Nz >= 4) && expect == e->Nba && ascending(d->bna_ixyz, e->Nba);
      for (i64 i = 0; ok && i < e->Nba; i++) {
         const i64 v = d->bna_ixyz[i], row = v / Nz, iz = v - row * Nz, ix = row / Ny, iy = row - ix * Ny;
         const int Q = qx(ix) + qy(iy) + qz(iz);
         ok = Q > 0 && Q == d->Q_bna[i] && even(ix, iy, iz);
      }
      e->fuse_ok = ok;
      // do boundary / source nodes stay clear of the shell (the usual case; the reference's Python engine assumes it,
      // sim_fdtd.py:152)?  Then the shell update commutes with the boundary kernels.
      auto on_shell = [&](i64 v) {
         const i64 row = v / Nz, iz = v - row * Nz, ix = row / Ny, iy = row - ix * Ny;
         return qx(ix) || qy(iy) || qz(iz);
      };
      bool clear = ok;
      for (i64 i = 0; clear && i < e->Nb; i++) clear = !on_shell(d->bn_ixyz[i]);
      for (i64 i = 0; clear && i < e->Ns; i++) clear = !on_shell(d->in_ixyz[i]);
      e->abc_disjoint = clear;
      // z halos written by k_abc (unfused step): needs the canonical shell (every row's z ends are in the list; not the checkerboard
      // layout, whose odd-parity nodes are not) and nothing else writing the mirrored nodes z = 2 / Nz-3 afterwards
      bool zok = ok && !checker && Nz >= 7;
      auto z_src = [&](i64 v) {
         const i64 iz = v % Nz;
         return iz == 2 || iz == Nz - 3;
      };
      for (i64 i = 0; zok && i < e->Nb; i++) zok = !z_src(d->bn_ixyz[i]);
      for (i64 i = 0; zok && i < e->Ns; i++) zok = !z_src(d->in_ixyz[i]);
      e->zflip_ok = zok;
      if (zok && e->Nba) {
         std::vector<uint8_t> zf((size_t)e->Nba, 0);
         for (i64 i = 0; i < e->Nba; i++) {
            const i64 v = d->bna_ixyz[i], row = v / Nz, iz = v - row * Nz, ix = row / Ny, iy = row - ix * Ny;
            const bool shell_row = qx(ix) || qy(iy);
            zf[(size_t)i] = shell_row ? (uint8_t)((iz == 2 ? 4 : 0) | (iz == Nz - 3 ? 8 : 0)) : (uint8_t)((iz == 1 ? 1 : 0) | (iz == Nz - 2 ? 2 : 0));
         }
         if (dalloc(e, &e->zf, zf.size())) return PFFDTD_ECUDA;
         CU(cudaMemcpy(e->zf, zf.data(), zf.size(), cudaMemcpyHostToDevice));
      }
   }
   // does any boundary node also sit in the absorbing-shell list?  (a bitmap of the shell list; any layout, any order)
   {
      std::vector<uint8_t> bits((size_t)((e->Nx * e->Ny * e->Nz + 7) / 8), 0);
      for (i64 i = 0; i < e->Nba; i++) bits[(size_t)(d->bna_ixyz[i] >> 3)] |= (uint8_t)(1u << (d->bna_ixyz[i] & 7));
      bool off = true;
      for (i64 i = 0; off && i < e->Nb; i++) off = !((bits[(size_t)(d->bn_ixyz[i] >> 3)] >> (d->bn_ixyz[i] & 7)) & 1u);
      e->bn_off_abc = off;
   }
   // late halo mirrors: nodes written after the air kernel (boundary, source) whose value belongs in a halo (all the combinations:
   // the 13-point stencil reads halo edges; on the folded grid row Ny-2 goes to the seam row Ny-1 and the high y end has no mirror)
   {
      std::vector<std::pair<i64, i64>> pr;  // (src, dst) in the padded layout
      auto add_node = [&](i64 v) {
         const i64 row = v / e->Nz, iz = v - row * e->Nz, ix = row / e->Ny, iy = row - ix * e->Ny;
         i64 ox[3], oy[3], oz[3];
         int nx = 0, ny = 0, nz = 0;
         ox[nx++] = ix;
         if (e->x_lo_edge && ix == 2) ox[nx++] = 0;
         if (e->x_hi_edge && ix == e->Nx - 3) ox[nx++] = e->Nx - 1;
         oy[ny++] = iy;
         if (iy == 2) oy[ny++] = 0;
         if (e->fcc != 2 && iy == e->Ny - 3) oy[ny++] = e->Ny - 1;
         if (e->fcc == 2 && iy == e->Ny - 2) oy[ny++] = e->Ny - 1;
         oz[nz++] = iz;
         if (iz == 2) oz[nz++] = 0;
         if (iz == e->Nz - 3) oz[nz++] = e->Nz - 1;
         const i64 src = (ix * e->Ny + iy) * e->Nzp + iz;
         // a masked node at z = 1 / Nz-2 can make the whole 16-byte vector that holds the z halo masked (fp64: [halo, node]); the air
         // kernel then never stores that vector and the halo misses its mirror-on-write: copy it from z = 2 / Nz-3 afterwards
         if (iz == 1) pr.push_back({src + 1, src - 1});
         if (iz == e->Nz - 2) pr.push_back({src - 1, src + 1});
         for (int a = 0; a < nx; a++)
            for (int b = 0; b < ny; b++)
               for (int c = 0; c < nz; c++)
                  if (a | b | c) pr.push_back({src, (ox[a] * e->Ny + oy[b]) * e->Nzp + oz[c]});
      };
      for (i64 i = 0; i < e->Nb; i++) add_node(d->bn_ixyz[i]);
      for (i64 i = 0; i < e->Ns; i++) add_node(d->in_ixyz[i]);
      std::sort(pr.begin(), pr.end());
      pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
      e->np = (i64)pr.size();
      std::vector<i64> ps(pr.size()), pd(pr.size());
      const i64 P = e->Ny * e->Nzp;
      for (size_t i = 0; i < pr.size(); i++) {
         ps[i] = pr[i].first, pd[i] = pr[i].second;
         if (pr[i].first < 2 * P) e->np_lo++;
         if (pr[i].first >= (e->Nx - 2) * P) e->np_hi++;
      }
      if ((rc = upload_raw(e, &e->pair_src, ps.data(), e->np, "pair_src"))) return rc;
      if ((rc = upload_raw(e, &e->pair_dst, pd.data(), e->np, "pair_dst"))) return rc;
   }

   // mask
   {
      const i64 words = e->Nx * e->Ny * e->mwpr;
      pf::k_mask_init<<<(unsigned)((words + 255) / 256), 256, 0, e->s_main>>>(e->mask, e->Nx, e->Ny, e->Nz, e->mwpr, e->fcc, e->ix0);
      if (e->Nb) pf::k_mask_nodes<<<(unsigned)((e->Nb + 255) / 256), 256, 0, e->s_main>>>(e->mask, e->bn, e->Nb, e->Nzp, e->mwpr);
      CU(cudaGetLastError());
   }
   if (dalloc(e, &e->tma.ctr, 4)) return PFFDTD_ECUDA;  // two {next item, CTAs done} pairs: launches on s_main / on s_edge
   if (dalloc(e, &e->d_n, 1)) return PFFDTD_ECUDA;
   const int svc_cfg = want_svc(e) && e->bn_off_abc && e->Nb > 0;
   if ((rc = pf::air_tma_setup(&e->tma, e->precision, e->fcc, e->Nx, e->Ny, e->Nz, e->Nzp, e->mwpr, e->u[0], e->u[1], e->mask, -1, svc_cfg))) {
      // not fatal: fall back to the generic kernel, remember why
      e->air_kernel = 0;
   }
   CU(cudaStreamSynchronize(e->s_main));
   if ((rc = build_service(e))) return rc;
   return PFFDTD_OK;
}

extern "C" int pffdtd_create(const pffdtd_desc *desc, int device, pffdtd_engine **out) {
   if (!desc || !out) return fail(PFFDTD_EINVAL, "NULL argument");
   if (desc->struct_size != (int32_t)sizeof(pffdtd_desc))
      return fail(PFFDTD_EINVAL, "pffdtd_desc size mismatch: caller %d, library %d", desc->struct_size, (int)sizeof(pffdtd_desc));
   pffdtd_engine *e = new pffdtd_engine();
   int rc = create_impl(desc, device, e);
   if (rc != PFFDTD_OK) {
      std::string keep = g_err;
      pffdtd_destroy(e);
      g_err = keep;
      return rc;
   }
   *out = e;
   return PFFDTD_OK;
}

// ------------------------------------------------------------------------------------------------
// comm
// ------------------------------------------------------------------------------------------------
extern "C" int pffdtd_comm_unique_id(void *id128) {
   if (!id128) return fail(PFFDTD_EINVAL, "NULL id");
   if (nccl_load()) return PFFDTD_ENCCL;
   NC(g_nccl.GetUniqueId(id128));
   return PFFDTD_OK;
}

extern "C" int pffdtd_comm_init(pffdtd_engine *e, const void *id128, int rank, int nranks) {
   if (!e || !id128 || rank < 0 || rank >= nranks) return fail(PFFDTD_EINVAL, "bad comm arguments");
   if (nccl_load()) return PFFDTD_ENCCL;
   CU(cudaSetDevice(e->device));
   if ((rank > 0) == (e->x_lo_edge != 0) || (rank < nranks - 1) == (e->x_hi_edge != 0))
      return fail(PFFDTD_EINVAL, "slab edges (%d,%d) disagree with rank %d of %d", e->x_lo_edge, e->x_hi_edge, rank, nranks);
   Id128 id;
   memcpy(id.b, id128, 128);
   NC(g_nccl.CommInitRank(&e->comm, nranks, id, rank));
   e->rank = rank;
   e->nranks = nranks;
   return PFFDTD_OK;
}


// ------------------------------------------------------------------------------------------------
// halo exchange between processes over peer memory (CUDA IPC), without NCCL in the step
// ------------------------------------------------------------------------------------------------
static void drop_graphs(pffdtd_engine *e);
struct PeerBlob {
   cudaIpcMemHandle_t u[2], flags;
   int64_t Nx, Ny, Nzp;
   int32_t rs, device;
   char pad[PFFDTD_PEER_BLOB - 3 * sizeof(cudaIpcMemHandle_t) - 3 * 8 - 2 * 4];
};
static_assert(sizeof(PeerBlob) == PFFDTD_PEER_BLOB, "PeerBlob size");

extern "C" int pffdtd_peer_export(pffdtd_engine *e, void *blob) {
   if (!e || !blob) return fail(PFFDTD_EINVAL, "NULL argument");
   CU(cudaSetDevice(e->device));
   if (!e->flags) {
      if (dalloc(e, &e->flags, 4)) return PFFDTD_ECUDA;
   }
   PeerBlob b;
   memset(&b, 0, sizeof b);
   CU(cudaIpcGetMemHandle(&b.u[0], e->u[0]));
   CU(cudaIpcGetMemHandle(&b.u[1], e->u[1]));
   CU(cudaIpcGetMemHandle(&b.flags, e->flags));
   b.Nx = e->Nx, b.Ny = e->Ny, b.Nzp = e->Nzp, b.rs = (int32_t)e->rs, b.device = e->device;
   memcpy(blob, &b, sizeof b);
   return PFFDTD_OK;
}

extern "C" int pffdtd_peer_connect(pffdtd_engine *e, const void *blob_lo, const void *blob_hi) {
   if (!e) return fail(PFFDTD_EINVAL, "NULL engine");
   if ((blob_lo != nullptr) == (e->x_lo_edge != 0) || (blob_hi != nullptr) == (e->x_hi_edge != 0))
      return fail(PFFDTD_EINVAL, "neighbour blobs disagree with the slab's edges (%d,%d)", e->x_lo_edge, e->x_hi_edge);
   if (!e->flags) return fail(PFFDTD_ESTATE, "call pffdtd_peer_export first");
   CU(cudaSetDevice(e->device));
   auto open = [&](const void *blob, void **u, long long **fl, i64 *nx) -> int {
      PeerBlob b;
      memcpy(&b, blob, sizeof b);
      if (b.Ny != e->Ny || b.Nzp != e->Nzp || b.rs != (int32_t)e->rs) return fail(PFFDTD_EINVAL, "neighbour slab has another plane shape");
      for (int k = 0; k < 2; k++) CU(cudaIpcOpenMemHandle(&u[k], b.u[k], cudaIpcMemLazyEnablePeerAccess));
      CU(cudaIpcOpenMemHandle((void **)fl, b.flags, cudaIpcMemLazyEnablePeerAccess));
      if (nx) *nx = b.Nx;
      return 0;
   };
   int rc;
   if (blob_lo && (rc = open(blob_lo, e->ipc_u_lo, &e->ipc_flags_lo, &e->nx_lo))) return rc;
   if (blob_hi && (rc = open(blob_hi, e->ipc_u_hi, &e->ipc_flags_hi, nullptr))) return rc;
   drop_graphs(e);
   e->p2p = 1;
   return PFFDTD_OK;
}

// ------------------------------------------------------------------------------------------------
// options / stats
// ------------------------------------------------------------------------------------------------
static void drop_graphs(pffdtd_engine *e) {
   for (int c = 0; c < 2; c++) {
      if (e->graph[c]) cudaGraphExecDestroy(e->graph[c]);
      if (e->hgraph[c]) cudaGraphExecDestroy(e->hgraph[c]);
      if (e->graphA[c]) cudaGraphExecDestroy(e->graphA[c]);
      if (e->graphB[c]) cudaGraphExecDestroy(e->graphB[c]);
      e->graph[c] = e->hgraph[c] = e->graphA[c] = e->graphB[c] = nullptr;
   }
}

extern "C" int pffdtd_set_option(pffdtd_engine *e, const char *key, int64_t value) {
   if (!e || !key) return fail(PFFDTD_EINVAL, "NULL argument");
   std::string k(key);
   // any option may change what a step launches: drop the captured graphs
   drop_graphs(e);
   e->steps_plain = 0;
   if (k == "use_graph") {
      e->use_graph = value != 0;
   } else if (k == "air_kernel") {
      if (value == 1 && !e->tma.ok) return fail(PFFDTD_ESTATE, "tiled air kernel unavailable: %s", e->tma.why.c_str());
      if (value < 0 || value > 1) return fail(PFFDTD_EINVAL, "air_kernel must be 0 or 1");
      e->air_kernel = (int)value;
      e->halo_dirty = 1;
      e->zhalo_ok = 0;
      int rc = build_service(e);  // (whether the lists carry the shell's z faces follows the kind of step)
      if (rc) return rc;
   } else if (k == "fuse") {
      e->fuse = value < 0 ? -1 : (value != 0);
      e->halo_dirty = 1;
      e->zhalo_ok = 0;
      int rc = build_service(e);
      if (rc) return rc;
   } else if (k == "zflip") {
      e->zflip_want = value != 0;
      e->zhalo_ok = 0;
   } else if (k == "overlap") {
      e->overlap = value != 0;
   } else if (k == "profile_air") {
      e->profile_air = value != 0;
   } else if (k == "air_xc") {
      e->tma.xc = (int)value;
   } else if (k == "air_cfg" || k == "svc" || k == "svc_cap") {
      // tile configuration (-1 = the default for this grid), in-kernel boundary work on / off, its density threshold: all three
      // decide the shape of the service lists
      CU(cudaSetDevice(e->device));
      CU(cudaStreamSynchronize(e->s_main));
      int cfg = e->tma.cfg;
      if (k == "svc") e->svc_want = value < 0 ? -1 : (value != 0), cfg = -1;
      else if (k == "svc_cap") e->svc_cap = (int)std::max<int64_t>(0, std::min<int64_t>(value, PF_SVC_CAP));
      else cfg = (int)value;
      const int svc_cfg = want_svc(e) && e->bn_off_abc && e->Nb > 0;
      if (k != "svc_cap" &&
          pf::air_tma_setup(&e->tma, e->precision, e->fcc, e->Nx, e->Ny, e->Nz, e->Nzp, e->mwpr, e->u[0], e->u[1], e->mask, cfg, svc_cfg))
         return fail(PFFDTD_EINVAL, "air_cfg %lld: %s", (long long)value, e->tma.why.c_str());
      int rc = build_service(e);
      if (rc) return rc;
      e->halo_dirty = 1;
      e->zhalo_ok = 0;
   } else if (k == "p2p") {
      if (value && !(e->ipc_flags_lo || e->ipc_flags_hi)) return fail(PFFDTD_ESTATE, "p2p needs pffdtd_peer_connect first");
      e->p2p = value != 0;
   } else if (k == "edge_overlap") {
      e->edge_overlap = value != 0;
   } else if (k == "fd_bulk") {
      e->fd_bulk = value != 0;
   } else if (k == "abc_overlap") {
      e->abc_overlap = value != 0;
   } else if (k == "manual_halo") {
      e->manual_halo = value != 0;
   } else {
      return fail(PFFDTD_EINVAL, "unknown option %s", key);
   }
   return PFFDTD_OK;
}

static int drain_air_events(pffdtd_engine *e) {
   for (size_t i = 0; i < e->air_ev_used; i++) {
      float ms = 0;
      CU(cudaEventSynchronize(e->air_ev[i].b));
      CU(cudaEventElapsedTime(&ms, e->air_ev[i].a, e->air_ev[i].b));
      e->air_ms += ms;
      e->air_timed++;
   }
   e->air_ev_used = 0;
   return 0;
}

extern "C" int pffdtd_get_stat(pffdtd_engine *e, const char *key, double *out) {
   if (!e || !key || !out) return fail(PFFDTD_EINVAL, "NULL argument");
   std::string k(key);
   if (k == "launches") *out = e->launches;
   else if (k == "steps") *out = (double)e->steps_done;
   else if (k == "air_ms" || k == "air_launches_timed") {
      CU(cudaSetDevice(e->device));
      if (drain_air_events(e)) return PFFDTD_ECUDA;
      *out = k == "air_ms" ? e->air_ms : (double)e->air_timed;
   } else if (k == "timer_start") {
      // device-side stopwatch on the engine's own stream (torch.cuda.Event would not see it)
      CU(cudaSetDevice(e->device));
      CU(cudaEventRecord(e->ev_t0, e->s_main));
      *out = 0;
   } else if (k == "timer_stop_ms") {
      CU(cudaSetDevice(e->device));
      if (e->comm_pending) CU(cudaStreamWaitEvent(e->s_main, e->ev_comm, 0));
      CU(cudaEventRecord(e->ev_t1, e->s_main));
      CU(cudaEventSynchronize(e->ev_t1));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
      *out = ms;
   } else if (k == "air_kernel") *out = e->air_kernel;
   else if (k == "air_cfg") *out = e->tma.ok ? e->tma.cfg : -1;
   else if (k == "air_lanes_z") {
      int rpt = 0, nw = 0, lz = 0;
      pf::air_cfg_shape(e->tma.cfg, &rpt, &nw, &lz);
      *out = e->tma.ok ? lz : 0;
   } else if (k == "Nzp") *out = (double)e->Nzp;
   else if (k == "zflip") *out = zflip_on(e) && !(step_fused(e) && (e->fcc == 0 || (e->svc_on && e->svc_shell))) ? 1 : 0;
   else if (k == "fused") *out = (step_fused(e) && (e->fcc == 0 || (e->svc_on && e->svc_shell))) ? 1 : 0;
   else if (k == "energy") *out = e->energy_on;
   else if (k == "mirror_pairs") *out = (double)e->np;
   else if (k == "abc_disjoint") *out = e->abc_disjoint;
   else if (k == "p2p") *out = e->p2p;
   else if (k == "svc") *out = e->svc_on;
   else if (k == "svc_entries") *out = (double)e->svc_entries;
   else if (k == "nb_left") *out = (double)e->Nb_left;
   else return fail(PFFDTD_EINVAL, "unknown stat %s", key);
   return PFFDTD_OK;
}

extern "C" int pffdtd_reset_stats(pffdtd_engine *e) {
   if (!e) return fail(PFFDTD_EINVAL, "NULL engine");
   CU(cudaSetDevice(e->device));
   if (drain_air_events(e)) return PFFDTD_ECUDA;
   e->launches = 0;
   e->air_ms = 0;
   e->air_timed = 0;
   return PFFDTD_OK;
}

// ------------------------------------------------------------------------------------------------
// one time step
// ------------------------------------------------------------------------------------------------
static inline unsigned nblk(i64 n, int b) { return (unsigned)((n + b - 1) / b); }

// a contiguous piece of one step's work: x-planes [xb,xe) and the node-list ranges that live in them
struct Part {
   i64 xb, xe, b0, nb, l0, nbl, a0, nba, s0, ns, p0, np;
   bool recv;  // this part also reads the receivers
};

template <typename Real>
struct Step {
   pffdtd_engine *e;
   Real *u1, *u0;
   cudaStream_t s;
   i64 n;
   bool fused;
   bool svc;  // the air kernel's service warp does the sparse rigid nodes and the shell's z faces; k_rigid gets the dense rest
   int *ctr = nullptr;  // work counters of the air launches of this chain (null = the engine's first pair)

   // 4. air update of planes [xb, xe)
   int air(i64 xb, i64 xe) {
      if (xe <= xb) return 0;
      const bool timed = e->profile_air;
      EvPair *ev = nullptr;
      if (timed) {
         if (e->air_ev_used == e->air_ev.size()) {
            if (e->air_ev.size() >= 4096) {
               if (drain_air_events(e)) return PFFDTD_ECUDA;
            } else {
               EvPair p;
               CU(cudaEventCreate(&p.a));
               CU(cudaEventCreate(&p.b));
               e->air_ev.push_back(p);
            }
         }
         ev = &e->air_ev[e->air_ev_used++];
         CU(cudaEventRecord(ev->a, s));
      }
      if (e->air_kernel == 1) {
         pf::AirEdge<Real> eg;
         memset(&eg, 0, sizeof eg);
         eg.fuse = fused, eg.x_lo = e->x_lo_edge, eg.x_hi = e->x_hi_edge, eg.Nx = (int)e->Nx;
         eg.zstash = svc ? 0 : 1, eg.sl2 = (Real)e->sl2, eg.negzero = e->negzero, eg.folded = e->fcc == 2;
         eg.zold = (Real *)e->zold, eg.yold = (Real *)e->yold, eg.xold = (Real *)e->xold;
         // cpu_engine.h:226-228: Real lQ = l*Q; ... /(1.0 + lQ)
         eg.lQ1 = (Real)((Real)e->l * (Real)1), eg.lQ2 = (Real)((Real)e->l * (Real)2), eg.lQ3 = (Real)((Real)e->l * (Real)3);
         eg.den1 = 1.0 + (double)eg.lQ1, eg.den2 = 1.0 + (double)eg.lQ2, eg.den3 = 1.0 + (double)eg.lQ3;
         eg.rden1 = 1.0 / eg.den1, eg.rden2 = 1.0 / eg.den2, eg.rden3 = 1.0 / eg.den3;
         int rc = pf::air_tma_launch<Real>(&e->tma, e->cur, u0, xb, xe, (Real)e->a1, (Real)e->a2, eg, svc, s, ctr);
         if (rc) return fail(PFFDTD_ECUDA, "tiled air kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
      } else {
         dim3 blk(64, 4, 1);
         dim3 grd(nblk(e->Nzp, 64), nblk(e->Ny, 4), (unsigned)(xe - xb));
         if (e->fcc)
            pf::k_air_generic<Real, 12><<<grd, blk, 0, s>>>(u1, u0, e->mask, e->Ny, e->Nzp, e->mwpr, xb, (Real)e->a1, (Real)e->a2, e->off);
         else
            pf::k_air_generic<Real, 6><<<grd, blk, 0, s>>>(u1, u0, e->mask, e->Ny, e->Nzp, e->mwpr, xb, (Real)e->a1, (Real)e->a2, e->off);
      }
      e->launches += 1;
      if (timed) CU(cudaEventRecord(ev->b, s));
      if (fused && e->air_kernel == 1) {
         // the absorbing shell, from the values the air kernel stashed (see k_abc_faces)
         pf::FacesArgs<Real> fa;
         fa.u0 = u0, fa.zold = (const Real *)e->zold, fa.yold = (const Real *)e->yold, fa.xold = (const Real *)e->xold;
         fa.Nx = (int)e->Nx, fa.Ny = (int)e->Ny, fa.Nz = (int)e->Nz, fa.Nzp = (int)e->Nzp, fa.xb = (int)xb, fa.xe = (int)xe;
         fa.x_lo = e->x_lo_edge, fa.x_hi = e->x_hi_edge, fa.do_z = svc ? 0 : 1, fa.folded = e->fcc == 2, fa.edges = e->fcc != 0;
         fa.checker = e->fcc == 1 ? 1 + (int)(e->ix0 & 1) : 0;
         fa.lQ1 = (Real)((Real)e->l * (Real)1), fa.lQ2 = (Real)((Real)e->l * (Real)2), fa.lQ3 = (Real)((Real)e->l * (Real)3);
         // lines: (xe-xb)*2 rows of the y faces, 2*Ny rows of the x faces, (when the service warp does not do them) (xe-xb)*2 lines of
         // the z faces along y
         const i64 lines = (xe - xb) * 2 + 2 * e->Ny + (svc ? 0 : (xe - xb) * 2);
         const dim3 grd(nblk(std::max(e->Nz, svc ? (i64)0 : e->Ny), 128), (unsigned)std::min<i64>(lines, 65535));
         if (e->abc_overlap && e->abc_disjoint && !e->comm && !e->p2p && !e->peer_lo && !e->peer_hi) {  // one GPU only: with slabs the edge parts order their work around the exchange
            // no boundary or source node lies on the shell: the shell update commutes with the boundary kernels
            CU(cudaEventRecord(e->ev_abc0, s));
            CU(cudaStreamWaitEvent(e->s_abc, e->ev_abc0, 0));
            pf::k_abc_faces<Real><<<grd, 128, 0, e->s_abc>>>(fa);
            CU(cudaEventRecord(e->ev_abc1, e->s_abc));
            e->abc_pending = 1;
         } else {
            pf::k_abc_faces<Real><<<grd, 128, 0, s>>>(fa);
         }
         e->launches += 1;
      }
      return 0;
   }
   // steps 4-9 for one part, in the reference's order: air, ABC, rigid, FD, receivers/sources, late mirrors
   int part(const Part &p) {
      int rc = air(p.xb, p.xe);
      if (rc) return rc;
      if (!fused && p.nba > 0) {
         pf::k_abc<Real><<<nblk(p.nba, 128), 128, 0, s>>>(u0, e->bna, e->Q, (const Real *)e->u2ba, p.a0, p.nba, (Real)e->l,
                                                          zflip_on(e) ? e->zf : nullptr);
         e->launches += 1;
      }
      if (p.nb > 0) {
         const i64 *bn = svc ? e->bn_left : e->bn;
         const uint16_t *adj = svc ? e->adj_left : e->adj;
         if (e->fcc)
            pf::k_rigid<Real, 12><<<nblk(p.nb, 128), 128, 0, s>>>(u1, u0, bn, adj, p.b0, p.nb, (Real)e->sl2, (Real)e->a2, e->off);
         else
            pf::k_rigid<Real, 6><<<nblk(p.nb, 128), 128, 0, s>>>(u1, u0, bn, adj, p.b0, p.nb, (Real)e->sl2, (Real)e->a2, e->off);
         e->launches += 1;
      }
      if (p.nbl > 0) {
         const int nq = e->Nm * PFFDTD_MMB * 4;
         const size_t sm = (size_t)(nq / 4 * 5) * sizeof(Real);
         if (e->fd_bulk) {
            // blocks own 4 whole groups of 32 nodes; the range may start / end inside a group
            const i64 gfirst = p.l0 >> 5, gend = (p.l0 + p.nbl + 31) >> 5;
            const unsigned nb = (unsigned)((gend - gfirst + 3) / 4);
            const size_t smb = (size_t)(4 * 2 * PFFDTD_MMB * 32 + (nq / 4 * 5 + 1) / 2 * 2) * sizeof(Real) + 16;
            pf::k_fd_bulk<Real, PFFDTD_MMB><<<nb, 128, smb, s>>>(u0, e->bnl, e->matmb, (const Real *)e->lo2Kbg, (const Real *)e->facb,
                                                              (Real *)e->hist[0], (Real *)e->hist[1], (Real *)e->vh1, p.l0, p.nbl,
                                                              (const Real *)e->quads, nq, e->d_n, e->mb_max);
         } else {
            pf::k_fd<Real, PFFDTD_MMB><<<nblk(p.nbl, 128), 128, sm, s>>>(u0, e->bnl, e->matmb, (const Real *)e->lo2Kbg, (const Real *)e->facb,
                                                                         (Real *)e->hist[0], (Real *)e->hist[1], (Real *)e->vh1, (Real *)e->gh1,
                                                                         p.l0, p.nbl, e->Nblp, (const Real *)e->quads, nq, e->d_n, e->mb_max);
         }
         e->launches += 1;
      }
      if (e->abc_pending) {  // join the absorbing-shell kernel running beside the boundary kernels
         CU(cudaStreamWaitEvent(s, e->ev_abc1, 0));
         e->abc_pending = 0;
      }
      const i64 nr = p.recv ? e->Nr : 0;
      if (nr > 0 || p.ns > 0) {
         // the step's last part (the one that reads the receivers), one block, no energy bookkeeping after it: k_io also ticks
         const unsigned nb = nblk(std::max(nr, p.ns), 128);
         const int tick = (p.recv && nb == 1 && !e->energy_on) ? 1 : 0;
         pf::k_io<Real><<<nb, 128, 0, s>>>(u1, u0, e->out, (Real *)e->uout, nr, e->Nr, e->in, (Real *)e->insig, e->Ns, p.s0, p.ns,
                                           e->serial_src, e->d_n, e->host_mode & 1 ? (const Real *)e->in_stage : nullptr,
                                           e->host_mode & 2 ? (Real *)e->out_stage : nullptr, tick);
         e->launches += 1;
         if (tick) e->ticked = 1;
      }
      if (fused && p.np > 0) {
         pf::k_pairs<Real><<<nblk(p.np, 128), 128, 0, s>>>(u0, e->pair_src, e->pair_dst, p.p0, p.np);
         e->launches += 1;
      }
      return 0;
   }
};

// ------------------------------------------------------------------------------------------------
// energy balance (energy.cuh; sim_fdtd.py:587-620)
// ------------------------------------------------------------------------------------------------
static inline int en_blocks(i64 n) { return (int)std::max<i64>(1, std::min<i64>(pf::EN_BLOCKS, (n + pf::EN_THREADS - 1) / pf::EN_THREADS)); }

// before the step's updates: u1 = state n (halos mirrored), u0 = state n-1, Lu = Laplacian of state n-1
template <typename Real>
static int energy_pre(pffdtd_engine *e, Real *u1, Real *u0, i64 n, cudaStream_t s) {
   double *P = e->en_part;
   const int EB = pf::EN_BLOCKS, nB = en_blocks(e->Nba), nC = en_blocks(e->Nbl);
   Real *Lu = (Real *)e->Lu;
   pf::k_energy_int<Real><<<EB, pf::EN_THREADS, 0, s>>>(u1, u0, Lu, e->Nx, e->Ny, e->Nz, e->Nzp, e->l2, P);
   pf::k_energy_abc_corr<Real><<<nB, pf::EN_THREADS, 0, s>>>(u1, u0, Lu, e->bna, e->Q, e->Nba, e->l2, P + EB);
   pf::k_energy_branches<Real, PFFDTD_MMB><<<nC, pf::EN_THREADS, 0, s>>>((const Real *)e->ssaf, e->matmb, (const Real *)e->vh1, (const Real *)e->gh1,
                                                                        e->Nbl, e->Nblp, e->en_def, e->en_Ts, 0, P + 2 * EB);
   pf::k_energy_finish_H<<<1, pf::EN_THREADS, 0, s>>>(P, P + EB, P + 2 * EB, EB, nB, nC, e->en_k, e->en_H, n);
   e->launches += 4;
   if (e->Ns) {
      pf::k_gather<Real><<<nblk(e->Ns, 128), 128, 0, s>>>(u0, e->in, (Real *)e->u2in, e->Ns);
      e->launches += 1;
   }
   if (e->Nbl) CU(cudaMemcpyAsync(e->vold, e->vh1, (size_t)e->Nblp * PFFDTD_MMB * 2 * e->rs, cudaMemcpyDeviceToDevice, s));
   // Lu <- Laplacian of state n, for the next step's H_tot (the reference keeps Lu1 the same way, sim_fdtd.py:603-604)
   const double lfac = e->fcc ? 0.25 : 1.0;
   dim3 blk(64, 4, 1), grd(nblk(e->Nzp, 64), nblk(e->Ny, 4), (unsigned)(e->Nx - 2));
   if (e->fcc) {
      pf::k_lap_air<Real, 12><<<grd, blk, 0, s>>>(u1, Lu, e->mask, e->Ny, e->Nzp, e->mwpr, 1, lfac, e->off);
      if (e->Nb) pf::k_lap_bn<Real, 12><<<nblk(e->Nb, 128), 128, 0, s>>>(u1, Lu, e->bn, e->adj, e->Nb, lfac, e->off);
   } else {
      pf::k_lap_air<Real, 6><<<grd, blk, 0, s>>>(u1, Lu, e->mask, e->Ny, e->Nzp, e->mwpr, 1, lfac, e->off);
      if (e->Nb) pf::k_lap_bn<Real, 6><<<nblk(e->Nb, 128), 128, 0, s>>>(u1, Lu, e->bn, e->adj, e->Nb, lfac, e->off);
   }
   e->launches += e->Nb ? 2 : 1;
   CU(cudaGetLastError());
   return 0;
}

// after the step's updates: u0 = state n+1, vh1 = the reference's vh0, vold = its vh1
template <typename Real>
static int energy_post(pffdtd_engine *e, Real *u0, i64 n, cudaStream_t s) {
   double *P = e->en_part;
   const int EB = pf::EN_BLOCKS, nD = en_blocks(e->Nbl), nE = en_blocks(e->Nba);
   pf::k_energy_branches<Real, PFFDTD_MMB><<<nD, pf::EN_THREADS, 0, s>>>((const Real *)e->ssaf, e->matmb, (const Real *)e->vh1, (const Real *)e->vold,
                                                                        e->Nbl, e->Nblp, e->en_def, e->en_Ts, 1, P + 3 * EB);
   pf::k_energy_abc_loss<Real><<<nE, pf::EN_THREADS, 0, s>>>(u0, (const Real *)e->u2ba, e->bna, e->Q, e->Nba, P + 4 * EB);
   pf::k_energy_in<Real><<<1, pf::EN_THREADS, 0, s>>>(u0, (const Real *)e->u2in, e->in, (const Real *)e->insig + n * e->Ns, e->Ns, P + 5 * EB);
   pf::k_energy_finish_E<<<1, pf::EN_THREADS, 0, s>>>(P + 3 * EB, P + 4 * EB, P + 5 * EB, nD, nE, e->en_k, e->en_lost, e->en_in, n);
   e->launches += 4;
   CU(cudaGetLastError());
   return 0;
}

extern "C" int pffdtd_energy_enable(pffdtd_engine *e, const pffdtd_energy_desc *d) {
   if (!e || !d) return fail(PFFDTD_EINVAL, "NULL argument");
   if (d->struct_size != (int32_t)sizeof(pffdtd_energy_desc)) return fail(PFFDTD_EINVAL, "pffdtd_energy_desc size mismatch");
   if (e->energy_on) return fail(PFFDTD_ESTATE, "energy balance already enabled");
   if (e->steps_done != 0) return fail(PFFDTD_ESTATE, "enable the energy balance before the first step");
   if (e->fcc == 2) return fail(PFFDTD_ESTATE, "no energy balance on folded FCC grids (fcc_flag 2): use the fcc_flag 1 folder");
   if (!(d->h > 0) || !(d->c > 0) || !(d->Ts > 0)) return fail(PFFDTD_EINVAL, "h, c, Ts must be positive");
   if (e->Nbl > 0 && !d->mat_DEF) return fail(PFFDTD_EINVAL, "lossy nodes but no mat_DEF");
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->s_main));
   const size_t npad = (size_t)(e->Nx * e->Ny * e->Nzp);
   if (dalloc_bytes(e, &e->Lu, npad * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->vold, (size_t)e->Nblp * PFFDTD_MMB * 2 * e->rs)) return PFFDTD_ECUDA;
   if (dalloc_bytes(e, &e->u2in, (size_t)e->Ns * e->rs)) return PFFDTD_ECUDA;
   if (dalloc(e, &e->en_def, (size_t)std::max(e->Nm, 1) * PFFDTD_MMB * 3)) return PFFDTD_ECUDA;
   if (e->Nm && d->mat_DEF) CU(cudaMemcpy(e->en_def, d->mat_DEF, (size_t)e->Nm * PFFDTD_MMB * 3 * 8, cudaMemcpyHostToDevice));
   if (dalloc(e, &e->en_part, (size_t)6 * pf::EN_BLOCKS)) return PFFDTD_ECUDA;
   if (dalloc(e, &e->en_H, (size_t)e->Nt + 1)) return PFFDTD_ECUDA;
   if (dalloc(e, &e->en_lost, (size_t)e->Nt + 1)) return PFFDTD_ECUDA;
   if (dalloc(e, &e->en_in, (size_t)e->Nt + 1)) return PFFDTD_ECUDA;
   e->en_k = pf::EnergyCoef{e->fcc ? 2.0 : 1.0, d->h, d->c, e->l, e->l2};
   e->en_Ts = d->Ts;
   drop_graphs(e);
   e->halo_dirty = 1;
   e->zhalo_ok = 0;
   e->energy_on = 1;
   return build_service(e);  // (energy steps use the list kernels only)
}

extern "C" int pffdtd_read_energy(pffdtd_engine *e, double *H_tot, double *E_lost, double *E_in) {
   if (!e || !H_tot || !E_lost || !E_in) return fail(PFFDTD_EINVAL, "NULL argument");
   if (!e->energy_on) return fail(PFFDTD_ESTATE, "energy balance not enabled");
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->s_main));
   if (e->Nt) CU(cudaMemcpy(H_tot, e->en_H, (size_t)e->Nt * 8, cudaMemcpyDeviceToHost));
   CU(cudaMemcpy(E_lost, e->en_lost, (size_t)(e->Nt + 1) * 8, cudaMemcpyDeviceToHost));
   CU(cudaMemcpy(E_in, e->en_in, (size_t)(e->Nt + 1) * 8, cudaMemcpyDeviceToHost));
   return PFFDTD_OK;
}

// the reference's mirror pass on one grid (cpu_engine.h:135-172): seam row, then z, y, x faces in that order
template <typename Real>
static void mirror_pass(pffdtd_engine *e, Real *u, cudaStream_t s, bool skip_z = false) {
   const i64 Nx = e->Nx, Ny = e->Ny, Nz = e->Nz, Nzp = e->Nzp;
   if (e->fcc == 2) {
      pf::k_fold_seam<Real><<<dim3(nblk(Nz, 128), (unsigned)std::min<i64>(Nx, 65535)), 128, 0, s>>>(u, Nx, Ny, Nz, Nzp);
      e->launches += 1;
   }
   if (!skip_z) {
      pf::k_flip_z<Real><<<nblk(Nx * Ny, 128), 128, 0, s>>>(u, Nx * Ny, Nz, Nzp);
      e->launches += 1;
   }
   pf::k_flip_y<Real><<<dim3(nblk(Nz, 128), (unsigned)std::min<i64>(Nx, 65535)), 128, 0, s>>>(u, Nx, Ny, Nz, Nzp, e->fcc != 2);
   e->launches += 1;
   if (e->x_lo_edge || e->x_hi_edge) {
      pf::k_flip_x<Real><<<dim3(nblk(Nz, 128), (unsigned)std::min<i64>(Ny, 65535)), 128, 0, s>>>(u, Nx, Ny, Nz, Nzp, e->x_lo_edge, e->x_hi_edge);
      e->launches += 1;
   }
}

// exchange of the new state's edge planes (unew = u0 before the swap):
// plane 1 -> lower neighbour's plane Nx-1, plane Nx-2 -> upper neighbour's plane 0.
// Replaces the four cudaMemcpyPeerAsync waves of gpu_engine.h:1086-1126.
static int exchange(pffdtd_engine *e, void *unew, cudaStream_t s) {
   const size_t pb = (size_t)(e->Ny * e->Nzp) * e->rs;
   char *g = (char *)unew;
   if (e->peer_lo || e->peer_hi) {
      // one host thread drives every slab (gpu_engine.h:994, 1086-1126): push the new edge planes into the neighbours' halo planes of
      // the grid with the same role (all slabs step in lockstep, so `cur` agrees).  The neighbour's next step waits for ev_comm.
      const int role = e->cur ^ 1;
      auto push = [&](pffdtd_engine *to, size_t dst_plane, size_t src_plane) -> cudaError_t {
         char *dst = (char *)to->u[role] + dst_plane * pb;
         if (to->device == e->device) return cudaMemcpyAsync(dst, g + src_plane * pb, pb, cudaMemcpyDeviceToDevice, s);
         return cudaMemcpyPeerAsync(dst, to->device, g + src_plane * pb, e->device, pb, s);
      };
      if (e->peer_lo) CU(push(e->peer_lo, (size_t)(e->peer_lo->Nx - 1), 1));
      if (e->peer_hi) CU(push(e->peer_hi, 0, (size_t)(e->Nx - 2)));
      return 0;
   }
   if (e->p2p) {
      const int role = e->cur ^ 1;
      if (e->ipc_u_lo[role]) {
         CU(cudaMemcpyAsync((char *)e->ipc_u_lo[role] + (size_t)(e->nx_lo - 1) * pb, g + pb, pb, cudaMemcpyDeviceToDevice, s));
         CU(cudaMemcpyAsync(e->ipc_flags_lo + 1, e->flags + 2, 8, cudaMemcpyDeviceToDevice, s));  // I am its upper neighbour
      }
      if (e->ipc_u_hi[role]) {
         CU(cudaMemcpyAsync((char *)e->ipc_u_hi[role], g + (size_t)(e->Nx - 2) * pb, pb, cudaMemcpyDeviceToDevice, s));
         CU(cudaMemcpyAsync(e->ipc_flags_hi + 0, e->flags + 2, 8, cudaMemcpyDeviceToDevice, s));  // I am its lower neighbour
      }
      CU(cudaGetLastError());
      return 0;
   }
   if (!e->comm) return 0;
   NC(g_nccl.GroupStart());
   if (!e->x_lo_edge) {
      NC(g_nccl.Send(g + pb, pb, /*ncclInt8*/ 0, e->rank - 1, e->comm, s));
      NC(g_nccl.Recv(g, pb, 0, e->rank - 1, e->comm, s));
   }
   if (!e->x_hi_edge) {
      NC(g_nccl.Send(g + (size_t)(e->Nx - 2) * pb, pb, 0, e->rank + 1, e->comm, s));
      NC(g_nccl.Recv(g + (size_t)(e->Nx - 1) * pb, pb, 0, e->rank + 1, e->comm, s));
   }
   NC(g_nccl.GroupEnd());
   return 0;
}

// Phases of a step: PH_A the work before the halo exchange (with slabs: the two edge planes and everything that lives in them),
// PH_X the exchange itself, PH_B the rest.  One call normally does all three.  A slab with an NCCL communicator replays PH_A and
// PH_B from captured graphs and issues PH_X eagerly between them: NCCL send/recv captured INSIDE a graph hung on the 2-GPU box
// (NCCL 2.28.9, B200), so the exchange stays outside the graphs.
enum { PH_A = 1, PH_X = 2, PH_B = 4, PH_ALL = 7 };

static bool step_split(const pffdtd_engine *e) {
   const bool linked = e->comm || e->p2p || e->peer_lo || e->peer_hi;
   return linked && (!e->x_lo_edge || !e->x_hi_edge) && e->overlap && e->sorted && e->Nx >= 5;
}

template <typename Real>
static int step_impl(pffdtd_engine *e, i64 n, int phases) {
   if (n < 0 || n >= e->Nt) return fail(PFFDTD_EINVAL, "step %lld outside [0,%lld)", (long long)n, (long long)e->Nt);
   const bool fused = step_fused(e) && (e->fcc == 0 || (e->svc_on && e->svc_shell));
   const bool svc = e->svc_on && e->air_kernel == 1 && (e->svc_shell != 0) == fused;
   Step<Real> st{e, (Real *)e->u[e->cur], (Real *)e->u[e->cur ^ 1], e->s_main, n, fused, svc};
   const i64 NB = svc ? e->Nb_left : e->Nb, nb_lo = svc ? e->nbL_lo : e->nb_lo, nb_hi = svc ? e->nbL_hi : e->nb_hi;
   Real *u1 = st.u1, *u0 = st.u0;
   cudaStream_t s = e->s_main;
   const i64 Nx = e->Nx;
   int rc;
   const bool linked = e->comm || e->p2p || e->peer_lo || e->peer_hi;
   const bool lo = linked && !e->x_lo_edge, hi = linked && !e->x_hi_edge;
   const bool split = step_split(e);
   // the step's opening belongs to PH_A when the step is split around the exchange, else to PH_B (PH_A is then empty)
   if (phases & (split ? PH_A : PH_B)) {
      if (e->n_dev != n) {
         pf::k_set_n<<<1, 1, 0, s>>>(e->d_n, n);
         e->launches += 1;
         e->n_dev = n;
      }
      // the halo planes of u1 come from the previous step's exchange
      if (e->comm_pending) {
         CU(cudaStreamWaitEvent(s, e->ev_comm, 0));
         e->comm_pending = 0;
      }
      if (e->p2p) {
         pf::k_p2p_wait<<<1, 1, 0, s>>>(e->flags, e->ipc_flags_lo != nullptr, e->ipc_flags_hi != nullptr, e->flags + 2);
         e->launches += 1;
      }
      // (single process: the neighbours pushed them; their events were recorded when the host queued their previous step)
      if (e->peer_lo && n > e->first_step) CU(cudaStreamWaitEvent(s, e->peer_lo->ev_push, 0));
      if (e->peer_hi && n > e->first_step) CU(cudaStreamWaitEvent(s, e->peer_hi->ev_push, 0));
      if (!fused) {
         // 1. previous-state values at the ABC nodes; 2.+3. seam row and halo mirrors of u1
         if (e->Nba) {
            pf::k_gather<Real><<<nblk(e->Nba, 128), 128, 0, s>>>(u0, e->bna, (Real *)e->u2ba, e->Nba);
            e->launches += 1;
         }
         // (the z halos are in place when the previous step's k_abc wrote them: the seam copy and the z mirror commute, and the y / x
         // mirrors, which copy whole rows, still come after both)
         mirror_pass<Real>(e, u1, s, zflip_on(e) && e->zhalo_ok);
         e->halo_dirty = 1;  // the state this step produces has no mirrored halos yet
      } else if (e->halo_dirty) {
         mirror_pass<Real>(e, u1, s);
         e->halo_dirty = 0;
      }
      if (e->energy_on && (rc = energy_pre<Real>(e, u1, u0, n, s))) return rc;
   }
   if (split) {
      if (phases & PH_A) {
         // planes the neighbours need first.  A slab with two neighbours runs its two edge planes side by side on two streams:
         // each is a chain of five short kernels (a one-plane air launch, shell, rigid, branches, io) that leaves the GPU mostly idle
         const bool both = lo && hi && e->edge_overlap && !e->fd_bulk && !e->profile_air;
         Step<Real> st2 = st;
         if (both) {
            CU(cudaEventRecord(e->ev_e0, s));
            CU(cudaStreamWaitEvent(e->s_edge, e->ev_e0, 0));
            st2.s = e->s_edge, st2.ctr = e->tma.ctr + 2;
         }
         if (lo && (rc = st.part(Part{1, 2, 0, nb_lo, 0, e->nbl_lo, 0, e->nba_lo, 0, e->ns_lo, 0, e->np_lo, false}))) return rc;
         if (hi && (rc = st2.part(Part{Nx - 2, Nx - 1, NB - nb_hi, nb_hi, e->Nbl - e->nbl_hi, e->nbl_hi, e->Nba - e->nba_hi,
                                       e->nba_hi, e->Ns - e->ns_hi, e->ns_hi, e->np - e->np_hi, e->np_hi, false})))
            return rc;
         if (both) {
            CU(cudaEventRecord(e->ev_e1, e->s_edge));
            CU(cudaStreamWaitEvent(s, e->ev_e1, 0));
         }
         CU(cudaGetLastError());
      }
      if (phases & PH_X) {
         // then the exchange on the comm stream while the interior runs
         CU(cudaEventRecord(e->ev_edge, s));
         CU(cudaStreamWaitEvent(e->s_comm, e->ev_edge, 0));
         if ((rc = exchange(e, u0, e->s_comm))) return rc;
         CU(cudaEventRecord(e->ev_comm, e->s_comm));
         if (e->peer_lo || e->peer_hi) CU(cudaEventRecord(e->ev_push, e->s_comm));
         e->comm_pending = 1;
      }
      if (phases & PH_B) {
         const i64 b0 = lo ? nb_lo : 0, b1 = hi ? nb_hi : 0, l0 = lo ? e->nbl_lo : 0, l1 = hi ? e->nbl_hi : 0;
         const i64 a0 = lo ? e->nba_lo : 0, a1 = hi ? e->nba_hi : 0, s0 = lo ? e->ns_lo : 0, s1 = hi ? e->ns_hi : 0;
         const i64 p0 = lo ? e->np_lo : 0, p1 = hi ? e->np_hi : 0;
         if ((rc = st.part(Part{lo ? 2 : 1, hi ? Nx - 2 : Nx - 1, b0, NB - b0 - b1, l0, e->Nbl - l0 - l1, a0, e->Nba - a0 - a1, s0,
                                e->Ns - s0 - s1, p0, e->np - p0 - p1, true})))
            return rc;
         CU(cudaGetLastError());
      }
   } else {
      if (phases & PH_B) {
         if ((rc = st.part(Part{1, Nx - 1, 0, NB, 0, e->Nbl, 0, e->Nba, 0, e->Ns, 0, e->np, true}))) return rc;
         CU(cudaGetLastError());
      }
      if (phases & PH_X) {
         if ((rc = exchange(e, u0, s))) return rc;
         if (e->peer_lo || e->peer_hi) CU(cudaEventRecord(e->ev_push, s));
      }
   }
   if (phases & PH_B) {
      if (e->energy_on && (rc = energy_post<Real>(e, u0, n, s))) return rc;
      // 10. advance the device step counter (unless the step's last k_io did), swap (the boundary history rotates with the
      // counter's parity)
      if (!e->ticked) {
         pf::k_tick<<<1, 1, 0, s>>>(e->d_n);
         e->launches += 1;
      }
      e->ticked = 0;
      e->zhalo_ok = (!fused && zflip_on(e)) ? 1 : 0;
      e->n_dev = n + 1;
      e->cur ^= 1;
      e->steps_done = n + 1;
   }
   return PFFDTD_OK;
}

static int step_any(pffdtd_engine *e, i64 n, int phases = PH_ALL) {
   return e->precision == 1 ? step_impl<float>(e, n, phases) : step_impl<double>(e, n, phases);
}

// can a step starting now be replayed from a captured graph?
static bool graphable(const pffdtd_engine *e) {
   return e->use_graph && !e->energy_on && !e->profile_air && !e->manual_halo && !e->peer_lo && !e->peer_hi && e->steps_plain >= 2 &&
          !(e->halo_dirty && step_fused(e)) && !(zflip_on(e) && !step_fused(e) && !e->zhalo_ok);
}

// Host-side bookkeeping a step changes; capturing a step must leave it as it was (nothing ran).
struct HostState {
   int cur, halo_dirty, comm_pending, abc_pending, host_mode, zhalo_ok;
   i64 n_dev, steps_done;
   double launches;
};
static HostState save_state(const pffdtd_engine *e) {
   return HostState{e->cur, e->halo_dirty, e->comm_pending, e->abc_pending, e->host_mode, e->zhalo_ok, e->n_dev, e->steps_done, e->launches};
}
static void restore_state(pffdtd_engine *e, const HostState &h) {
   e->cur = h.cur, e->halo_dirty = h.halo_dirty, e->comm_pending = h.comm_pending, e->abc_pending = h.abc_pending;
   e->host_mode = h.host_mode, e->zhalo_ok = h.zhalo_ok, e->n_dev = h.n_dev, e->steps_done = h.steps_done, e->launches = h.launches;
}

// a halo exchange still in flight on the comm stream is joined for real (never from inside a capture)
static int join_comm(pffdtd_engine *e) {
   if (e->comm_pending) {
      CU(cudaStreamWaitEvent(e->s_main, e->ev_comm, 0));
      e->comm_pending = 0;
   }
   return 0;
}

// Capture `count` consecutive steps (or the given phases of one step) that start at time index n with grid role `c` (u[c] =
// current state) into an executable graph; `host_io`: the sequence is one host-driven step with the H2D / D2H copies of its
// samples.  Nothing runs and no host-side state changes, whatever the outcome.
static int capture_steps(pffdtd_engine *e, int c, i64 n, int count, bool host_io, cudaGraphExec_t *out, double *launches,
                         int phases = PH_ALL) {
   const HostState keep = save_state(e);
   cudaGraph_t g = nullptr;
   e->cur = c, e->n_dev = n, e->comm_pending = 0, e->abc_pending = 0;
   cudaError_t ce = cudaStreamBeginCapture(e->s_main, cudaStreamCaptureModeThreadLocal);
   if (ce != cudaSuccess) {
      restore_state(e, keep);
      return fail(PFFDTD_ECUDA, "graph capture: %s", cudaGetErrorString(ce));
   }
   int rc = PFFDTD_OK;
   cudaError_t cc = cudaSuccess;
   if (host_io) {
      e->host_mode = 3;
      cc = cudaMemcpyAsync(e->in_stage, e->h_in, (size_t)e->Ns * e->rs, cudaMemcpyHostToDevice, e->s_main);
   }
   for (int k = 0; k < count && rc == PFFDTD_OK; k++) rc = step_any(e, n + k, phases);
   if (host_io && cc == cudaSuccess)
      cc = cudaMemcpyAsync(e->h_out, e->out_stage, (size_t)e->Nr * e->rs, cudaMemcpyDeviceToHost, e->s_main);
   if (e->comm_pending && cc == cudaSuccess) cc = cudaStreamWaitEvent(e->s_main, e->ev_comm, 0);  // re-join the comm stream
   ce = cudaStreamEndCapture(e->s_main, &g);
   *launches = e->launches - keep.launches;
   restore_state(e, keep);
   if (rc || cc != cudaSuccess || ce != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      return rc ? rc : fail(PFFDTD_ECUDA, "graph capture: %s", cudaGetErrorString(cc != cudaSuccess ? cc : ce));
   }
   ce = cudaGraphInstantiate(out, g, 0);
   cudaGraphDestroy(g);
   if (ce != cudaSuccess) {
      *out = nullptr;
      return fail(PFFDTD_ECUDA, "graph instantiate: %s", cudaGetErrorString(ce));
   }
   return PFFDTD_OK;
}

extern "C" int pffdtd_run_steps(pffdtd_engine *e, int64_t nstart, int64_t nsteps) {
   if (!e) return fail(PFFDTD_EINVAL, "NULL engine");
   if (nsteps < 0 || nstart < 0 || nstart + nsteps > e->Nt)
      return fail(PFFDTD_EINVAL, "steps [%lld,%lld) outside [0,%lld)", (long long)nstart, (long long)(nstart + nsteps), (long long)e->Nt);
   if ((!e->x_lo_edge || !e->x_hi_edge) && !e->comm && !e->p2p && !e->manual_halo)
      return fail(PFFDTD_ESTATE, e->peer_lo || e->peer_hi ? "slab of a pffdtd_multi: step it with pffdtd_multi_run_steps"
                                                         : "slab engine without communicator: call pffdtd_comm_init");
   CU(cudaSetDevice(e->device));
   i64 n = nstart;
   const i64 nend = nstart + nsteps;
   while (n < nend) {
      // two steps bring `cur` back: a captured pair replays as one CUDA graph (fused or not, one GPU or a slab with its
      // halo exchange), once the halos are clean and the first plain steps have sized the launches and opened the
      // NCCL connections.  Both grid roles are captured at the first use, so no later call pays for a capture.
      if (e->comm && !e->p2p && graphable(e)) {
         // a slab with an NCCL communicator: the work before and after the exchange replays from two graphs per grid role, the
         // exchange (event, ncclSend/ncclRecv on the comm stream, event) is issued eagerly between them
         const int c = e->cur;
         const bool split = step_split(e);
         int rc = join_comm(e);
         if (rc) return rc;
         for (int k = 0; k < 2; k++) {
            const int cc = c ^ k;
            if (split && !e->graphA[cc] && (rc = capture_steps(e, cc, n, 1, false, &e->graphA[cc], &e->graphA_launches[cc], PH_A))) return rc;
            if (!e->graphB[cc] && (rc = capture_steps(e, cc, n, 1, false, &e->graphB[cc], &e->graphB_launches[cc], PH_B))) return rc;
         }
         if (e->n_dev != n) {
            pf::k_set_n<<<1, 1, 0, e->s_main>>>(e->d_n, n);
            e->n_dev = n;
         }
         if (split) {
            CU(cudaGraphLaunch(e->graphA[c], e->s_main));
            e->launches += e->graphA_launches[c];
            if ((rc = step_any(e, n, PH_X))) return rc;
         }
         CU(cudaGraphLaunch(e->graphB[c], e->s_main));
         e->launches += e->graphB_launches[c];
         if (!split && (rc = step_any(e, n, PH_X))) return rc;
         e->cur ^= 1;
         n += 1;
         e->n_dev = n;
         e->steps_done = n;
         continue;
      }
      const bool graph_ok = nend - n >= 2 && graphable(e) && (!e->comm || e->p2p);
      if (graph_ok) {
         const int c = e->cur;
         int rc = join_comm(e);
         if (rc) return rc;
         for (int k = 0; k < 2; k++) {
            const int cc = c ^ k;
            if (!e->graph[cc] && (rc = capture_steps(e, cc, n, 2, false, &e->graph[cc], &e->graph_launches[cc]))) return rc;
         }
         if (e->n_dev != n) {
            pf::k_set_n<<<1, 1, 0, e->s_main>>>(e->d_n, n);
            e->n_dev = n;
         }
         CU(cudaGraphLaunch(e->graph[c], e->s_main));
         e->launches += e->graph_launches[c];
         n += 2;
         e->n_dev = n;
         e->steps_done = n;
         continue;
      }
      int rc = step_any(e, n);
      if (rc) return rc;
      e->steps_plain++;
      n++;
   }
   return PFFDTD_OK;
}

extern "C" int pffdtd_sync(pffdtd_engine *e) {
   if (!e) return fail(PFFDTD_EINVAL, "NULL engine");
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->s_main));
   CU(cudaStreamSynchronize(e->s_comm));
   if (e->p2p && e->flags) {
      long long err = 0;
      CU(cudaMemcpy(&err, e->flags + 3, 8, cudaMemcpyDeviceToHost));
      if (err) return fail(PFFDTD_ESTATE, "halo exchange over peer memory: a neighbour's plane never arrived (the wait gave up)");
   }
   return PFFDTD_OK;
}

extern "C" int pffdtd_step_host(pffdtd_engine *e, int64_t n, const double *in_samples, double *out_samples) {
   if (!e) return fail(PFFDTD_EINVAL, "NULL engine");
   if (n < 0 || n >= e->Nt) return fail(PFFDTD_EINVAL, "step %lld outside [0,%lld)", (long long)n, (long long)e->Nt);
   if ((!e->x_lo_edge || !e->x_hi_edge) && !e->comm && !e->p2p && !e->manual_halo)
      return fail(PFFDTD_ESTATE, "slab engine without communicator: call pffdtd_comm_init");
   CU(cudaSetDevice(e->device));
   const bool has_in = in_samples && e->Ns, has_out = out_samples && e->Nr;
   if (has_in) {
      if (e->precision == 1) for (i64 s = 0; s < e->Ns; s++) ((float *)e->h_in)[s] = (float)in_samples[s];
      else memcpy(e->h_in, in_samples, (size_t)e->Ns * 8);
   }
   int rc = PFFDTD_OK;
   if (has_in && has_out && graphable(e) && (!e->comm || e->p2p)) {
      // H2D of the source samples, the step, D2H of the receiver samples: one replayed graph per grid role
      const int c = e->cur;
      if ((rc = join_comm(e))) return rc;
      for (int k = 0; k < 2; k++) {
         const int cc = c ^ k;
         if (!e->hgraph[cc] && (rc = capture_steps(e, cc, n, 1, true, &e->hgraph[cc], &e->hgraph_launches[cc]))) return rc;
      }
      if (e->n_dev != n) {
         pf::k_set_n<<<1, 1, 0, e->s_main>>>(e->d_n, n);
         e->n_dev = n;
      }
      CU(cudaGraphLaunch(e->hgraph[c], e->s_main));
      e->launches += e->hgraph_launches[c];
      e->cur ^= 1;
      e->n_dev = n + 1;
      e->steps_done = n + 1;
   } else {
      e->host_mode = (has_in ? 1 : 0) | (has_out ? 2 : 0);
      cudaError_t ce = cudaSuccess;
      if (has_in) ce = cudaMemcpyAsync(e->in_stage, e->h_in, (size_t)e->Ns * e->rs, cudaMemcpyHostToDevice, e->s_main);
      rc = step_any(e, n);
      e->host_mode = 0;
      if (rc) return rc;
      if (ce != cudaSuccess) return fail(PFFDTD_ECUDA, "source sample upload: %s", cudaGetErrorString(ce));
      e->steps_plain++;
      if (has_out) CU(cudaMemcpyAsync(e->h_out, e->out_stage, (size_t)e->Nr * e->rs, cudaMemcpyDeviceToHost, e->s_main));
   }
   CU(cudaStreamSynchronize(e->s_main));
   if (has_out) {
      if (e->precision == 1) for (i64 r = 0; r < e->Nr; r++) out_samples[r] = (double)((float *)e->h_out)[r];
      else memcpy(out_samples, e->h_out, (size_t)e->Nr * 8);
   }
   return PFFDTD_OK;
}

extern "C" int pffdtd_read_outputs(pffdtd_engine *e, int64_t n0, int64_t n1, double *u_out) {
   if (!e || !u_out) return fail(PFFDTD_EINVAL, "NULL argument");
   if (n0 < 0 || n1 < n0 || n1 > e->Nt) return fail(PFFDTD_EINVAL, "bad step range");
   CU(cudaSetDevice(e->device));
   CU(cudaStreamSynchronize(e->s_main));
   const i64 W = n1 - n0;
   if (W == 0 || e->Nr == 0) return PFFDTD_OK;
   std::vector<char> tmp((size_t)(W * e->Nr) * e->rs);
   CU(cudaMemcpy(tmp.data(), (char *)e->uout + (size_t)(n0 * e->Nr) * e->rs, tmp.size(), cudaMemcpyDeviceToHost));
   for (i64 k = 0; k < W; k++)
      for (i64 r = 0; r < e->Nr; r++)
         u_out[r * W + k] = e->precision == 1 ? (double)((float *)tmp.data())[k * e->Nr + r] : ((double *)tmp.data())[k * e->Nr + r];
   return PFFDTD_OK;
}

template <typename Real>
static int grid_io(pffdtd_engine *e, int which, double *host, bool to_host) {
   Real *g = (Real *)e->u[which ? e->cur : e->cur ^ 1];
   const i64 nrows = e->Nx * e->Ny;
   // through a device staging buffer of one x-plane at a time in the reference layout
   const i64 rows_per = e->Ny;
   double *stage = nullptr;
   CU(cudaMalloc((void **)&stage, (size_t)(rows_per * e->Nz) * 8));
   int rc = PFFDTD_OK;
   for (i64 r0 = 0; r0 < nrows && rc == PFFDTD_OK; r0 += rows_per) {
      dim3 grd(nblk(e->Nz, 128), (unsigned)rows_per, 1);
      cudaError_t ce;
      if (to_host) {
         pf::k_unpad<Real><<<grd, 128, 0, e->s_main>>>(g + r0 * e->Nzp, stage, rows_per, e->Nz, e->Nzp);
         ce = cudaMemcpyAsync(host + r0 * e->Nz, stage, (size_t)(rows_per * e->Nz) * 8, cudaMemcpyDeviceToHost, e->s_main);
      } else {
         ce = cudaMemcpyAsync(stage, host + r0 * e->Nz, (size_t)(rows_per * e->Nz) * 8, cudaMemcpyHostToDevice, e->s_main);
         pf::k_pad<Real><<<grd, 128, 0, e->s_main>>>(g + r0 * e->Nzp, stage, rows_per, e->Nz, e->Nzp);
      }
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->s_main);
      if (ce != cudaSuccess) rc = fail(PFFDTD_ECUDA, "grid transfer: %s", cudaGetErrorString(ce));
   }
   cudaFree(stage);
   return rc;
}

extern "C" int pffdtd_read_grid(pffdtd_engine *e, int which, double *out) {
   if (!e || !out) return fail(PFFDTD_EINVAL, "NULL argument");
   CU(cudaSetDevice(e->device));
   int rc = pffdtd_sync(e);
   if (rc) return rc;
   return e->precision == 1 ? grid_io<float>(e, which, out, true) : grid_io<double>(e, which, out, true);
}

extern "C" int pffdtd_write_grid(pffdtd_engine *e, int which, const double *in) {
   if (!e || !in) return fail(PFFDTD_EINVAL, "NULL argument");
   CU(cudaSetDevice(e->device));
   int rc = pffdtd_sync(e);
   if (rc) return rc;
   e->halo_dirty = 1;
   e->zhalo_ok = 0;
   return e->precision == 1 ? grid_io<float>(e, which, (double *)in, false) : grid_io<double>(e, which, (double *)in, false);
}

extern "C" int pffdtd_read_boundary_state(pffdtd_engine *e, double *vh1, double *gh1) {
   if (!e || !vh1 || !gh1) return fail(PFFDTD_EINVAL, "NULL argument");
   CU(cudaSetDevice(e->device));
   int rc = pffdtd_sync(e);
   if (rc) return rc;
   const size_t n = (size_t)e->Nblp * PFFDTD_MMB * 2;
   if (e->Nbl == 0) return PFFDTD_OK;
   std::vector<char> tv(n * e->rs);
   CU(cudaMemcpy(tv.data(), e->vh1, tv.size(), cudaMemcpyDeviceToHost));
   for (i64 i = 0; i < e->Nbl; i++)
      for (int m = 0; m < PFFDTD_MMB; m++) {
         const size_t src = (size_t)pf::st_idx<PFFDTD_MMB>(m, i), dst = (size_t)i * PFFDTD_MMB + m;
         vh1[dst] = e->precision == 1 ? (double)((float *)tv.data())[src] : ((double *)tv.data())[src];
         gh1[dst] = e->precision == 1 ? (double)((float *)tv.data())[src + 32] : ((double *)tv.data())[src + 32];
      }
   return PFFDTD_OK;
}

// device self-test of the constant-divisor division used by the fused absorbing shell: compares
// div_by_const with __ddiv_rn on `count` pseudo-random numerators for divisor 1 + l*Q; *mismatches out
extern "C" int pffdtd_selftest(int device, double l, int precision, int64_t count, int64_t *mismatches) {
   if (!mismatches || count < 0) return fail(PFFDTD_EINVAL, "bad selftest arguments");
   CU(cudaSetDevice(device));
   unsigned long long *bad = nullptr, h = 0;
   CU(cudaMalloc((void **)&bad, 8));
   CU(cudaMemset(bad, 0, 8));
   const int per_thread = 256, threads = 256;
   const unsigned blocks = (unsigned)std::max<int64_t>(1, count / (3LL * per_thread * threads));
   for (int Q = 1; Q <= 3; Q++) {
      const double lQ = precision == 1 ? (double)((float)l * (float)Q) : l * (double)Q;
      const double b = 1.0 + lQ;
      pf::k_selftest_div<<<blocks, threads>>>(b, 1.0 / b, 0x1234567ull * Q, per_thread, precision == 1, bad);
   }
   CU(cudaGetLastError());
   CU(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
   cudaFree(bad);
   *mismatches = (int64_t)h;
   return PFFDTD_OK;
}

extern "C" int64_t pffdtd_air_chunk_plan(int64_t n_planes, int xc, int64_t *bounds, int64_t max_bounds) {
   if (n_planes < 0 || n_planes > 0x7fffffff || !bounds) return fail(PFFDTD_EINVAL, "bad chunk plan arguments");
   pf::AirJob jb;
   pf::air_plan_chunks(&jb, (int)n_planes, xc);
   if ((i64)jb.nch + 1 > max_bounds) return fail(PFFDTD_EINVAL, "chunk plan needs %d bounds", jb.nch + 1);
   for (int k = 0; k <= jb.nch; k++)
      bounds[k] = k < jb.n_main ? (i64)k * jb.xc : (i64)jb.tail[k - jb.n_main];
   return jb.nch;
}

extern "C" int pffdtd_run_sim(const pffdtd_desc *desc, int device, double *u_out, double *elapsed_s) {
   if (!desc || !u_out) return fail(PFFDTD_EINVAL, "NULL argument");
   pffdtd_engine *e = nullptr;
   int rc = pffdtd_create(desc, device, &e);
   if (rc) return rc;
   auto t0 = std::chrono::steady_clock::now();
   rc = pffdtd_run_steps(e, 0, desc->Nt);
   if (rc == PFFDTD_OK) rc = pffdtd_sync(e);
   auto t1 = std::chrono::steady_clock::now();
   if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
   if (rc == PFFDTD_OK) rc = pffdtd_read_outputs(e, 0, desc->Nt, u_out);
   std::string keep = g_err;
   pffdtd_destroy(e);
   g_err = keep;
   return rc;
}

// ------------------------------------------------------------------------------------------------
// single-process multi-GPU: one host thread drives every slab, as the reference's run_sim does
// (gpu_engine.h:679-691 device count, :516-662 split_data, :994 the per-step loop over devices, :1086-1126 the exchange)
// ------------------------------------------------------------------------------------------------
struct SlabData {
   pffdtd_desc d{};
   std::vector<int64_t> bn, bnl, bna, in, out;
   std::vector<uint16_t> adj;
   std::vector<int8_t> mat, Q;
   std::vector<double> ssaf, insig;
};
struct pffdtd_multi {
   std::vector<pffdtd_engine *> eng;
   std::vector<i64> x0, nx;  // first owned plane and owned planes per slab
   i64 Nr = 0, Nt = 0;
};

// per x-plane cost, the C++ twin of SimData.plane_costs (pffdtd_b200/sim_data.py): air nodes + weighted list entries
static std::vector<double> plane_costs(const pffdtd_desc *d) {
   const double COST_BN = 3.0, COST_BNL_BASE = 4.0, COST_BNL_BRANCH = 1.5, COST_BNA = 2.0;
   const i64 P = d->Ny * d->Nz;
   std::vector<double> c((size_t)d->Nx, (double)P);
   if (d->x_lo_edge) c[0] = 0.0;
   if (d->x_hi_edge) c[(size_t)d->Nx - 1] = 0.0;
   int mb = 0;
   for (int k = 0; k < d->Nm; k++) mb = std::max<int>(mb, d->Mb[k]);
   for (i64 i = 0; i < d->Nb; i++) c[(size_t)(d->bn_ixyz[i] / P)] += COST_BN;
   for (i64 i = 0; i < d->Nbl; i++) c[(size_t)(d->bnl_ixyz[i] / P)] += COST_BNL_BASE + COST_BNL_BRANCH * mb;
   for (i64 i = 0; i < d->Nba; i++) c[(size_t)(d->bna_ixyz[i] / P)] += COST_BNA;
   return c;
}

// owned planes per slab: the reference's equal split (gpu_engine.h:532-543) or, with `cost`, slabs of about equal cost
// (the twin of SimData.slab_planes)
static int slab_planes(i64 Nx, int n, const std::vector<double> *cost, std::vector<i64> *starts, std::vector<i64> *sizes) {
   sizes->assign((size_t)n, 0);
   if (!cost) {
      for (int r = 0; r < n; r++) (*sizes)[(size_t)r] = Nx / n + (r < Nx % n ? 1 : 0);
   } else {
      if (Nx < 2 * (i64)n) return fail(PFFDTD_EINVAL, "too many slabs for this grid");
      std::vector<double> cum((size_t)Nx + 1, 0.0);
      for (i64 x = 0; x < Nx; x++) cum[(size_t)x + 1] = cum[(size_t)x] + (*cost)[(size_t)x];
      std::vector<i64> cuts{0};
      for (int r = 1; r < n; r++) {
         const double target = cum[(size_t)Nx] * r / n;
         i64 x = std::lower_bound(cum.begin(), cum.end(), target) - cum.begin();
         if (x > 0 && std::fabs(cum[(size_t)x - 1] - target) <= std::fabs(cum[(size_t)x] - target)) x--;
         x = std::min(std::max(x, cuts.back() + 2), Nx - 2 * (i64)(n - r));
         cuts.push_back(x);
      }
      cuts.push_back(Nx);
      for (int r = 0; r < n; r++) (*sizes)[(size_t)r] = cuts[(size_t)r + 1] - cuts[(size_t)r];
   }
   starts->assign((size_t)n, 0);
   for (int r = 1; r < n; r++) (*starts)[(size_t)r] = (*starts)[(size_t)r - 1] + (*sizes)[(size_t)r - 1];
   for (int r = 0; r < n; r++)
      if ((*sizes)[(size_t)r] < 2) return fail(PFFDTD_EINVAL, "too many slabs for this grid");
   return 0;
}

// the part of `d` slab r owns, re-based to slab-local indices, one halo plane towards each neighbour (gpu_engine.h:755-823;
// the twin of SimData.slab).  The node lists must be sorted (check_sorted, gpu_engine.h:497-513).
static void slab_desc(const pffdtd_desc *d, i64 start, i64 size, bool first, bool last, SlabData *s) {
   const i64 P = d->Ny * d->Nz, lo = start * P, hi = (start + size) * P;
   const i64 first_plane = start - (first ? 0 : 1), off = first_plane * P;
   auto own = [&](const int64_t *a, i64 n, i64 *b, i64 *e) {
      *b = std::lower_bound(a, a + n, lo) - a;
      *e = std::lower_bound(a, a + n, hi) - a;
   };
   auto rebase = [&](const int64_t *a, i64 b, i64 e, std::vector<int64_t> *out) {
      out->resize((size_t)(e - b));
      for (i64 i = b; i < e; i++) (*out)[(size_t)(i - b)] = a[i] - off;
   };
   i64 b, e;
   s->d = *d;
   s->d.Nx = size + (first ? 0 : 1) + (last ? 0 : 1);
   s->d.ix0 = d->ix0 + first_plane;
   s->d.x_lo_edge = d->x_lo_edge && first, s->d.x_hi_edge = d->x_hi_edge && last;
   own(d->bn_ixyz, d->Nb, &b, &e);
   rebase(d->bn_ixyz, b, e, &s->bn);
   s->adj.assign(d->adj_bn + b, d->adj_bn + e);
   s->d.Nb = e - b;
   own(d->bnl_ixyz, d->Nbl, &b, &e);
   rebase(d->bnl_ixyz, b, e, &s->bnl);
   s->mat.assign(d->mat_bnl + b, d->mat_bnl + e);
   s->ssaf.assign(d->ssaf_bnl + b, d->ssaf_bnl + e);
   s->d.Nbl = e - b;
   own(d->bna_ixyz, d->Nba, &b, &e);
   rebase(d->bna_ixyz, b, e, &s->bna);
   s->Q.assign(d->Q_bna + b, d->Q_bna + e);
   s->d.Nba = e - b;
   own(d->in_ixyz, d->Ns, &b, &e);
   rebase(d->in_ixyz, b, e, &s->in);
   s->insig.assign(d->in_sigs + b * d->Nt, d->in_sigs + e * d->Nt);
   s->d.Ns = e - b;
   own(d->out_ixyz, d->Nr, &b, &e);
   rebase(d->out_ixyz, b, e, &s->out);
   s->d.Nr = e - b;
   s->d.bn_ixyz = s->bn.data(), s->d.adj_bn = s->adj.data(), s->d.bnl_ixyz = s->bnl.data(), s->d.mat_bnl = s->mat.data();
   s->d.ssaf_bnl = s->ssaf.data(), s->d.bna_ixyz = s->bna.data(), s->d.Q_bna = s->Q.data(), s->d.in_ixyz = s->in.data();
   s->d.out_ixyz = s->out.data(), s->d.in_sigs = s->insig.data();
}

// Diagnostic (no device needed): the slab plan pffdtd_multi_create would use for `desc`
extern "C" int pffdtd_slab_plan(const pffdtd_desc *desc, int nslabs, int balance, int64_t *starts, int64_t *sizes) {
   if (!desc || !starts || !sizes || nslabs < 1) return fail(PFFDTD_EINVAL, "bad slab plan arguments");
   std::vector<double> cost;
   if (balance) cost = plane_costs(desc);
   std::vector<i64> st, sz;
   int rc = slab_planes(desc->Nx, nslabs, balance ? &cost : nullptr, &st, &sz);
   if (rc) return rc;
   for (int r = 0; r < nslabs; r++) starts[r] = st[(size_t)r], sizes[r] = sz[(size_t)r];
   return PFFDTD_OK;
}

extern "C" int pffdtd_multi_destroy(pffdtd_multi *m) {
   if (!m) return PFFDTD_OK;
   for (pffdtd_engine *e : m->eng)
      if (e) {
         cudaSetDevice(e->device);
         cudaStreamSynchronize(e->s_main);
         cudaStreamSynchronize(e->s_comm);
      }
   for (pffdtd_engine *e : m->eng) pffdtd_destroy(e);
   delete m;
   return PFFDTD_OK;
}

extern "C" int pffdtd_multi_create(const pffdtd_desc *desc, int nslabs, const int *devices, int balance, pffdtd_multi **out) {
   if (!desc || !out) return fail(PFFDTD_EINVAL, "NULL argument");
   if (desc->struct_size != (int32_t)sizeof(pffdtd_desc)) return fail(PFFDTD_EINVAL, "pffdtd_desc size mismatch");
   int ndev = 0;
   CU(cudaGetDeviceCount(&ndev));
   if (ndev < 1) return fail(PFFDTD_ECUDA, "no CUDA device");
   if (nslabs <= 0) nslabs = ndev;  // every visible device, the reference's rule (CUDA_VISIBLE_DEVICES picks them)
   if (nslabs > 1 && (!ascending(desc->bn_ixyz, desc->Nb) || !ascending(desc->bnl_ixyz, desc->Nbl) || !ascending(desc->bna_ixyz, desc->Nba) ||
                      !ascending(desc->in_ixyz, desc->Ns, false) || !ascending(desc->out_ixyz, desc->Nr, false)))
      return fail(PFFDTD_EINVAL, "more than one slab needs sorted node lists (a sort_sim_data'd gpu folder, gpu_engine.h:497-513)");
   pffdtd_multi *m = new pffdtd_multi();
   m->Nr = desc->Nr, m->Nt = desc->Nt;
   int rc = PFFDTD_OK;
   if (nslabs == 1) {
      m->x0 = {0}, m->nx = {desc->Nx};
      pffdtd_engine *e = nullptr;
      rc = pffdtd_create(desc, devices ? devices[0] : 0, &e);
      m->eng.push_back(e);
   } else {
      std::vector<double> cost;
      if (balance) cost = plane_costs(desc);
      rc = slab_planes(desc->Nx, nslabs, balance ? &cost : nullptr, &m->x0, &m->nx);
      for (int r = 0; r < nslabs && rc == PFFDTD_OK; r++) {
         SlabData sd;
         slab_desc(desc, m->x0[(size_t)r], m->nx[(size_t)r], r == 0, r == nslabs - 1, &sd);
         pffdtd_engine *e = nullptr;
         rc = pffdtd_create(&sd.d, devices ? devices[r] : r % ndev, &e);
         m->eng.push_back(e);
      }
      for (int r = 0; r < nslabs && rc == PFFDTD_OK; r++) {
         pffdtd_engine *e = m->eng[(size_t)r];
         e->peer_lo = r > 0 ? m->eng[(size_t)r - 1] : nullptr;
         e->peer_hi = r < nslabs - 1 ? m->eng[(size_t)r + 1] : nullptr;
         for (pffdtd_engine *p : {e->peer_lo, e->peer_hi})
            if (p && p->device != e->device) {
               cudaSetDevice(e->device);
               int can = 0;
               cudaDeviceCanAccessPeer(&can, e->device, p->device);
               if (can && cudaDeviceEnablePeerAccess(p->device, 0) != cudaSuccess) cudaGetLastError();  // (already enabled is fine)
            }
      }
   }
   if (rc != PFFDTD_OK) {
      std::string keep = g_err;
      pffdtd_multi_destroy(m);
      g_err = keep;
      return rc;
   }
   *out = m;
   return PFFDTD_OK;
}

extern "C" int pffdtd_multi_slabs(pffdtd_multi *m, int64_t *planes, int max) {
   if (!m) return fail(PFFDTD_EINVAL, "NULL argument");
   for (size_t r = 0; r < m->eng.size() && planes && (int)r < max; r++) planes[r] = m->nx[r];
   return (int)m->eng.size();
}

extern "C" int pffdtd_multi_engine(pffdtd_multi *m, int slab, pffdtd_engine **e) {
   if (!m || !e || slab < 0 || slab >= (int)m->eng.size()) return fail(PFFDTD_EINVAL, "no such slab");
   *e = m->eng[(size_t)slab];
   return PFFDTD_OK;
}

extern "C" int pffdtd_multi_run_steps(pffdtd_multi *m, int64_t nstart, int64_t nsteps) {
   if (!m) return fail(PFFDTD_EINVAL, "NULL argument");
   if (m->eng.size() == 1) return pffdtd_run_steps(m->eng[0], nstart, nsteps);
   if (nsteps < 0 || nstart < 0 || nstart + nsteps > m->Nt)
      return fail(PFFDTD_EINVAL, "steps [%lld,%lld) outside [0,%lld)", (long long)nstart, (long long)(nstart + nsteps), (long long)m->Nt);
   // all slabs step in lockstep, queued by this one thread; nothing here waits for the devices
   for (pffdtd_engine *e : m->eng) e->first_step = nstart == 0 ? 0 : -1;  // later batches: the neighbours' events of the previous batch hold
   for (i64 n = nstart; n < nstart + nsteps; n++)
      for (pffdtd_engine *e : m->eng) {
         CU(cudaSetDevice(e->device));
         int rc = step_any(e, n);
         if (rc) return rc;
         e->steps_plain++;
      }
   return PFFDTD_OK;
}

extern "C" int pffdtd_multi_sync(pffdtd_multi *m) {
   if (!m) return fail(PFFDTD_EINVAL, "NULL argument");
   for (pffdtd_engine *e : m->eng) {
      int rc = pffdtd_sync(e);
      if (rc) return rc;
   }
   return PFFDTD_OK;
}

// rows in slab order == the sorted receiver order of the whole grid (slabs are contiguous in x, lists sorted)
extern "C" int pffdtd_multi_read_outputs(pffdtd_multi *m, int64_t n0, int64_t n1, double *u_out) {
   if (!m || !u_out) return fail(PFFDTD_EINVAL, "NULL argument");
   int rc = pffdtd_multi_sync(m);
   i64 row = 0;
   for (size_t r = 0; r < m->eng.size() && rc == PFFDTD_OK; r++) {
      rc = pffdtd_read_outputs(m->eng[r], n0, n1, u_out + row * (n1 - n0));
      row += m->eng[r]->Nr;
   }
   return rc;
}

extern "C" int pffdtd_run_sim_multi(const pffdtd_desc *desc, int nslabs, const int *devices, double *u_out, double *elapsed_s) {
   if (!desc || !u_out) return fail(PFFDTD_EINVAL, "NULL argument");
   pffdtd_multi *m = nullptr;
   int rc = pffdtd_multi_create(desc, nslabs, devices, 1, &m);
   if (rc) return rc;
   auto t0 = std::chrono::steady_clock::now();
   rc = pffdtd_multi_run_steps(m, 0, desc->Nt);
   if (rc == PFFDTD_OK) rc = pffdtd_multi_sync(m);
   auto t1 = std::chrono::steady_clock::now();
   if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
   if (rc == PFFDTD_OK) rc = pffdtd_multi_read_outputs(m, 0, desc->Nt, u_out);
   std::string keep = g_err;
   pffdtd_multi_destroy(m);
   g_err = keep;
   return rc;
}

// ------------------------------------------------------------------------------------------------
// the voxeliser's ray casting (VoxScene.calc_adj) -- see vox.cuh
// ------------------------------------------------------------------------------------------------
struct pffdtd_vox {
   int NN = 0;
   std::vector<int64_t> bn;
   std::vector<uint8_t> adj;
   std::vector<int32_t> tidx;
   std::vector<double> ndist;
};

extern "C" int pffdtd_vox_run(const pffdtd_vox_desc *d, int device, pffdtd_vox **out) {
   if (!d || !out) return fail(PFFDTD_EINVAL, "NULL argument");
   if (d->struct_size != (int32_t)sizeof(pffdtd_vox_desc)) return fail(PFFDTD_EINVAL, "pffdtd_vox_desc size mismatch");
   if ((d->NN != 6 && d->NN != 12) || d->Nvox < 0 || d->Ntris < 0) return fail(PFFDTD_EINVAL, "bad voxeliser description");
   int ndev = 0;
   CU(cudaGetDeviceCount(&ndev));
   if (device < 0 || device >= ndev) return fail(PFFDTD_ECUDA, "no CUDA device %d (%d visible)", device, ndev);
   CU(cudaSetDevice(device));
   // PFFDTD_VOX_TIMING=1: wall time of the phases on stderr (diagnostic; adds device synchronisations)
   const bool timing = getenv("PFFDTD_VOX_TIMING") != nullptr;
   auto t_last = std::chrono::steady_clock::now();
   auto lap = [&](const char *what) {
      if (!timing) return;
      cudaDeviceSynchronize();
      const auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[pffdtd_vox_run] %-34s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
   };
   std::vector<void *> dev;
   auto freeall = [&]() {
      for (void *p : dev) cudaFree(p);
   };
   auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
      void *p = nullptr;
      if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return 1;
      dev.push_back(p);
      if (bytes && src && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
      *dst = p;
      return 0;
   };
   // points of every voxel (halo layer included)
   std::vector<long long> pt_off((size_t)d->Nvox + 1, 0);
   for (i64 v = 0; v < d->Nvox; v++) {
      const int64_t *sh = d->vox_shape + 3 * v, *st = d->vox_start + 3 * v;
      if (sh[0] < 3 || sh[1] < 3 || sh[2] < 3 || st[0] < 0 || st[1] < 0 || st[2] < 0 || st[0] + sh[0] > d->Nx || st[1] + sh[1] > d->Ny ||
          st[2] + sh[2] > d->Nz || sh[0] * sh[1] * sh[2] > 0x7fffffff)
         return fail(PFFDTD_EINVAL, "voxel %lld outside the grid", (long long)v);
      pt_off[(size_t)v + 1] = pt_off[(size_t)v] + sh[0] * sh[1] * sh[2];
   }
   const size_t npt = (size_t)pt_off[(size_t)d->Nvox];
   const i64 ntl = d->Nvox ? d->vox_tri_off[d->Nvox] : 0;
   pf::VoxArgs a;
   memset(&a, 0, sizeof a);
   a.c = VoxConst{d->hf, d->c_bb, d->c_near, d->c_far, d->d_eps, d->cp_eps, -2.220446049250313e-16};
   a.NN = d->NN, a.fcc = d->fcc, a.Ny = d->Ny, a.Nz = d->Nz;
   int bad = 0;
   bad |= up(d->xv, (size_t)d->Nx * 8, (const void **)&a.xv) | up(d->yv, (size_t)d->Ny * 8, (const void **)&a.yv) | up(d->zv, (size_t)d->Nz * 8, (const void **)&a.zv);
   bad |= up(d->vvh, (size_t)d->NN * 24, (const void **)&a.vvh) | up(d->ray_un, (size_t)d->NN * 24, (const void **)&a.ray_un);
   bad |= up(d->vox_start, (size_t)d->Nvox * 24, (const void **)&a.vox_start) | up(d->vox_shape, (size_t)d->Nvox * 24, (const void **)&a.vox_shape);
   bad |= up(d->vox_tri_off, ((size_t)d->Nvox + 1) * 8, (const void **)&a.vox_tri_off) | up(pt_off.data(), pt_off.size() * 8, (const void **)&a.pt_off);
   bad |= up(d->vox_tri, (size_t)ntl * 4, (const void **)&a.vox_tri);
   const size_t t3 = (size_t)d->Ntris * 24;
   bad |= up(d->unor, t3, (const void **)&a.unor) | up(d->cent, t3, (const void **)&a.cent) | up(d->bmin, t3, (const void **)&a.bmin) | up(d->bmax, t3, (const void **)&a.bmax);
   bad |= up(d->v, t3 * 3, (const void **)&a.v) | up(d->eab, t3, (const void **)&a.eab) | up(d->ebc, t3, (const void **)&a.ebc) | up(d->eca, t3, (const void **)&a.eca);
   const void *p_nd = nullptr, *p_ti = nullptr, *p_cu = nullptr, *p_fl = nullptr, *p_cnt = nullptr, *p_off = nullptr;
   bad |= up(nullptr, (size_t)d->Nvox * 8, &p_cnt) | up(nullptr, ((size_t)d->Nvox + 1) * 8, &p_off);
   if (!bad && npt) {
      // per-point scratch of every voxel: nearest hit, its triangle, cut links, flags
      const size_t sizes[4] = {npt * 8, npt * 4, npt * 2, npt};
      const void **dst[4] = {&p_nd, &p_ti, &p_cu, &p_fl};
      for (int k = 0; k < 4 && !bad; k++) bad |= up(nullptr, sizes[k], dst[k]);
   }
   if (bad) {
      freeall();
      return fail(PFFDTD_ECUDA, "voxeliser: device allocation / upload failed: %s", cudaGetErrorString(cudaGetLastError()));
   }
   lap("allocations + uploads");
   a.ndist = (double *)p_nd, a.tidx = (int *)p_ti, a.cut = (unsigned short *)p_cu, a.fl = (unsigned char *)p_fl;
   a.count = (long long *)p_cnt, a.off = (const long long *)p_off;
   pffdtd_vox *R = new pffdtd_vox();
   R->NN = d->NN;
   cudaError_t ce = cudaSuccess;
   if (d->Nvox && npt) {
      // 1. ray casting, one block per voxel; 2. the voxels' boundary-point counts -> offsets; 3. compaction on the device, in the
      // reference's order (voxel by voxel, ascending inside a voxel: vox_scene.py:246-279, 343-366)
      pf::k_vox_calc_adj<<<(unsigned)d->Nvox, 256>>>(a);
      ce = cudaGetLastError();
      lap("k_vox_calc_adj");
      std::vector<long long> cnt((size_t)d->Nvox), off((size_t)d->Nvox + 1, 0);
      if (ce == cudaSuccess) ce = cudaMemcpy(cnt.data(), a.count, cnt.size() * 8, cudaMemcpyDeviceToHost);
      if (ce == cudaSuccess) {
         for (i64 v = 0; v < d->Nvox; v++) off[(size_t)v + 1] = off[(size_t)v] + cnt[(size_t)v];
         ce = cudaMemcpy((void *)a.off, off.data(), off.size() * 8, cudaMemcpyHostToDevice);
      }
      const size_t nb = (size_t)off[(size_t)d->Nvox];
      lap("counts -> offsets");
      if (ce == cudaSuccess && nb) {
         const void *o1 = nullptr, *o2 = nullptr, *o3 = nullptr, *o4 = nullptr;
         if (up(nullptr, nb * 8, &o1) | up(nullptr, nb * (size_t)d->NN, &o2) | up(nullptr, nb * 4, &o3) | up(nullptr, nb * 8, &o4)) {
            ce = cudaGetLastError();
            if (ce == cudaSuccess) ce = cudaErrorMemoryAllocation;
         } else {
            a.o_bn = (long long *)o1, a.o_adj = (unsigned char *)o2, a.o_tidx = (int *)o3, a.o_ndist = (double *)o4;
            pf::k_vox_emit<<<(unsigned)d->Nvox, 256>>>(a);
            ce = cudaGetLastError();
            lap("result allocation + k_vox_emit");
            R->bn.resize(nb), R->adj.resize(nb * (size_t)d->NN), R->tidx.resize(nb), R->ndist.resize(nb);
            if (ce == cudaSuccess) ce = cudaMemcpy(R->bn.data(), a.o_bn, nb * 8, cudaMemcpyDeviceToHost);
            if (ce == cudaSuccess) ce = cudaMemcpy(R->adj.data(), a.o_adj, nb * (size_t)d->NN, cudaMemcpyDeviceToHost);
            if (ce == cudaSuccess) ce = cudaMemcpy(R->tidx.data(), a.o_tidx, nb * 4, cudaMemcpyDeviceToHost);
            if (ce == cudaSuccess) ce = cudaMemcpy(R->ndist.data(), a.o_ndist, nb * 8, cudaMemcpyDeviceToHost);
         }
      }
   }
   lap("results to the host");
   freeall();
   lap("cudaFree");
   if (ce != cudaSuccess) {
      delete R;
      return fail(PFFDTD_ECUDA, "voxeliser kernel: %s", cudaGetErrorString(ce));
   }
   *out = R;
   return PFFDTD_OK;
}

extern "C" int64_t pffdtd_vox_count(const pffdtd_vox *r) { return r ? (int64_t)r->bn.size() : -1; }

extern "C" int pffdtd_vox_read(const pffdtd_vox *r, int64_t *bn_ixyz, uint8_t *adj, int32_t *tidx, double *ndist) {
   if (!r || !bn_ixyz || !adj || !tidx || !ndist) return fail(PFFDTD_EINVAL, "NULL argument");
   if (r->bn.empty()) return PFFDTD_OK;
   memcpy(bn_ixyz, r->bn.data(), r->bn.size() * 8);
   memcpy(adj, r->adj.data(), r->adj.size());
   memcpy(tidx, r->tidx.data(), r->tidx.size() * 4);
   memcpy(ndist, r->ndist.data(), r->ndist.size() * 8);
   return PFFDTD_OK;
}

extern "C" int pffdtd_vox_free(pffdtd_vox *r) {
   delete r;
   return PFFDTD_OK;
}

// VoxGridBase.fill: triangle lists of every voxel -- see vox.cuh (k_vox_fill)
struct pffdtd_voxfill {
   std::vector<int64_t> off;
   std::vector<int32_t> tri;
};

extern "C" int pffdtd_voxfill_run(const pffdtd_voxfill_desc *d, int device, pffdtd_voxfill **out) {
   if (!d || !out) return fail(PFFDTD_EINVAL, "NULL argument");
   if (d->struct_size != (int32_t)sizeof(pffdtd_voxfill_desc)) return fail(PFFDTD_EINVAL, "pffdtd_voxfill_desc size mismatch");
   if (d->Nvox < 0 || d->Ntris < 0 || d->Nvox > 0x7fffffffLL / 8 * 8 || d->Ntris > 0x7fffffffLL - 32)
      return fail(PFFDTD_EINVAL, "bad voxel grid description");
   int ndev = 0;
   CU(cudaGetDeviceCount(&ndev));
   if (device < 0 || device >= ndev) return fail(PFFDTD_ECUDA, "no CUDA device %d (%d visible)", device, ndev);
   CU(cudaSetDevice(device));
   std::vector<void *> dev;
   auto freeall = [&]() {
      for (void *p : dev) cudaFree(p);
   };
   auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
      void *p = nullptr;
      if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return 1;
      dev.push_back(p);
      if (bytes && src && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
      *dst = p;
      return 0;
   };
   pf::VoxFillArgs a;
   memset(&a, 0, sizeof a);
   a.Nvox = d->Nvox, a.Ntris = d->Ntris;
   const size_t t3 = (size_t)d->Ntris * 24, v3 = (size_t)d->Nvox * 24;
   int bad = up(d->vbmin, v3, (const void **)&a.vbmin) | up(d->vbmax, v3, (const void **)&a.vbmax) | up(d->v, t3 * 3, (const void **)&a.v);
   bad |= up(d->nor, t3, (const void **)&a.nor) | up(d->cent, t3, (const void **)&a.cent) | up(d->bmin, t3, (const void **)&a.bmin) | up(d->bmax, t3, (const void **)&a.bmax);
   const void *p_cnt = nullptr, *p_off = nullptr;
   bad |= up(nullptr, (size_t)d->Nvox * 8, &p_cnt) | up(nullptr, ((size_t)d->Nvox + 1) * 8, &p_off);
   if (bad) {
      freeall();
      return fail(PFFDTD_ECUDA, "voxel grid fill: device allocation / upload failed: %s", cudaGetErrorString(cudaGetLastError()));
   }
   a.count = (long long *)p_cnt, a.off = (const long long *)p_off;
   pffdtd_voxfill *R = new pffdtd_voxfill();
   R->off.assign((size_t)d->Nvox + 1, 0);
   cudaError_t ce = cudaSuccess;
   if (d->Nvox && d->Ntris) {
      const unsigned blocks = (unsigned)((d->Nvox + 7) / 8);
      std::vector<long long> cnt((size_t)d->Nvox);
      pf::k_vox_fill<false><<<blocks, 256>>>(a);
      ce = cudaGetLastError();
      if (ce == cudaSuccess) ce = cudaMemcpy(cnt.data(), a.count, cnt.size() * 8, cudaMemcpyDeviceToHost);
      if (ce == cudaSuccess) {
         for (i64 v = 0; v < d->Nvox; v++) R->off[(size_t)v + 1] = R->off[(size_t)v] + cnt[(size_t)v];
         ce = cudaMemcpy((void *)a.off, R->off.data(), R->off.size() * 8, cudaMemcpyHostToDevice);
      }
      const size_t total = (size_t)R->off[(size_t)d->Nvox];
      if (ce == cudaSuccess && total) {
         void *q = nullptr;
         ce = cudaMalloc(&q, total * 4);
         if (ce == cudaSuccess) {
            dev.push_back(q);
            a.tri = (int *)q;
            pf::k_vox_fill<true><<<blocks, 256>>>(a);
            ce = cudaGetLastError();
            R->tri.resize(total);
            if (ce == cudaSuccess) ce = cudaMemcpy(R->tri.data(), a.tri, total * 4, cudaMemcpyDeviceToHost);
         }
      }
   }
   freeall();
   if (ce != cudaSuccess) {
      delete R;
      return fail(PFFDTD_ECUDA, "voxel grid fill: %s", cudaGetErrorString(ce));
   }
   *out = R;
   return PFFDTD_OK;
}

extern "C" int64_t pffdtd_voxfill_count(const pffdtd_voxfill *r) { return r ? (int64_t)r->tri.size() : -1; }

extern "C" int pffdtd_voxfill_read(const pffdtd_voxfill *r, int64_t *off, int32_t *tri) {
   if (!r || !off) return fail(PFFDTD_EINVAL, "NULL argument");
   memcpy(off, r->off.data(), r->off.size() * 8);
   if (!r->tri.empty()) {
      if (!tri) return fail(PFFDTD_EINVAL, "NULL argument");
      memcpy(tri, r->tri.data(), r->tri.size() * 4);
   }
   return PFFDTD_OK;
}

extern "C" int pffdtd_voxfill_free(pffdtd_voxfill *r) {
   delete r;
   return PFFDTD_OK;
}
