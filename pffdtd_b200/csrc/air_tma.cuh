// air_tma.cuh -- the interior air update as a persistent 2.5-D blocked sweep for sm_100a.
//
// Replaces KernelAirCart of the reference (c_cuda/gpu_engine.h:220-242; one thread per node, seven
// scalar loads through L1) with a tile kernel built for Blackwell:
//   * a CTA owns a (TY x TZ) tile of the y-z plane and sweeps consecutive x-planes;
//   * EVERY operand of a plane arrives by TMA (cp.async.bulk.tensor.3d, mbarrier complete_tx) in one
//     shared-memory stage: the u1 tile with its 1-node halo, the u0 tile and the tile's words of the
//     node mask; S stages deep, so the loads of the next planes are in flight while a plane is
//     computed and no thread ever waits on a global load.  Out-of-grid parts of a box are zero-filled
//     by the TMA unit, so ragged tiles need no special code;
//   * warp-specialised: one producer warp issues the TMA loads, NW consumer warps compute; they meet
//     only through full/empty mbarriers (no CTA-wide barrier in the loop);
//   * a consumer thread owns RPT rows x one 16-byte vector of z (4 fp32 / 2 fp64) and keeps the x-1,
//     x, x+1 values of its own columns in registers, rotating them along the sweep: the +-x taps never
//     touch memory again, +-y taps of inner rows come from the thread's own registers, only the two
//     rows next to the thread's strip and the +-z end taps are read from shared memory;
//   * the new u0 goes straight to HBM with 128-bit stores; masked lanes store the old value back, so
//     every store is a full aligned vector;
//   * persistent grid with an exactly balanced static partition of the tile-planes (see AirJob).
// Arithmetic is the reference CPU engine's (cpu_engine.h:182-189): a1*u1 - u0, then six separately
// rounded a2*u1[nb] products added in the order +x -x +y -y +z -z.
//
// Algorithmic HBM traffic per node: u1 read once (4|8 B) + u0 read (4|8 B) + u0 write (4|8 B)
// + 1 mask bit = 12.125 B fp32 / 24.125 B fp64.
#pragma once
#include <cstdint>
#include <algorithm>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace pf {

// In-kernel boundary work ("service warp", kernels built with SVC): a list of the SPARSE special nodes of every tile-plane --
// rigid-boundary nodes (cpu_engine.h:234-287) of tile-planes that hold few of them, and the z faces of the absorbing shell
// (cpu_engine.h:225-229) -- is processed by one extra warp from the shared-memory stages, where the node's seven (thirteen) u1
// values and its old u0 already are: the warp writes the finished value into the stage's u0 tile (and sets the node's bit in the
// stage's mask words) BEFORE the consumers read that tile, so the consumers carry it to HBM inside their ordinary full-vector
// stores (and into the halo mirrors).  List-driven kernels pay a 32-byte DRAM sector per tap for exactly these nodes (z walls:
// one or two nodes per row): round 1's k_rigid moved 86 B and k_abc_faces 76 B per node for them.  Dense runs of boundary nodes
// (walls perpendicular to x or y, contiguous along z) stay with the list kernel, where they coalesce.
// Entry: bits 0-6 column in the tile, 7-12 row in the tile, 13-15 kind (0 rigid, 1 shell face Q=1, 7 none), 16-27 adjacency.
struct AirSvc {
   const uint32_t *list;  // entries grouped by (tile, plane)
   const uint32_t *off;   // [tiles][pitch]: entries of plane x of a tile = [off[x], off[x+1])
   int pitch;             // Nx + 1
};
#define PF_SVC_NONE 0xE000u
#define PF_SVC_CAP 192  // most entries of one tile-plane (a stage holds them; segments are padded to 4 entries = 16 bytes)

struct AirTma {
   bool ok = false;
   std::string why = "not set up";
   int precision = 0, fcc = 0;
   i64 Nx = 0, Ny = 0, Nz = 0, Nzp = 0, mwpr = 0;
   CUtensorMap map_u1[2];  // haloed boxes over u[0], u[1]
   CUtensorMap map_u0[2];  // plain tiles over u[0], u[1]
   CUtensorMap map_mk;     // mask words
   void *base[2] = {nullptr, nullptr};
   void *mask = nullptr;
   int cfg = 0;         // tile configuration, see PF_AIR_CONFIGS
   int z_edge = 0;      // the shell node z = Nz-2 opens a z tile of this configuration (7-point: handled in the kernel; 13-point: no fused step)
   int xc = 0;          // planes per (long) x-chunk of the work order, 0 = default
   int sm_count = 148;
   int slots = 0;       // resident CTAs of the chosen configuration on this device
   int *ctr = nullptr;  // device: {next item, CTAs done}, zero between launches
   int svc = 0;         // the configuration has the service warp
   int ty = 0, tzn = 0; // tile shape in nodes (rows, columns)
   AirSvc sv{nullptr, nullptr, 0};  // the service warp's lists for this tile shape (engine-built), null = none
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
   uint32_t ok;
   const uint32_t a = smem_u32(bar);
   do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(a), "r"(parity)
          : "memory");
   } while (!ok);
}
// plain bulk copy global -> shared, completing on the same mbarrier as the stage's tensor loads (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
   asm volatile(
       "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
           smem_u32(dst)),
       "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
       : "memory");
}

template <typename Real> struct VecOf;
template <> struct VecOf<float> { typedef float4 type; };
template <> struct VecOf<double> { typedef double2 type; };

template <typename Real, int VEC>
__device__ __forceinline__ void ld_vec(const Real *p, Real (&d)[VEC]) {
   typedef typename VecOf<Real>::type V;
   const V v = *reinterpret_cast<const V *>(p);
   const Real *s = reinterpret_cast<const Real *>(&v);
#pragma unroll
   for (int k = 0; k < VEC; k++) d[k] = s[k];
}
template <typename Real, int VEC>
__device__ __forceinline__ void st_vec(Real *p, const Real (&d)[VEC]) {
   typedef typename VecOf<Real>::type V;
   V v;
   Real *s = reinterpret_cast<Real *>(&v);
#pragma unroll
   for (int k = 0; k < VEC; k++) s[k] = d[k];
   *reinterpret_cast<V *>(p) = v;
}

// ---------------------------------------------------------------- packed fp32 arithmetic (sm_100a: FFMA2 / FADD2)
// Two fp32 operations per issued instruction.  The air kernels are issue-limited next to their HBM limit (round 1: 84 % of the
// issue slots busy at 92 % of the copy bandwidth), and 56 of the ~90 instructions of a row-vector were scalar FMUL / FADD.
// Every operation stays a separately rounded IEEE multiply or add in the reference's order (cpu_engine.h:182-189), element by
// element, so the bits do not change:
//  * a product a*b is fma.rn(a, b, nz) with nz = -0.0 (exactly RN(a*b), signed zeros included).  nz comes in as a kernel
//    argument: ptxas contracts mul.rn.f32x2 / fma(.., -0.0 literal) + add.rn.f32x2 into ONE FFMA2 even under --fmad=false
//    (verified in SASS), which would round once instead of twice; an addend it cannot see through keeps the two roundings;
//  * sums are add.rn.f32x2 / sub.rn.f32x2 (FADD2).
typedef unsigned long long u64;
struct F4 {
   u64 lo, hi;  // elements 0,1 and 2,3 of a 16-byte vector
};
__device__ __forceinline__ u64 pk2(float lo, float hi) {
   u64 r;
   asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
   return r;
}
__device__ __forceinline__ void upk2(u64 p, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b, u64 nz) {
   u64 r;
   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
   return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
   u64 r;
   asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
   return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
   u64 r;
   asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
   return r;
}
__device__ __forceinline__ F4 f4_ld(const float *p) {
   const float4 v = *reinterpret_cast<const float4 *>(p);
   return F4{pk2(v.x, v.y), pk2(v.z, v.w)};
}
__device__ __forceinline__ F4 f4_ld(const double *) { return F4{0, 0}; }  // (never called: the packed path is fp32 only)
__device__ __forceinline__ F4 f4_mul(F4 a, u64 c, u64 nz) { return F4{mul2(a.lo, c, nz), mul2(a.hi, c, nz)}; }
__device__ __forceinline__ F4 f4_add(F4 a, F4 b) { return F4{add2(a.lo, b.lo), add2(a.hi, b.hi)}; }
__device__ __forceinline__ F4 f4_sub(F4 a, F4 b) { return F4{sub2(a.lo, b.lo), sub2(a.hi, b.hi)}; }
__device__ __forceinline__ void f4_get(F4 a, float (&d)[4]) {
   upk2(a.lo, d[0], d[1]);
   upk2(a.hi, d[2], d[3]);
}
__device__ __forceinline__ void f4_get(F4, double (&)[2]) {}

// LZ = lanes of a warp along z (32, 16 or 8); the other 32/LZ = LR lane groups take further rows, so a warp covers
// LZ vectors x LR*RPT rows.  Narrow tiles cut the padding a grid pays when Nz is not a multiple of 32 vectors.
template <typename Real, int RPT, int NW, int S, int LZ = 32, bool SVC = false>
struct AirCfg {
   static constexpr int VEC = 16 / (int)sizeof(Real);
   static constexpr int LR = 32 / LZ;
   static constexpr int TZ = LZ * VEC;
   static constexpr int TY = NW * RPT * LR;
   static constexpr int BZ = TZ + 2 * VEC;  // u1 box starts one vector left of the tile: 16-byte aligned columns
   static constexpr int ROWS = TY + 2;
   static constexpr int MKW = 4;            // mask words per tile row in a stage (TZ/32 used, 16-byte TMA minimum)
   static constexpr int U1_BYTES = ROWS * BZ * (int)sizeof(Real);
   static constexpr int U0_BYTES = TY * TZ * (int)sizeof(Real);
   static constexpr int MK_BYTES = TY * MKW * 4;
   static constexpr int U0_OFF = (U1_BYTES + 127) / 128 * 128;
   static constexpr int MK_OFF = U0_OFF + U0_BYTES;
   static constexpr int SV_CAP = 192;       // (SVC) most list entries of one tile-plane (PF_SVC_CAP: the engine's builder keeps to it)
   static constexpr int SV_OFF = (MK_OFF + MK_BYTES + 15) / 16 * 16;
   static constexpr int STAGE_PITCH = (SV_OFF + (SVC ? SV_CAP * 4 : 0) + 127) / 128 * 128;
   // stages, full/empty/patched barriers, item headers, entry counts
   static constexpr int SMEM_BYTES = S * STAGE_PITCH + 3 * S * 8 + S * 16 + S * 4 + 128;
   static constexpr int THREADS = (NW + 1 + (SVC ? 1 : 0)) * 32;  // NW consumer warps + 1 TMA producer warp (+ 1 service warp)
};

// ---------------------------------------------------------------- work decomposition
// The job "planes [x_begin, x_begin+n) x all y-z tiles" is cut into ITEMS = (x-chunk, tile): consecutive planes
// of one tile.  A persistent grid (one CTA per resident slot) pulls items from an atomic counter, in the order
// (chunk, tile): CTAs running at the same time work on neighbouring tiles of the same x-chunk, so tile halos
// hit in L2.  Chunks get shorter towards the end of the job (guided self-scheduling), so that all CTAs finish
// within a few planes of each other whatever the per-tile cost; every item costs two extra u1 plane loads.
#define PF_AIR_MAXTAIL 24
struct AirJob {
   int x_begin, n, tz, tiles;  // n planes, tiles = tz*ty
   int nch, n_items;           // chunks, items = nch*tiles
   int Ny, Nz, Nzp;
   i64 plane;                  // Ny*Nzp
   int *ctr;                   // {next item, CTAs done}; left at {0,0} by the last CTA
   // chunk k < n_main covers planes [k*xc, (k+1)*xc) of the job (the last one clipped to n); the guided tail that follows is
   // chunk n_main + j = [tail[j], tail[j+1]).  Arithmetic bounds: the job may have any number of planes.
   int xc, n_main;
   int tail[PF_AIR_MAXTAIL + 1];
};
struct AirSeg {
   int xa, cnt, z0, y0;
};
template <int TZ, int TY>
__device__ __forceinline__ AirSeg air_item(const AirJob &jb, int item) {
   const int k = item / jb.tiles, t = item - k * jb.tiles;
   AirSeg s;
   int b0, b1;
   if (k < jb.n_main) {
      b0 = k * jb.xc;
      b1 = min(b0 + jb.xc, jb.n);
   } else {
      b0 = jb.tail[k - jb.n_main];
      b1 = jb.tail[k - jb.n_main + 1];
   }
   s.xa = jb.x_begin + b0;
   s.cnt = b1 - b0;
   s.z0 = (t % jb.tz) * TZ;
   s.y0 = 1 + (t / jb.tz) * TY;
   return s;
}

// Chunk plan of a job of n planes: chunks of xc planes while more than a tail's worth of planes is left, then lengths that halve
// down to 4, so that the items handed out last are small and all CTAs finish within a few planes of each other.
static void air_plan_chunks(AirJob *jb, int n, int xc) {
   xc = std::max(4, xc);
   jb->n = n, jb->xc = xc;
   const int tail_from = n - std::min(n / 4, 2 * xc);
   jb->n_main = (tail_from + xc - 1) / xc;
   int x = std::min(n, jb->n_main * xc), j = 0;
   jb->tail[0] = x;
   while (x < n) {
      int len = std::max(4, std::min(xc, (n - x + 1) / 2));
      if (j == PF_AIR_MAXTAIL - 1) len = n - x;
      x = std::min(n, x + len);
      jb->tail[++j] = x;
   }
   jb->nch = jb->n_main + j;
}

// Fused extras of the Cartesian step (all optional, `fuse` = 0 gives the plain masked air update):
//  * the absorbing shell (cpu_engine.h:225-229): a node with Q = #{axes on which its index is 1 or
//    N-2} > 0 becomes (v + lQ*u0_old)/(1.0 + lQ), v = the fresh air value (or the untouched old value
//    of a masked node, exactly what the reference's ABC loop sees).  The air kernel only STASHES u0_old of
//    the shell nodes (it has it in registers; a list-driven gather would cost a DRAM line per node) and
//    k_abc_faces finishes them; that keeps every row of the sweep equally cheap, which matters because
//    the persistent partition assumes equal cost per tile-plane;
//  * the halo mirrors (cpu_engine.h:145-172) are applied when a value is WRITTEN instead of before it is
//    read: whoever stores index 2 (N-3) of an axis also stores it to index 0 (N-1).  The new state then
//    leaves the kernel with its face halos complete and the next step needs no mirror pass.
template <typename Real>
struct AirEdge {
   int fuse, x_lo, x_hi, Nx;
   float negzero;  // -0.0f, opaque to the compiler: the addend that makes fma.rn.f32x2 a multiplication (see "packed fp32 arithmetic")
   int folded;  // folded FCC grid (fcc_flag 2): seam row instead of a mirror / shell at the high y end
   int zstash;  // the consumers stash the pre-update values of the shell's z faces for k_abc_faces (0 when the service warp does them)
   Real sl2;    // rigid update: b1 = 2 - sl2*K
   // pre-update values of the shell nodes, for k_abc_faces:
   Real *zold;  // [Nx][Ny][2]    z=1 / z=Nz-2 of every row
   Real *yold;  // [Nx][2][Nzp]   rows y=1 / y=Ny-2
   Real *xold;  // [2][Ny][Nzp]   planes x=1 / x=Nx-2 (global ends only)
   Real lQ1, lQ2, lQ3;
   double den1, den2, den3;     // 1.0 + lQ
   double rden1, rden2, rden3;  // RN(1/den)
};

// a / b correctly rounded, given rb = RN(1/b): q = RN(a*rb) is within an ulp of a/b, the residual
// r = a - b*q is then exact in one fma, and RN(q + r*rb) is the correctly rounded quotient (Markstein's
// theorem; b = 1 + l*Q is far from the all-ones significand it excludes).  Tiny numerators, where the
// residual could underflow, take the general division.  Checked against __ddiv_rn on the device by
// pffdtd_selftest().
__device__ __forceinline__ double div_by_const(double a, double b, double rb) {
   if (a == 0.0) return a;  // +-0 / b (b > 0) keeps its sign; also the common case while the wave has not arrived
   if (fabs(a) < 1e-280) return __ddiv_rn(a, b);
   const double q = __dmul_rn(a, rb);
   const double r = __fma_rn(-b, q, a);
   return __fma_rn(r, rb, q);
}

// (v + lQ*u0_old) / (1.0 + lQ) with the reference's types: numerator in Real, division in double
template <typename Real>
__device__ __forceinline__ Real abc_apply(Real v, Real old, Real lQ, double den, double rden) {
   typedef Ops<Real> O;
   const Real num = O::add(v, O::mul(lQ, old));
   return (Real)div_by_const((double)num, den, rden);
}

__global__ void k_selftest_div(double b, double rb, unsigned long long seed, int per_thread, int as_float, unsigned long long *bad) {
   unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
   unsigned long long nbad = 0;
   for (int i = 0; i < per_thread; i++) {
      x ^= x << 13, x ^= x >> 7, x ^= x << 17;  // xorshift64
      double a;
      if (as_float) {
         a = (double)__uint_as_float((unsigned)(x >> 32) & 0xdfffffffu);
      } else {
         a = __longlong_as_double((long long)(x & 0xbfffffffffffffffull));
      }
      if (isnan(a) || isinf(a)) continue;
      const double q0 = __ddiv_rn(a, b), q1 = div_by_const(a, b, rb);
      if (__double_as_longlong(q0) != __double_as_longlong(q1) && !(q0 == 0.0 && q1 == 0.0)) nbad++;
   }
   if (nbad) atomicAdd(bad, nbad);
}

// The absorbing shell (cpu_engine.h:225-229) for the fused step.  The air kernel leaves the plain air value
// in u0 and stashes the pre-update value of every shell node; this kernel finishes them,
//     u0 = (u0 + lQ*old)/(1.0 + lQ),   lQ = l*Q,  Q = number of axes on which the index is 1 or N-2,
// and repeats the halo mirror of every node it changes.  Thread classes over the planes [xb,xe):
//   X: all interior nodes of the planes on the x shell            (old values in xold, coalesced)
//   Y: rows y=1, y=Ny-2 of the other planes                       (old values in yold, coalesced)
//   Z: z=1, z=Nz-2 of the remaining rows                          (old values in zold, one node per row end)
template <typename Real>
struct FacesArgs {
   Real *u0;
   const Real *zold, *yold, *xold;
   int Nx, Ny, Nz, Nzp, xb, xe, x_lo, x_hi;
   int do_z;    // 0: the z faces were finished inside the air kernel (service warp)
   int folded;  // folded FCC grid: no shell and no mirror at the high y end, row Ny-2 is copied to the seam row Ny-1
   int edges;   // 13-point stencil: halo edges are read too, a changed node is copied to its doubly mirrored positions as well
   int checker; // checkerboard FCC grid (fcc_flag 1): 1 + the parity of the slab's first plane; odd-parity nodes are not grid nodes
   Real lQ1, lQ2, lQ3;
};
template <typename Real>
__device__ __forceinline__ void abc_face_node(const FacesArgs<Real> &a, const int x, const int y, const int z, const Real old, const bool fold,
                                              const bool xshell);
template <typename Real>
__global__ void k_abc_faces(const FacesArgs<Real> a) {
   typedef Ops<Real> O;
   // 2-D launch, no divisions: blockIdx.y walks "lines" -- nx*2 rows of the y faces (x, side), then 2*Ny rows of the x faces
   // (side, y), then (do_z) nx*2 lines of the z faces (x, side) that run along y -- and the x dimension of the grid walks the
   // line (z, or y for the z faces).  (Round 2's first version decoded one linear 64-bit index with three runtime divisions per
   // thread: 12 us on c2 for 4 MB of traffic.)
   const int nx = a.xe - a.xb;
   const int rowsY = nx * 2, rowsX = 2 * a.Ny, rowsZ = a.do_z ? nx * 2 : 0;
   const int w = blockIdx.x * blockDim.x + threadIdx.x;  // position along the line
   int x, y, z;
   Real old;
   auto on_xshell = [&](int xx) { return (a.x_lo && xx == 1) || (a.x_hi && xx == a.Nx - 2); };
   const bool fold = a.folded != 0;
   for (int line = blockIdx.y; line < rowsY + rowsX + rowsZ; line += gridDim.y) {
      if (line < rowsY) {
         // descending x: the planes the air kernel wrote last are still in L2
         const int l = rowsY - 1 - line, side = l & 1;
         x = a.xb + (l >> 1), y = side ? a.Ny - 2 : 1, z = w;
         if (z < 1 || z > a.Nz - 2 || on_xshell(x) || (fold && side)) continue;
         old = a.yold[((i64)x * 2 + side) * a.Nzp + z];
      } else if (line < rowsY + rowsX) {
         const int l = line - rowsY, side = l >= a.Ny ? 1 : 0;
         x = side ? a.Nx - 2 : 1, y = l - side * a.Ny, z = w;
         if (!(side ? a.x_hi : a.x_lo) || x < a.xb || x >= a.xe) continue;
         if (y < 1 || y > a.Ny - 2 || z < 1 || z > a.Nz - 2) continue;
         old = a.xold[((i64)side * a.Ny + y) * a.Nzp + z];
      } else {
         const int l = rowsZ - 1 - (line - rowsY - rowsX), side = l & 1;
         x = a.xb + (l >> 1), y = w, z = side ? a.Nz - 2 : 1;
         if (y < 2 || y > (fold ? a.Ny - 2 : a.Ny - 3) || on_xshell(x)) continue;
         old = a.zold[((i64)x * a.Ny + y) * 2 + side];
      }
      abc_face_node<Real>(a, x, y, z, old, fold, on_xshell(x));
   }
}

// one shell node of k_abc_faces: the update and the halo copies of the changed value
template <typename Real>
__device__ __forceinline__ void abc_face_node(const FacesArgs<Real> &a, const int x, const int y, const int z, const Real old, const bool fold,
                                              const bool xshell) {
   typedef Ops<Real> O;
   // (the unused parity of a checkerboard grid normally holds zeros, on which the update is the identity -- but only exactly so for
   // zeros: leave those nodes alone, as the reference's list does)
   if (a.checker && ((a.checker - 1 + x + y + z) & 1)) return;
   const int Q = (xshell ? 1 : 0) + ((y == 1 || (!fold && y == a.Ny - 2)) ? 1 : 0) + ((z == 1 || z == a.Nz - 2) ? 1 : 0);
   const Real lQ = Q == 1 ? a.lQ1 : (Q == 2 ? a.lQ2 : a.lQ3);
   const i64 P = (i64)a.Ny * a.Nzp;
   Real *p = a.u0 + ((i64)x * a.Ny + y) * a.Nzp + z;
   const Real num = O::add(*p, O::mul(lQ, old));
   const Real v = (Real)__ddiv_rn((double)num, __dadd_rn(1.0, (double)lQ));
   *p = v;
   // halo mirrors of the changed node: one copy per axis on which its index is 2 / N-3 (the seam copy for row Ny-2 of a folded
   // grid); the 13-point stencil also reads halo edges, so there the copies combine
   const i64 oz[3] = {0, z == 2 ? -2 : 0, z == a.Nz - 3 ? 2 : 0};
   const i64 oy[3] = {0, y == 2 ? -2 * (i64)a.Nzp : 0, fold ? (y == a.Ny - 2 ? (i64)a.Nzp : 0) : (y == a.Ny - 3 ? 2 * (i64)a.Nzp : 0)};
   const i64 ox[3] = {0, (a.x_lo && x == 2) ? -2 * P : 0, (a.x_hi && x == a.Nx - 3) ? 2 * P : 0};
#pragma unroll
   for (int ix = 0; ix < 3; ix++)
#pragma unroll
      for (int iy = 0; iy < 3; iy++)
#pragma unroll
         for (int iz = 0; iz < 3; iz++) {
            const int axes = (ix ? 1 : 0) + (iy ? 1 : 0) + (iz ? 1 : 0);
            if (axes == 0 || (ix && !ox[ix]) || (iy && !oy[iy]) || (iz && !oz[iz])) continue;
            if (axes > 1 && !a.edges) continue;
            p[ox[ix] + oy[iy] + oz[iz]] = v;
         }
}

// ---------------------------------------------------------------- the kernel (7-point Cartesian)
// full[s] flips when all bytes of a plane have landed in stage s, empty[s] when all NW consumer warps
// are done with it.  Loads are numbered consecutively over all segments of the CTA; load i uses stage
// i % S.  A segment of cnt planes loads planes xa-1 .. xa+cnt: the first and the last only as u1.
template <typename Real, int RPT, int NW, int S, int MAXR, bool FCC, int LZ, bool SVC, bool FFUSE, bool ZE = false>
__global__ void __maxnreg__(MAXR)
    k_air_tma_cart(const __grid_constant__ CUtensorMap map_u1, const __grid_constant__ CUtensorMap map_u0,
                   const __grid_constant__ CUtensorMap map_mk, Real *__restrict__ u0g, const AirJob jb, const Real a1, const Real a2,
                   const AirEdge<Real> eg, const AirSvc sv) {
   typedef AirCfg<Real, RPT, NW, S, LZ, SVC> C;
   typedef Ops<Real> O;
   constexpr int VEC = C::VEC, BZ = C::BZ, TZ = C::TZ;
   constexpr uint32_t VMASK = (1u << VEC) - 1u;
   // no integer round trip on this pointer: it must stay a SHARED-space pointer so that the tile reads compile to
   // LDS (a pointer rebuilt from an integer becomes generic: LD.E with 64-bit address arithmetic).  There is no
   // static shared memory in this kernel, so the dynamic window starts at offset 0 of the CTA's shared memory.
   extern __shared__ __align__(1024) unsigned char smem[];
   uint64_t *full = (uint64_t *)(smem + S * C::STAGE_PITCH);
   uint64_t *empty = full + S;
   uint64_t *patched = empty + S;  // (SVC) flips when the service warp is done with the load in stage s

   int4 *hdr = (int4 *)(patched + S);  // per stage: the item (xa, cnt, z0, y0) whose first plane it holds; cnt < 0 = stop
   int *nent = (int *)(hdr + S);       // (SVC) per stage: list entries that came with the centre plane

   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const int lz = lane % LZ;                     // this thread's vector within its row
   const int wr = w * C::LR + lane / LZ;         // this thread's row group within the tile ("virtual warp")

   if (tid == 0) {
      for (int s = 0; s < S; s++) {
         mbar_init(&full[s], 1);
         mbar_init(&empty[s], NW + (SVC ? 1 : 0));
         mbar_init(&patched[s], 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
   }
   __syncthreads();

   if (w == NW) {
      // ---------------- producer: fetch items, stream their planes (lane 0 issues; with SVC the whole warp holds the item's list
      // offsets, one or two per lane, and the plane's list entries travel into the stage with a bulk copy on the same barrier)
      int s = 0;
      uint32_t ph = 1u;  // parity of the "previous round released" phase; the first round needs no wait
      bool first = true;
      for (;;) {
         int item = 0;
         if (lane == 0) item = atomicAdd(&jb.ctr[0], 1);
         item = __shfl_sync(0xffffffffu, item, 0);
         const bool stop = item >= jb.n_items;
         AirSeg sg = {0, -1, 0, 0};
         if (!stop) sg = air_item<C::TZ, C::TY>(jb, item);
         const int nq = stop ? 1 : sg.cnt + 2;  // planes xa-1 .. xa+cnt, or the stop marker
         uint32_t o0 = 0u, o1 = 0u;
         const bool lists = SVC && sv.list != nullptr && !stop;
         if (lists) {
            const int tile = ((sg.y0 - 1) / C::TY) * jb.tz + sg.z0 / C::TZ;
            const uint32_t *ob = sv.off + (size_t)tile * sv.pitch + sg.xa;  // ob[j], ob[j+1]: entries of plane xa + j
            o0 = lane <= sg.cnt ? ob[lane] : 0u, o1 = lane + 32 <= sg.cnt ? ob[lane + 32] : 0u;
         }
         for (int q = 0; q < nq; q++) {
            const bool centre = !stop && q >= 1 && q <= sg.cnt;
            uint32_t eb = 0u, ee = 0u;
            if (SVC) {  // (warp-uniform: every lane takes part in the shuffles)
               const int j = centre ? q - 1 : 0;
               const uint32_t a0 = __shfl_sync(0xffffffffu, o0, j & 31), a1 = __shfl_sync(0xffffffffu, o1, j & 31);
               const uint32_t b0 = __shfl_sync(0xffffffffu, o0, (j + 1) & 31), b1 = __shfl_sync(0xffffffffu, o1, (j + 1) & 31);
               eb = j < 32 ? a0 : a1, ee = j + 1 < 32 ? b0 : b1;
               if (!centre || !lists) eb = ee = 0u;
            }
            if (lane == 0) {
               unsigned char *st = smem + s * C::STAGE_PITCH;
               if (!first) mbar_wait(&empty[s], ph);
               if (q == 0) hdr[s] = make_int4(sg.xa, sg.cnt, sg.z0, sg.y0);
               if (stop) {
                  mbar_arrive(&full[s]);
               } else {
                  const uint32_t lbytes = (ee - eb) * 4u;  // segments start and end on multiples of 4 entries
                  if (SVC) nent[s] = (int)(ee - eb);
                  mbar_expect_tx(&full[s], (centre ? C::U1_BYTES + C::U0_BYTES + C::MK_BYTES : C::U1_BYTES) + lbytes);
                  const int x = sg.xa - 1 + q;
                  tma_load_3d(st, &map_u1, &full[s], sg.z0 - VEC, sg.y0 - 1, x);
                  if (centre) {
                     tma_load_3d(st + C::U0_OFF, &map_u0, &full[s], sg.z0, sg.y0, x);
                     tma_load_3d(st + C::MK_OFF, &map_mk, &full[s], (sg.z0 >> 7) << 2, sg.y0, x);  // box start must be 16-byte aligned
                     if (lbytes) bulk_load(st + C::SV_OFF, sv.list + eb, lbytes, &full[s]);
                  }
               }
            }
            if (++s == S) s = 0, ph ^= 1u, first = false;
         }
         if (stop) break;
      }
      // the last CTA out re-arms the counters for the next launch
      if (lane == 0) {
         __threadfence();
         if (atomicAdd(&jb.ctr[1], 1) == (int)gridDim.x - 1) {
            jb.ctr[0] = 0;
            jb.ctr[1] = 0;
            __threadfence();
         }
      }
      return;
   }

   // ---------------- service warp: the sparse special nodes of every centre plane, from the stages (see AirSvc)
   // It walks the same sequence of loads as the consumers.  For the centre plane x in stage s it needs the u1 boxes of x-1, x, x+1
   // (stages s-1, s, s+1), patches the u0 tile and the mask words of stage s and arrives on patched[s]; patched[] flips once per
   // load (also for the planes that are only read as x-1 / x+1), so its parity follows the ring like full[] / empty[].  It counts as
   // one more consumer on empty[]: a stage is refilled only after this warp has read its last tap from it.
   if constexpr (SVC) {
      if (w == NW + 1) {
         constexpr int NN = FCC ? 12 : 6;
         struct R2 {
            int s;
            uint32_t ph;
            __device__ __forceinline__ void next() {
               if (++s == S) s = 0, ph ^= 1u;
            }
         };
         auto done = [&](uint64_t *bar) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
         };
         const Real lQ1 = eg.lQ1;
         const double den1 = eg.den1, rden1 = eg.rden1;
         R2 g0{0, 0u};
         for (;;) {
            mbar_wait(&full[g0.s], g0.ph);
            const int4 h = hdr[g0.s];
            if (h.y < 0) break;
            const int cnt = h.y;
            done(&patched[g0.s]);
            R2 gm = g0, gc = g0;
            gc.next();
            mbar_wait(&full[gc.s], gc.ph);
            for (int j = 0; j < cnt; j++) {
               R2 gu = gc;
               gu.next();
               mbar_wait(&full[gu.s], gu.ph);
               const int ne = nent[gc.s];  // the plane's entries came into its stage with the plane (no global load in this warp)
               unsigned char *stc = smem + gc.s * C::STAGE_PITCH;
               const Real *sc = (const Real *)stc, *sm = (const Real *)(smem + gm.s * C::STAGE_PITCH),
                          *su = (const Real *)(smem + gu.s * C::STAGE_PITCH);
               Real *u0s = (Real *)(stc + C::U0_OFF);
               uint32_t *mks = (uint32_t *)(stc + C::MK_OFF);
               const uint32_t *ents = (const uint32_t *)(stc + C::SV_OFF);
               for (int i = lane; i < ne; i += 32) {
                  const uint32_t ent = ents[i];
                  const unsigned kind = (ent >> 13) & 7u;
                  if (kind < 2u) {
                     const int c = (int)(ent & 127u), r = (int)((ent >> 7) & 63u);
                     const unsigned adj = kind == 1u ? 0xfffu : (ent >> 16) & 0xfffu;
                     const int idx = (r + 1) * BZ + VEC + c;
                     const Real uc = sc[idx], uo = u0s[r * TZ + c];
                     Real b1 = a1;
                     if (kind == 0u) b1 = O::sub((Real)2.0, O::mul(eg.sl2, (Real)__popc(adj)));
                     Real p = O::sub(O::mul(b1, uc), uo);
                     Real t[NN];
                     if constexpr (!FCC) {  // +x -x +y -y +z -z (cpu_engine.h:249-254)
                        t[0] = su[idx], t[1] = sm[idx], t[2] = sc[idx + BZ], t[3] = sc[idx - BZ], t[4] = sc[idx + 1], t[5] = sc[idx - 1];
                     } else {  // cpu_engine.h:273-284
                        t[0] = su[idx + BZ], t[1] = sm[idx - BZ], t[2] = sc[idx + BZ + 1], t[3] = sc[idx - BZ - 1];
                        t[4] = su[idx + 1], t[5] = sm[idx - 1], t[6] = su[idx - BZ], t[7] = sm[idx + BZ];
                        t[8] = sc[idx + BZ - 1], t[9] = sc[idx - BZ + 1], t[10] = su[idx - 1], t[11] = sm[idx + 1];
                     }
#pragma unroll
                     for (int q = 0; q < NN; q++) p = O::add(p, O::mul(((adj >> q) & 1u) ? a2 : (Real)0.0, t[q]));
                     if (kind == 1u) {
                        p = abc_apply<Real>(p, uo, lQ1, den1, rden1);
                        const int zc = (h.z & 127) + c;
                        atomicOr(&mks[r * C::MKW + (zc >> 5)], 1u << (zc & 31));
                     }
                     u0s[r * TZ + c] = p;
                  }
               }
               done(&patched[gc.s]);
               asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
               done(&empty[gm.s]);
               gm = gc;
               gc = gu;
            }
            done(&patched[gc.s]);  // the last plane was only ever an "x+1" plane
            done(&empty[gm.s]);
            done(&empty[gc.s]);
            g0 = gc;
            g0.next();
         }
         return;
      }
   }

   // ---------------- consumers
   // ring position of a load: stage index and phase parity, advanced incrementally (no div/mod in the loop)
   struct Ring {
      int s;
      uint32_t ph;
      __device__ __forceinline__ void next() {
         if (++s == S) s = 0, ph ^= 1u;
      }
   };
   auto wait_full = [&](const Ring &g) { mbar_wait(&full[g.s], g.ph); };
   auto wait_patched = [&](const Ring &g) {  // the centre plane's u0 tile and mask words are final
      if constexpr (SVC) mbar_wait(&patched[g.s], g.ph);
   };
   auto release = [&](const Ring &g) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[g.s]);
   };
   auto stage = [&](const Ring &g) -> const unsigned char * { return smem + g.s * C::STAGE_PITCH; };
   const int soff = (wr * RPT + 1) * BZ + VEC + VEC * lz;  // strip row 0 inside the u1 box (box row 0 is y0-1)
   const int u0off = C::U0_OFF + ((wr * RPT) * TZ + VEC * lz) * (int)sizeof(Real);
   const int Ny = jb.Ny, Nz = jb.Nz, Nzp = jb.Nzp;
   const bool fuse = eg.fuse != 0;

   Ring g0{0, 0u};  // first plane of the current item
   for (;;) {
      wait_full(g0);
      const int4 h = hdr[g0.s];
      if (h.y < 0) break;  // stop marker
      const AirSeg sg = {h.x, h.y, h.z, h.w};
      // this thread's strip: rows y0 + wr*RPT + r, columns z0 + VEC*lz .. +VEC-1
      const int zv = sg.z0 + VEC * lz;
      const int ybase = sg.y0 + wr * RPT;
      const int mkoff = C::MK_OFF + ((wr * RPT) * C::MKW + ((zv & 127) >> 5)) * 4;  // the stage holds the words of z0 & ~127 ..
      const int mshift = zv & 31;
      int nrow = 0;  // active rows of the strip; vectors entirely in the far halo/padding are never touched
#pragma unroll
      for (int r = 0; r < RPT; r++) nrow += (zv < Nz - 1 && (ybase + r) <= Ny - 2) ? 1 : 0;
      // Roles of the fused step:
      //   zlo_tile / zhi_tile (warp-uniform): the tile holds the low / high z end;
      //   khs, khm (per thread): position of z=Nz-2 (shell) and z=Nz-3 (mirror source) in this vector, else -1;
      //   yrole (per row, 4 bits each): bit0 row on the y shell, bit1 y==2, bit2 y==Ny-3 (mirror sources), bit3 y==Ny-2 of a folded
      //   FCC grid (the seam: its row is copied to Ny-1; the folded grid has neither a shell nor a mirror at its high y end)
      const bool zlo_tile = fuse && sg.z0 == 0;
      const bool zhi_tile = fuse && sg.z0 + TZ > Nz - 3;
      const int khs = (unsigned)(Nz - 2 - zv) < (unsigned)VEC ? Nz - 2 - zv : -1;
      const int khm = (unsigned)(Nz - 3 - zv) < (unsigned)VEC ? Nz - 3 - zv : -1;
      unsigned yrole = 0;
      if (fuse) {
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            const int y = ybase + r;
            const bool fold = eg.folded != 0;
            yrole |= (((y == 1 || (!fold && y == Ny - 2)) ? 1u : 0u) | (y == 2 ? 2u : 0u) | ((!fold && y == Ny - 3) ? 4u : 0u) |
                      ((fold && y == Ny - 2) ? 8u : 0u))
                     << (4 * r);
         }
      }

      if constexpr (FCC) {
      // ---- the end of a row-vector's update in the FUSED 13-point step (FFUSE kernels): the fused step's extras, then the stores
         // (the 7-point path below has the same logic inline; kept apart: sharing it cost the 7-point kernel 14 % on B200).  `o` = the new values (masked elements carry their stage value), `u0v` = the old ones.
         //  * rows / planes on the absorbing shell keep the plain air value and stash their old values for k_abc_faces;
         //  * mirror-on-write: whoever holds index 2 / N-3 of an axis also writes the halo at 0 / N-1 -- inside the vector (or by a
         //    shuffle / one scalar store) for z, as whole-row copies for y and x; the 13-point stencil also reads halo EDGES, so its
         //    x copies carry the y copies too (the row vector already carries the z mirror); the seam row of a folded grid is one
         //    more row copy.
         // Everything lane-dependent is written as selects / single predicated stores: a divergent branch here would make the
         // one lane at a z end run the rest of the step on its own.
            auto emit = [&](const int r, const int x, const unsigned xrole, const unsigned am, Real(&o)[VEC], const Real(&u0v)[VEC], Real *dst,
                         Real *zo) {
            const unsigned yr = (yrole >> (4 * r)) & 15u;
            const bool shell = ((yr | xrole) & 1u) != 0;  // uniform within a row group
            if (shell) {
               // row / plane on the absorbing shell: one extra vector store (the planes on the x shell take precedence)
               Real *sp = (xrole & 1u) ? eg.xold + (((i64)(x == 1 ? 0 : 1) * Ny + (ybase + r)) * Nzp + zv)
                                       : eg.yold + ((((i64)x * 2 + ((ybase + r) == 1 ? 0 : 1)) * Nzp) + zv);
               st_vec<Real, VEC>(sp, u0v);
            }
            // Mirror targets that live in ANOTHER active lane's vector are delivered by a shuffle (that lane stores
            // its whole vector; a scalar store from here would race with it); only a target in a vector nobody
            // stores (the row's far padding) is written directly.
            __syncwarp(am);  // the row groups may have diverged on `shell`
            if (zlo_tile) {  // warp-uniform: the tile starts at z = 0
               // lane 0 holds z=1 (shell): stash its pre-update value; z=2 -> z=0 mirror
               if (eg.zstash && lz == 0 && !shell) zo[0] = u0v[1];
               if constexpr (VEC >= 4) {
                  o[0] = (lz == 0) ? o[2] : o[0];
               } else {
                  const Real t = __shfl_down_sync(am, o[0], 1);  // fp64: z=2 is the first element of the next lane
                  o[0] = (lz == 0) ? t : o[0];
               }
            }
            Real vm = o[0];  // value of z = Nz-3 if this thread holds it
            if (zhi_tile) {  // warp-uniform: the tile contains z = Nz-3 .. Nz-1
               Real vs = u0v[0];
#pragma unroll
               for (int k = 1; k < VEC; k++) {
                  vs = (k == khs) ? u0v[k] : vs;
                  vm = (k == khm) ? o[k] : vm;
               }
               if (eg.zstash && khs >= 0 && !shell) zo[1] = vs;  // z = Nz-2 (shell)
#pragma unroll
               for (int k = 0; k + 2 < VEC; k++) o[k + 2] = (k == khm) ? o[k] : o[k + 2];  // z=Nz-3 -> z=Nz-1 inside the vector
               // a vector that starts at z=Nz-2 holds z=Nz-1 as element 1 and finds z=Nz-3 at the end of the previous lane's
               // (never the first vector of a tile: the engine refuses the fused step for such grids, AirTma::z_edge)
               __syncwarp(am);
               const Real t = __shfl_up_sync(am, o[VEC - 1], 1);
               o[1] = (khs == 0) ? t : o[1];
            }
            // Every vector of an active row is stored, fully masked ones too: it may hold the z halo (mirror-on-write: a masked
            // node at z = 1 / Nz-2 must not keep the halo next to it from being refreshed) or a node the service warp finished
            // in the stage.  Masked elements carry their stage value.
            const bool tailz = zhi_tile && khm + 2 == VEC;  // z=Nz-1 opens the next vector, which nobody stores
            st_vec<Real, VEC>(dst, o);
            if (tailz) dst[VEC] = vm;
            if (((yr | xrole) & 14u) != 0u) {
               // mirror source of a y / x halo (or the seam): the same row goes there as well (warp-uniform, a few rows / planes)
#pragma unroll 1
               for (int tx = 0; tx < 3; tx++) {
                  const bool xon = tx == 0 || (tx == 1 ? (xrole & 2u) != 0 : (xrole & 4u) != 0);
                  const i64 xo = tx == 0 ? 0 : (tx == 1 ? -2 * jb.plane : 2 * jb.plane);
#pragma unroll 1
                  for (int ty = 0; ty < 4; ty++) {
                     const bool yon = ty == 0 || (ty == 1 ? (yr & 2u) != 0 : ty == 2 ? (yr & 4u) != 0 : (yr & 8u) != 0);
                     const i64 yo = ty == 0 ? 0 : (ty == 1 ? -2 * (i64)Nzp : ty == 2 ? 2 * (i64)Nzp : (i64)Nzp);
                     if (xon && yon && (tx | ty) != 0) {
                        Real *d = dst + xo + yo;
                        st_vec<Real, VEC>(d, o);
                        if (tailz) d[VEC] = vm;
                     }
                  }
               }
            }
         };

         // ---- 13-point FCC (cpu_engine.h:205-216; the same stencil on the checkerboard and on the folded grid):
         // the taps of the planes x-1 and x+1 are (y+-1, z) and (y, z+-1), those of plane x are (y+-1, z+-1).
         // Rows y-1, y, y+1 of the three planes live in registers and rotate along the sweep (x+1 -> x -> x-1), so a
         // plane costs three vector LDS per row; the z-1 / z+1 end taps come from the neighbouring lanes by shuffle,
         // only lanes 0 and 31 read theirs from the stage (the first version read all nine rows and eight scalar taps
         // from shared memory: 73 wavefronts per row-vector, the LSU data pipe 94 % busy -- the kernel's bound).
         // Every lane runs the loads and shuffles (out-of-grid parts of a box are zero-filled); only the store is
         // predicated.  The sum runs in the reference's order.
         const bool edge_lane = lz == 0 || lz == LZ - 1;
         const int eoff = lz == 0 ? -1 : VEC;  // where an edge lane finds its out-of-row neighbour in a box row
         auto zleft = [&](const Real(&v)[VEC], const Real e) {
            const Real t = __shfl_up_sync(0xffffffffu, v[VEC - 1], 1);
            return lz == 0 ? e : t;
         };
         auto zright = [&](const Real(&v)[VEC], const Real e) {
            const Real t = __shfl_down_sync(0xffffffffu, v[0], 1);
            return lz == LZ - 1 ? e : t;
         };
         if constexpr (sizeof(Real) == 4) {
            // ---- packed fp32 version (see "packed fp32 arithmetic"): the nine rows are kept as PRODUCTS a2*u1, made once when a plane
            // arrives (3 row vectors per plane) instead of 12 times per node; the four same-column terms of each half of the sum
            // are FADD2s on element pairs, the eight z-shifted terms stay per element (their neighbours sit one element over).
            // 54 floating-point instructions per row-vector instead of 104; the order of the sum is the reference's.
            const u64 A1 = pk2((float)a1, (float)a1), A2 = pk2((float)a2, (float)a2), NZ = pk2(eg.negzero, eg.negzero);
            Ring gm = g0;  // plane x-1
            Ring gc = g0;  // plane x
            gc.next();
            wait_full(gc);
            F4 qm0[RPT], qm1[RPT], qm2[RPT], qc0[RPT], qc1[RPT], qc2[RPT], ac1[RPT];
            {
               const Real *sm = (const Real *)stage(gm) + soff;
               const Real *sc = (const Real *)stage(gc) + soff;
#pragma unroll
               for (int r = 0; r < RPT; r++) {
                  qm0[r] = f4_mul(f4_ld(sm + (r - 1) * BZ), A2, NZ);
                  qm1[r] = f4_mul(f4_ld(sm + r * BZ), A2, NZ);
                  qm2[r] = f4_mul(f4_ld(sm + (r + 1) * BZ), A2, NZ);
                  qc0[r] = f4_mul(f4_ld(sc + (r - 1) * BZ), A2, NZ);
                  const F4 t = f4_ld(sc + r * BZ);
                  qc1[r] = f4_mul(t, A2, NZ), ac1[r] = f4_mul(t, A1, NZ);
                  qc2[r] = f4_mul(f4_ld(sc + (r + 1) * BZ), A2, NZ);
               }
            }
            Real *u0p = u0g + ((i64)sg.xa * Ny + ybase) * Nzp + zv;
            for (int j = 0; j < sg.cnt; j++) {
               const int x = sg.xa + j;
               const unsigned xrole = !FFUSE ? 0u
                                             : ((((eg.x_lo && x == 1) || (eg.x_hi && x == eg.Nx - 2)) ? 1u : 0u) | ((eg.x_lo && x == 2) ? 2u : 0u) |
                                                ((eg.x_hi && x == eg.Nx - 3) ? 4u : 0u));
               Ring gu = gc;  // plane x+1
               gu.next();
               wait_full(gu);
               wait_patched(gc);
               const unsigned char *stc = stage(gc);
               const Real *sm = (const Real *)stage(gm) + soff;
               const Real *sc = (const Real *)stc + soff;
               const Real *su = (const Real *)stage(gu) + soff;
#pragma unroll
               for (int r = 0; r < RPT; r++) {
                  const F4 t1 = f4_ld(su + r * BZ);
                  const F4 qp0 = f4_mul(f4_ld(su + (r - 1) * BZ), A2, NZ), qp1 = f4_mul(t1, A2, NZ),
                           qp2 = f4_mul(f4_ld(su + (r + 1) * BZ), A2, NZ), an1 = f4_mul(t1, A1, NZ);
                  // products of the out-of-row z neighbours (edge lanes; the others get theirs by shuffle)
                  Real em1 = (Real)0, ec0 = (Real)0, ec2 = (Real)0, ep1 = (Real)0;
                  if (edge_lane) {
                     em1 = sm[r * BZ + eoff];
                     ec0 = sc[(r - 1) * BZ + eoff];
                     ec2 = sc[(r + 1) * BZ + eoff];
                     ep1 = su[r * BZ + eoff];
                  }
                  em1 = O::mul(a2, em1), ec0 = O::mul(a2, ec0), ec2 = O::mul(a2, ec2), ep1 = O::mul(a2, ep1);
                  Real m1[VEC], c0[VEC], c2[VEC], p1[VEC];
                  f4_get(qm1[r], m1);
                  f4_get(qc0[r], c0);
                  f4_get(qc2[r], c2);
                  f4_get(qp1, p1);
                  const Real m1l = zleft(m1, em1), m1r = zright(m1, em1);
                  const Real c0l = zleft(c0, ec0), c0r = zright(c0, ec0);
                  const Real c2l = zleft(c2, ec2), c2r = zright(c2, ec2);
                  const Real p1l = zleft(p1, ep1), p1r = zright(p1, ep1);
                  Real u0v[VEC];
                  ld_vec<Real, VEC>((const Real *)(stc + u0off) + r * TZ, u0v);
                  const uint32_t m = (*(const uint32_t *)(stc + mkoff + r * C::MKW * 4) >> mshift) & VMASK;
                  F4 P = f4_sub(ac1[r], F4{pk2((float)u0v[0], (float)u0v[1]), pk2((float)u0v[2], (float)u0v[3])});
                  P = f4_add(P, qp2);     // +x +y
                  P = f4_add(P, qm0[r]);  // -x -y
                  Real pe[VEC];
                  f4_get(P, pe);
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     Real p = pe[k];
                     p = O::add(p, (k < VEC - 1) ? c2[k + 1 < VEC ? k + 1 : k] : c2r);  // +y +z
                     p = O::add(p, (k > 0) ? c0[k > 0 ? k - 1 : 0] : c0l);              // -y -z
                     p = O::add(p, (k < VEC - 1) ? p1[k + 1 < VEC ? k + 1 : k] : p1r);  // +x +z
                     p = O::add(p, (k > 0) ? m1[k > 0 ? k - 1 : 0] : m1l);              // -x -z
                     pe[k] = p;
                  }
                  P = F4{pk2((float)pe[0], (float)pe[1]), pk2((float)pe[2], (float)pe[3])};
                  P = f4_add(P, qp0);     // +x -y
                  P = f4_add(P, qm2[r]);  // -x +y
                  f4_get(P, pe);
                  Real o[VEC];
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     Real p = pe[k];
                     p = O::add(p, (k > 0) ? c2[k > 0 ? k - 1 : 0] : c2l);              // +y -z
                     p = O::add(p, (k < VEC - 1) ? c0[k + 1 < VEC ? k + 1 : k] : c0r);  // -y +z
                     p = O::add(p, (k > 0) ? p1[k > 0 ? k - 1 : 0] : p1l);              // +x -z
                     p = O::add(p, (k < VEC - 1) ? m1[k + 1 < VEC ? k + 1 : k] : m1r);  // -x +z
                     o[k] = ((m >> k) & 1u) ? u0v[k] : p;
                  }
                  if constexpr (FFUSE) {
                     if (r < nrow) {
                        const unsigned am = __activemask();
                        emit(r, x, xrole, am, o, u0v, u0p + (i64)r * Nzp, nullptr);
                     }
                  } else {
                     if (r < nrow && (SVC || m != VMASK)) st_vec<Real, VEC>(u0p + (i64)r * Nzp, o);
                  }
                  qm0[r] = qc0[r], qm1[r] = qc1[r], qm2[r] = qc2[r];
                  qc0[r] = qp0, qc1[r] = qp1, qc2[r] = qp2, ac1[r] = an1;
               }
               release(gm);
               gm = gc;
               gc = gu;
               u0p += jb.plane;
            }
            release(gm);  // the two planes still held: the last centre plane and the last "x+1" plane
            release(gc);
            g0 = gc;
            g0.next();
            continue;
         }
         Ring gm = g0;  // plane x-1
         Ring gc = g0;  // plane x
         gc.next();
         wait_full(gc);
         Real m0[RPT][VEC], m1[RPT][VEC], m2[RPT][VEC], c0[RPT][VEC], c1[RPT][VEC], c2[RPT][VEC];
         {
            const Real *sm = (const Real *)stage(gm) + soff;
            const Real *sc = (const Real *)stage(gc) + soff;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
               ld_vec<Real, VEC>(sm + (r - 1) * BZ, m0[r]);
               ld_vec<Real, VEC>(sm + r * BZ, m1[r]);
               ld_vec<Real, VEC>(sm + (r + 1) * BZ, m2[r]);
               ld_vec<Real, VEC>(sc + (r - 1) * BZ, c0[r]);
               ld_vec<Real, VEC>(sc + r * BZ, c1[r]);
               ld_vec<Real, VEC>(sc + (r + 1) * BZ, c2[r]);
            }
         }
         Real *u0p = u0g + ((i64)sg.xa * Ny + ybase) * Nzp + zv;
         for (int j = 0; j < sg.cnt; j++) {
            const int x = sg.xa + j;
            const unsigned xrole = !fuse ? 0u
                                      : ((((eg.x_lo && x == 1) || (eg.x_hi && x == eg.Nx - 2)) ? 1u : 0u) | ((eg.x_lo && x == 2) ? 2u : 0u) |
                                      ((eg.x_hi && x == eg.Nx - 3) ? 4u : 0u));
            Ring gu = gc;  // plane x+1
            gu.next();
            wait_full(gu);
            wait_patched(gc);
            const unsigned char *stc = stage(gc);
            const Real *sm = (const Real *)stage(gm) + soff;
            const Real *sc = (const Real *)stc + soff;
            const Real *su = (const Real *)stage(gu) + soff;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
               Real p0[VEC], p1[VEC], p2[VEC];
               ld_vec<Real, VEC>(su + (r - 1) * BZ, p0);
               ld_vec<Real, VEC>(su + r * BZ, p1);
               ld_vec<Real, VEC>(su + (r + 1) * BZ, p2);
               {
                  // every lane of the warp takes part (shuffles); rows beyond the grid compute on zero-filled stages
                  Real em1 = (Real)0, ec0 = (Real)0, ec2 = (Real)0, ep1 = (Real)0;
                  if (edge_lane) {
                     em1 = sm[r * BZ + eoff];
                     ec0 = sc[(r - 1) * BZ + eoff];
                     ec2 = sc[(r + 1) * BZ + eoff];
                     ep1 = su[r * BZ + eoff];
                  }
                  const Real m1l = zleft(m1[r], em1), m1r = zright(m1[r], em1);
                  const Real c0l = zleft(c0[r], ec0), c0r = zright(c0[r], ec0);
                  const Real c2l = zleft(c2[r], ec2), c2r = zright(c2[r], ec2);
                  const Real p1l = zleft(p1, ep1), p1r = zright(p1, ep1);
                  Real u0v[VEC];
                  ld_vec<Real, VEC>((const Real *)(stc + u0off) + r * TZ, u0v);
                  const uint32_t m = (*(const uint32_t *)(stc + mkoff + r * C::MKW * 4) >> mshift) & VMASK;
                  Real o[VEC];
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     Real p = O::sub(O::mul(a1, c1[r][k]), u0v[k]);
                     p = O::add(p, O::mul(a2, p2[k]));                                     // +x +y
                     p = O::add(p, O::mul(a2, m0[r][k]));                                  // -x -y
                     p = O::add(p, O::mul(a2, (k < VEC - 1) ? c2[r][k + 1] : c2r));        // +y +z
                     p = O::add(p, O::mul(a2, (k > 0) ? c0[r][k - 1] : c0l));              // -y -z
                     p = O::add(p, O::mul(a2, (k < VEC - 1) ? p1[k + 1] : p1r));           // +x +z
                     p = O::add(p, O::mul(a2, (k > 0) ? m1[r][k - 1] : m1l));              // -x -z
                     p = O::add(p, O::mul(a2, p0[k]));                                     // +x -y
                     p = O::add(p, O::mul(a2, m2[r][k]));                                  // -x +y
                     p = O::add(p, O::mul(a2, (k > 0) ? c2[r][k - 1] : c2l));              // +y -z
                     p = O::add(p, O::mul(a2, (k < VEC - 1) ? c0[r][k + 1] : c0r));        // -y +z
                     p = O::add(p, O::mul(a2, (k > 0) ? p1[k - 1] : p1l));                 // +x -z
                     p = O::add(p, O::mul(a2, (k < VEC - 1) ? m1[r][k + 1] : m1r));        // -x +z
                     o[k] = ((m >> k) & 1u) ? u0v[k] : p;
                  }
                  // (also a fully masked vector is stored: the service warp may have finished one of its nodes in the stage)
                  if constexpr (FFUSE) {
                     if (r < nrow) {
                        const unsigned am = __activemask();
                        emit(r, x, xrole, am, o, u0v, u0p + (i64)r * Nzp, nullptr);
                     }
                  } else {
                     if (r < nrow && (SVC || m != VMASK)) st_vec<Real, VEC>(u0p + (i64)r * Nzp, o);
                  }
               }
#pragma unroll
               for (int k = 0; k < VEC; k++) {
                  m0[r][k] = c0[r][k], m1[r][k] = c1[r][k], m2[r][k] = c2[r][k];
                  c0[r][k] = p0[k], c1[r][k] = p1[k], c2[r][k] = p2[k];
               }
            }
            release(gm);
            gm = gc;
            gc = gu;
            u0p += jb.plane;
         }
         release(gm);  // the two planes still held: the last centre plane and the last "x+1" plane
         release(gc);
         g0 = gc;
         g0.next();
         continue;
      }
      // PK (fp32): the strip's columns are kept as PRODUCTS, made once when a plane arrives and reused while it is x+1, x and x-1:
      // qp/qc/qm = a2*u1 of those planes, ac = a1*u1 of plane x.  Otherwise (fp64) um/uc/up hold the raw values.
      constexpr bool PK = sizeof(Real) == 4;
      Real um[RPT][VEC], uc[RPT][VEC], up[RPT][VEC];
      F4 qm[RPT], qc[RPT], qp[RPT], ac[RPT];
      u64 A1 = 0, A2 = 0, NZ = 0;
      if constexpr (PK) A1 = pk2((float)a1, (float)a1), A2 = pk2((float)a2, (float)a2), NZ = pk2(eg.negzero, eg.negzero);
      Ring gc = g0;  // plane xa-1 (already waited for)
      {
         const Real *s0 = (const Real *)stage(gc) + soff;
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if constexpr (PK) qm[r] = f4_mul(f4_ld(s0 + r * BZ), A2, NZ);
            else ld_vec<Real, VEC>(s0 + r * BZ, um[r]);
         }
      }
      release(gc);
      gc.next();  // plane xa: the first centre plane
      wait_full(gc);
      {
         const Real *s1 = (const Real *)stage(gc) + soff;
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if constexpr (PK) {
               const F4 t = f4_ld(s1 + r * BZ);
               qc[r] = f4_mul(t, A2, NZ), ac[r] = f4_mul(t, A1, NZ);
            } else {
               ld_vec<Real, VEC>(s1 + r * BZ, uc[r]);
            }
         }
      }
      Real *u0p = u0g + ((i64)sg.xa * Ny + ybase) * Nzp + zv;
      Real *zop = eg.zold + ((i64)sg.xa * Ny + ybase) * 2;  // stash of the z-shell values of this strip's rows

      for (int j = 0; j < sg.cnt; j++) {
         const int x = sg.xa + j;
         // bit0 plane on the x shell, bit1 x==2, bit2 x==Nx-3 (mirror sources), only at the global x ends
         const unsigned xrole = !fuse ? 0u
                                      : ((((eg.x_lo && x == 1) || (eg.x_hi && x == eg.Nx - 2)) ? 1u : 0u) | ((eg.x_lo && x == 2) ? 2u : 0u) |
                                         ((eg.x_hi && x == eg.Nx - 3) ? 4u : 0u));
         Ring gu = gc;
         gu.next();
         wait_full(gu);
         wait_patched(gc);
         const unsigned char *stc = stage(gc);
         const Real *sc = (const Real *)stc + soff;
         const Real *su = (const Real *)stage(gu) + soff;
         Real rowm[VEC], rowp[VEC], zl[RPT], zr[RPT];
         F4 an[RPT], qrm, qrp;  // PK: a1*u1 of plane x+1 (next centre), a2 * the rows next to the strip
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if constexpr (PK) {
               const F4 t = f4_ld(su + r * BZ);
               qp[r] = f4_mul(t, A2, NZ), an[r] = f4_mul(t, A1, NZ);
            } else {
               ld_vec<Real, VEC>(su + r * BZ, up[r]);
            }
         }
         if constexpr (PK) {
            qrm = f4_mul(f4_ld(sc - BZ), A2, NZ), qrp = f4_mul(f4_ld(sc + RPT * BZ), A2, NZ);
         } else {
            ld_vec<Real, VEC>(sc - BZ, rowm);
            ld_vec<Real, VEC>(sc + RPT * BZ, rowp);
         }
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            zl[r] = sc[r * BZ - 1];
            zr[r] = sc[r * BZ + VEC];
         }
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if (r < nrow) {
               // the lanes of this warp whose vector holds interior nodes; taken before any row-dependent branch (with
               // LZ < 32 the row groups of a warp can differ in their shell / mirror roles)
               const unsigned am = __activemask();
               Real u0v[VEC];
               ld_vec<Real, VEC>((const Real *)(stc + u0off) + r * TZ, u0v);
               const uint32_t m = (*(const uint32_t *)(stc + mkoff + r * C::MKW * 4) >> mshift) & VMASK;
               Real o[VEC];
               if constexpr (PK) {
                  // (a1*uc - u0) + a2*u[+x] + a2*u[-x] + a2*u[+y] + a2*u[-y] on two element pairs, then + a2*u[+z] + a2*u[-z]
                  // per element (the z neighbours sit one element over: pairing them would cost more moves than it saves)
                  F4 P = f4_sub(ac[r], F4{pk2((float)u0v[0], (float)u0v[1]), pk2((float)u0v[2], (float)u0v[3])});
                  P = f4_add(P, qp[r]);
                  P = f4_add(P, qm[r]);
                  P = f4_add(P, (r < RPT - 1) ? qc[r + 1 < RPT ? r + 1 : r] : qrp);
                  P = f4_add(P, (r > 0) ? qc[r > 0 ? r - 1 : 0] : qrm);
                  Real pe[VEC], q[VEC];
                  f4_get(P, pe);
                  f4_get(qc[r], q);
                  const Real qzl = O::mul(a2, zl[r]), qzr = O::mul(a2, zr[r]);
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     Real p = O::add(pe[k], (k < VEC - 1) ? q[k + 1 < VEC ? k + 1 : k] : qzr);
                     p = O::add(p, (k > 0) ? q[k > 0 ? k - 1 : 0] : qzl);
                     o[k] = ((m >> k) & 1u) ? u0v[k] : p;
                  }
               } else {
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     const Real yp = (r < RPT - 1) ? uc[r + 1][k] : rowp[k];
                     const Real ym = (r > 0) ? uc[r - 1][k] : rowm[k];
                     const Real zp = (k < VEC - 1) ? uc[r][k + 1] : zr[r];
                     const Real zm = (k > 0) ? uc[r][k - 1] : zl[r];
                     Real p = O::sub(O::mul(a1, uc[r][k]), u0v[k]);
                     p = O::add(p, O::mul(a2, up[r][k]));
                     p = O::add(p, O::mul(a2, um[r][k]));
                     p = O::add(p, O::mul(a2, yp));
                     p = O::add(p, O::mul(a2, ym));
                     p = O::add(p, O::mul(a2, zp));
                     p = O::add(p, O::mul(a2, zm));
                     o[k] = ((m >> k) & 1u) ? u0v[k] : p;
                  }
               }
               Real *dst = u0p + (i64)r * Nzp;
               const unsigned rrole = ((yrole >> (4 * r)) & 7u) | xrole;  // uniform within a row group
               // Fused extras.  Everything lane-dependent below is written as selects / single predicated stores:
               // a divergent branch here would make the one lane at a z end run the rest of the step on its own.
               const bool shell = (rrole & 1u) != 0;
               if (shell) {
                  // row / plane on the absorbing shell: keep the plain air value, stash the pre-update values for
                  // k_abc_faces (one extra vector store; the planes on the x shell take precedence)
                  Real *sp = (xrole & 1u) ? eg.xold + (((i64)(x == 1 ? 0 : 1) * Ny + (ybase + r)) * Nzp + zv)
                                          : eg.yold + ((((i64)x * 2 + ((ybase + r) == 1 ? 0 : 1)) * Nzp) + zv);
                  st_vec<Real, VEC>(sp, u0v);
               }
               // Mirror targets that live in ANOTHER active lane's vector are delivered by a shuffle (that lane stores
               // its whole vector; a scalar store from here would race with it); only a target in a vector nobody
               // stores (the row's far padding) is written directly.
               __syncwarp(am);  // the row groups may have diverged on `shell`
               if (zlo_tile) {  // warp-uniform: the tile starts at z = 0
                  // lane 0 holds z=1 (shell): stash its pre-update value; z=2 -> z=0 mirror
                  if (eg.zstash && lz == 0 && !shell) zop[2 * r] = u0v[1];
                  if constexpr (VEC >= 4) {
                     o[0] = (lz == 0) ? o[2] : o[0];
                  } else {
                     const Real t = __shfl_down_sync(am, o[0], 1);  // fp64: z=2 is the first element of the next lane
                     o[0] = (lz == 0) ? t : o[0];
                  }
               }
               Real vm = o[0];  // value of z = Nz-3 if this thread holds it
               if (zhi_tile) {  // warp-uniform: the tile contains z = Nz-3 .. Nz-1
                  Real vs = u0v[0];
#pragma unroll
                  for (int k = 1; k < VEC; k++) {
                     vs = (k == khs) ? u0v[k] : vs;
                     vm = (k == khm) ? o[k] : vm;
                  }
                  if (eg.zstash && khs >= 0 && !shell) zop[2 * r + 1] = vs;  // z = Nz-2 (shell)
#pragma unroll
                  for (int k = 0; k + 2 < VEC; k++) o[k + 2] = (k == khm) ? o[k] : o[k + 2];  // z=Nz-3 -> z=Nz-1 inside the vector
                  // a vector that starts at z=Nz-2 holds z=Nz-1 as element 1 and finds z=Nz-3 at the end of the previous lane's
                  // (never the first vector of a tile: the engine refuses the fused step for such grids, AirTma::z_edge)
                  __syncwarp(am);
                  const Real t = __shfl_up_sync(am, o[VEC - 1], 1);
                  o[1] = (khs == 0) ? t : o[1];
               }
               // Every vector of an active row is stored, fully masked ones too: it may hold the z halo (mirror-on-write: a masked
               // node at z = 1 / Nz-2 must not keep the halo next to it from being refreshed) or a node the service warp finished
               // in the stage.  Masked elements carry their stage value.
               // Where the z halo lands when its source z = Nz-3 is near the end of the vector:
               //   tail1: z = Nz-1 opens the next vector, which nobody stores;
               //   tail2: z = Nz-3 closes the vector, so z = Nz-1 is the SECOND element of the next one.  On grids with
               //          Nz = 2 (mod tile width) that vector opens the next TILE (`lone`), whose only node is the shell node z = Nz-2:
               //          its thread stores that one element alone, so that the halo written from here survives whichever tile runs first.
               // (ZE: a kernel variant of its own for those grids -- the extra live values cost the ordinary kernel 1-2 %)
               const bool tail1 = zhi_tile && khm + 2 == VEC, tail2 = ZE && zhi_tile && khm + 1 == VEC;
               const bool lone = ZE && fuse && sg.z0 == Nz - 2;  // warp-uniform
               if (lone) dst[0] = o[0];
               else st_vec<Real, VEC>(dst, o);
               if (tail1) dst[VEC] = vm;
               if (tail2) dst[VEC + 1] = vm;
               if (rrole & 6u) {
                  // mirror source of a y / x halo: the same row goes there as well (warp-uniform, a few rows / planes)
                  const int y = ybase + r;
#pragma unroll 1
                  for (int t = 1; t < 5; t++) {
                     const bool on = t == 1 ? y == 2 : t == 2 ? y == Ny - 3 : t == 3 ? (xrole & 2u) != 0 : (xrole & 4u) != 0;
                     if (on) {
                        Real *d = dst + (t == 1 ? -2 * (i64)Nzp : t == 2 ? 2 * (i64)Nzp : t == 3 ? -2 * jb.plane : 2 * jb.plane);
                        if (lone) d[0] = o[0];
                        else st_vec<Real, VEC>(d, o);
                        if (tail1) d[VEC] = vm;
                        if (tail2) d[VEC + 1] = vm;
                     }
                  }
               }
            }
         }
         release(gc);  // plane x's stage may be refilled; x-1 and x+1 live in registers / the next stage
         gc = gu;
         u0p += jb.plane;
         zop += 2 * Ny;
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if constexpr (PK) {
               qm[r] = qc[r], qc[r] = qp[r], ac[r] = an[r];
            } else {
#pragma unroll
               for (int k = 0; k < VEC; k++) {
                  um[r][k] = uc[r][k];
                  uc[r][k] = up[r][k];
               }
            }
         }
      }
      release(gc);  // the last plane was only ever an "x+1" plane
      g0 = gc;
      g0.next();
   }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tile configurations (id, rows per thread, consumer warps, stages, register cap, lanes along z, service warp); 0 / 5 are the
// defaults for whole-tile grids without in-kernel boundary work (7-point / FCC), 8..11 their narrow-tile versions for ragged Nz
// (see air_pick_cfg); 12..17 the same three widths with the service warp (one consumer warp fewer: 16 / 12 warps per CTA keep two
// CTAs per SM inside the register file)
#define PF_AIR_CONFIGS(X)        \
   X(0, 1, 15, 6, 64, 32, false)  \
   X(1, 2, 8, 4, 72, 32, false)   \
   X(2, 2, 8, 6, 112, 32, false)  \
   X(3, 1, 15, 4, 64, 32, false)  \
   X(4, 4, 8, 3, 112, 32, false)  \
   X(5, 1, 11, 6, 80, 32, false)  \
   X(6, 2, 12, 4, 72, 32, false)  \
   X(7, 1, 8, 6, 80, 32, false)   \
   X(8, 1, 15, 6, 64, 16, false)  \
   X(9, 1, 15, 6, 64, 8, false)   \
   X(10, 1, 11, 6, 80, 16, false) \
   X(11, 1, 11, 6, 80, 8, false)  \
   X(12, 1, 14, 6, 64, 32, true)  \
   X(13, 1, 14, 6, 64, 16, true)  \
   X(14, 1, 14, 6, 64, 8, true)   \
   X(15, 1, 11, 6, 72, 32, true)  \
   X(16, 1, 11, 6, 72, 16, true)  \
   X(17, 1, 11, 6, 72, 8, true)
#define PF_AIR_NCFG 18

template <typename Real>
static int air_tma_attr(int cfg) {
   cudaError_t rc = cudaErrorInvalidValue;
#define X(id, RPT, NW, S, MAXR, LZ, SVC)                                                                                          \
   if (cfg == id) {                                                                                                                   \
      rc = cudaFuncSetAttribute(k_air_tma_cart<Real, RPT, NW, S, MAXR, false, LZ, SVC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                AirCfg<Real, RPT, NW, S, LZ, SVC>::SMEM_BYTES);                                                       \
      if (rc == cudaSuccess)                                                                                                          \
         rc = cudaFuncSetAttribute(k_air_tma_cart<Real, RPT, NW, S, MAXR, true, LZ, SVC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   AirCfg<Real, RPT, NW, S, LZ, SVC>::SMEM_BYTES);                                                    \
      if (rc == cudaSuccess && SVC)                                                                                                   \
         rc = cudaFuncSetAttribute(k_air_tma_cart<Real, RPT, NW, S, MAXR, true, LZ, SVC, SVC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   AirCfg<Real, RPT, NW, S, LZ, SVC>::SMEM_BYTES);                                                    \
      if (rc == cudaSuccess)                                                                                                          \
         rc = cudaFuncSetAttribute(k_air_tma_cart<Real, RPT, NW, S, MAXR, false, LZ, SVC, false, true>,                               \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, AirCfg<Real, RPT, NW, S, LZ, SVC>::SMEM_BYTES);       \
   }
   PF_AIR_CONFIGS(X)
#undef X
   return (int)rc;
}

static void air_cfg_shape(int cfg, int *rpt, int *nw, int *lz, int *svc = nullptr) {
   *rpt = 4, *nw = 8, *lz = 32;
   int sv = 0;
#define X(id, RPT, NW, S, MAXR, LZ, SVC) \
   if (cfg == id) *rpt = RPT, *nw = NW, *lz = LZ, sv = SVC ? 1 : 0;
   PF_AIR_CONFIGS(X)
#undef X
   if (svc) *svc = sv;
}

// default configuration for a grid: the widest tile (32, 16, 8 lanes along z) that wastes less than 12 % of its columns on the
// padding behind Nz-1, else the one that wastes least.  Real rooms need this: the reference's gpu folders make z the SHORTEST
// axis (CTK church Nz = 180: 70 % useful columns with 32 lanes, 93 % with 16; Musikverein Nz = 258: 67 % / 80 % / 89 %).
static int air_pick_cfg(int fcc, int precision, i64 Nz, int svc = 0) {
   const int vec = precision == 1 ? 4 : 2;
   const int ids[3] = {svc ? (fcc ? 15 : 12) : (fcc ? 5 : 0), svc ? (fcc ? 16 : 13) : (fcc ? 10 : 8), svc ? (fcc ? 17 : 14) : (fcc ? 11 : 9)};
   const int lzs[3] = {32, 16, 8};
   // A width whose tiles make the shell node z = Nz-2 open a tile (AirTma::z_edge) cannot run the fused step: prefer a wider one
   // that does not, as long as it still fills half of its columns.  (Nz-2 a multiple of the widest tile defeats all three.)
   int best = -1;
   double best_eff = 0;
   for (int pass = 0; pass < 2 && best < 0; pass++) {
      for (int k = 0; k < 3; k++) {
         const i64 tz = (i64)lzs[k] * vec, cols = (Nz - 1 + tz - 1) / tz * tz;
         const double eff = (double)(Nz - 1) / (double)cols;
         const bool edge = (Nz - 2) % tz == 0;
         if (pass == 0 && (edge || eff < 0.5)) continue;
         if (eff >= 0.88) return ids[k];
         if (eff > best_eff + 1e-9) best_eff = eff, best = k;
      }
   }
   return ids[best < 0 ? 0 : best];
}

static int air_tma_setup(AirTma *t, int precision, int fcc, i64 Nx, i64 Ny, i64 Nz, i64 Nzp, i64 mwpr, void *u_a, void *u_b, void *mask,
                         int cfg = -1, int svc = 0) {
   // defaults (measured on B200, profiles/): 7-point: 15 consumer warps, 64 registers; 13-point FCC: its nine rotating
   // row vectors need 80 registers to stay out of local memory -> 11 consumer warps (c3s: 264 us vs 294 us with cfg 0)
   if (cfg < 0) cfg = air_pick_cfg(fcc, precision, Nz, svc);
   t->ok = false;
   t->sv = AirSvc{nullptr, nullptr, 0};  // lists belong to a tile shape: the engine rebuilds them after every set-up
   t->precision = precision, t->fcc = fcc, t->Nx = Nx, t->Ny = Ny, t->Nz = Nz, t->Nzp = Nzp, t->mwpr = mwpr;
   t->base[0] = u_a, t->base[1] = u_b, t->mask = mask;
   if (cfg < 0 || cfg >= PF_AIR_NCFG) {
      t->why = "no such tile configuration";
      return 1;
   }
   t->cfg = cfg;
   t->slots = 0;
   if (Nx > 0x7fffffff || Ny > 0x7fffffff || Nzp > 0x7fffffff) {
      t->why = "grid dimension exceeds 2^31";
      return 1;
   }
   void *fn = nullptr;
   cudaDriverEntryPointQueryResult qres;
   if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
       qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      t->why = "cuTensorMapEncodeTiled not available from the driver";
      return 1;
   }
   EncodeTiledFn encode = (EncodeTiledFn)fn;
   const size_t rs = precision == 1 ? 4 : 8;
   const int VEC = 16 / (int)rs;
   int rpt, nw, lz;
   air_cfg_shape(cfg, &rpt, &nw, &lz, &t->svc);
   const int ty = nw * rpt * (32 / lz);
   t->ty = ty, t->tzn = lz * VEC;
   // mirror-on-write finds the source z = Nz-3 of the halo z = Nz-1 in the same z tile; when the shell node z = Nz-2 opens a tile
   // the source sits in another CTA's tile and the fused step must not be used (tests cart_nz_e / cart_nz_f)
   t->z_edge = ((Nz - 2) % ((i64)lz * VEC) == 0) ? 1 : 0;
   const CUtensorMapDataType dt = precision == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
   const cuuint64_t gdim[3] = {(cuuint64_t)Nzp, (cuuint64_t)Ny, (cuuint64_t)Nx};
   const cuuint64_t gstr[2] = {(cuuint64_t)Nzp * rs, (cuuint64_t)Ny * Nzp * rs};
   const cuuint32_t box1[3] = {(cuuint32_t)(lz * VEC + 2 * VEC), (cuuint32_t)(ty + 2), 1};
   const cuuint32_t box0[3] = {(cuuint32_t)(lz * VEC), (cuuint32_t)ty, 1};
   const cuuint32_t estr[3] = {1, 1, 1};
   CUresult r = CUDA_SUCCESS;
   for (int k = 0; k < 2 && r == CUDA_SUCCESS; k++) {
      r = encode(&t->map_u1[k], dt, 3, t->base[k], gdim, gstr, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS)
         r = encode(&t->map_u0[k], dt, 3, t->base[k], gdim, gstr, box0, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   }
   if (r == CUDA_SUCCESS) {
      const cuuint64_t mdim[3] = {(cuuint64_t)mwpr, (cuuint64_t)Ny, (cuuint64_t)Nx};
      const cuuint64_t mstr[2] = {(cuuint64_t)mwpr * 4, (cuuint64_t)Ny * mwpr * 4};
      const cuuint32_t mbox[3] = {4, (cuuint32_t)ty, 1};
      r = encode(&t->map_mk, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, mask, mdim, mstr, mbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   }
   if (r != CUDA_SUCCESS) {
      t->why = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
      return 1;
   }
   int rc = precision == 1 ? air_tma_attr<float>(cfg) : air_tma_attr<double>(cfg);
   if (rc) {
      cudaGetLastError();
      t->why = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString((cudaError_t)rc);
      return 1;
   }
   int dev = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
   t->ok = true;
   t->why = "";
   return 0;
}

template <typename Real, int RPT, int NW, int S, int MAXR, int LZ, bool SVC>
static int air_tma_launch_cfg(AirTma *t, int cur, Real *u0, i64 xb, i64 xe, Real a1, Real a2, const AirEdge<Real> &eg, bool use_lists,
                              cudaStream_t s, int *ctr) {
   typedef AirCfg<Real, RPT, NW, S, LZ, SVC> C;
   // (the fused 13-point step is its own kernel, compiled for the service-warp configurations only: FFUSE = SVC there)
   auto kern = t->fcc ? ((eg.fuse && SVC) ? k_air_tma_cart<Real, RPT, NW, S, MAXR, true, LZ, SVC, SVC>
                                          : k_air_tma_cart<Real, RPT, NW, S, MAXR, true, LZ, SVC, false>)
                      : (t->z_edge ? k_air_tma_cart<Real, RPT, NW, S, MAXR, false, LZ, SVC, false, true>
                                   : k_air_tma_cart<Real, RPT, NW, S, MAXR, false, LZ, SVC, false, false>);
   if (t->slots <= 0) {
      int per_sm = 0;
      cudaError_t rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, C::SMEM_BYTES);
      if (rc != cudaSuccess) return (int)rc;
      t->slots = std::max(1, per_sm) * t->sm_count;
   }
   AirJob jb;
   jb.x_begin = (int)xb;
   jb.n = (int)(xe - xb);
   jb.tz = (int)((t->Nz - 1 + C::TZ - 1) / C::TZ);  // vectors starting at z >= Nz-1 hold no interior node
   const int ty = (int)((t->Ny - 2 + C::TY - 1) / C::TY);
   jb.tiles = jb.tz * ty;
   jb.Ny = (int)t->Ny, jb.Nz = (int)t->Nz, jb.Nzp = (int)t->Nzp;
   jb.plane = t->Ny * t->Nzp;
   jb.ctr = ctr ? ctr : t->ctr;  // (launches that may run side by side need their own work counters)
   // x-chunk length: 16 planes.  Longer chunks save the two extra u1 planes an item loads (4 % of the traffic at 16) but let the
   // CTAs drift apart along x, and neighbouring tiles stop finding each other's halo rows and columns in L2: measured on B200,
   // c5 (2046 planes): 16 planes 0.975 of the copy peak, 32: 0.963, 60: 0.945; c4 (fp64 1024^3): 16: 0.976, 60: 0.946.
   {
      int xc = t->xc > 0 ? t->xc : 16;
      if (SVC) xc = std::min(xc, 60);  // the service warp keeps an item's cnt+1 list offsets in two registers per lane
      air_plan_chunks(&jb, jb.n, xc);
   }
   const i64 items = (i64)jb.nch * jb.tiles;
   if (items > 0x7fffffff) return (int)cudaErrorInvalidValue;
   jb.n_items = (int)items;
   const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(t->slots, items));
   kern<<<grid, C::THREADS, C::SMEM_BYTES, s>>>(t->map_u1[cur], t->map_u0[cur ^ 1], t->map_mk, u0, jb, a1, a2, eg,
                                                use_lists ? t->sv : AirSvc{nullptr, nullptr, 0});
   return (int)cudaGetLastError();
}

// planes [xb, xe) of the slab; `cur` = index of the grid that currently is u1 (u0 = the other one)
template <typename Real>
static int air_tma_launch(AirTma *t, int cur, Real *u0, i64 xb, i64 xe, Real a1, Real a2, const AirEdge<Real> &eg, bool use_lists,
                          cudaStream_t s, int *ctr = nullptr) {
#define X(id, RPT, NW, S, MAXR, LZ, SVC) \
   if (t->cfg == id) return air_tma_launch_cfg<Real, RPT, NW, S, MAXR, LZ, SVC>(t, cur, u0, xb, xe, a1, a2, eg, use_lists, s, ctr);
   PF_AIR_CONFIGS(X)
#undef X
   return (int)cudaErrorInvalidValue;
}

}  // namespace pf
