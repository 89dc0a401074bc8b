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
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace pf {

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
   int xc = 0;          // planes per x-chunk of the work order, 0 = automatic
   int sm_count = 148;
   int slots = 0;       // resident CTAs of the chosen configuration on this device
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

template <typename Real, int RPT, int NW, int S>
struct AirCfg {
   static constexpr int VEC = 16 / (int)sizeof(Real);
   static constexpr int TZ = 32 * VEC;
   static constexpr int TY = NW * RPT;
   static constexpr int BZ = TZ + 2 * VEC;  // u1 box starts one vector left of the tile: 16-byte aligned columns
   static constexpr int ROWS = TY + 2;
   static constexpr int MKW = 4;            // mask words per tile row in a stage (TZ/32 used, 16-byte TMA minimum)
   static constexpr int U1_BYTES = ROWS * BZ * (int)sizeof(Real);
   static constexpr int U0_BYTES = TY * TZ * (int)sizeof(Real);
   static constexpr int MK_BYTES = TY * MKW * 4;
   static constexpr int U0_OFF = (U1_BYTES + 127) / 128 * 128;
   static constexpr int MK_OFF = U0_OFF + U0_BYTES;
   static constexpr int STAGE_PITCH = (MK_OFF + MK_BYTES + 127) / 128 * 128;
   static constexpr int SMEM_BYTES = S * STAGE_PITCH + 2 * S * 8 + 128;
   static constexpr int THREADS = (NW + 1) * 32;  // NW consumer warps + 1 TMA producer warp
};

// ---------------------------------------------------------------- work decomposition
// The job "planes [x_begin, x_begin+n) x all y-z tiles" is cut into units of one tile-plane, ordered
// (x-chunk of XC planes, tile, plane in chunk), and every CTA of a persistent grid (one CTA per
// resident slot) takes an equal contiguous range of units: all CTAs finish together (no tail), and CTAs
// running at the same time work on the same x-chunk of neighbouring tiles, so tile halos hit in L2.
// A CTA's range is a few "segments" (one tile, consecutive planes); each costs two extra u1 plane loads.
struct AirJob {
   int x_begin, n, XC, tz, tiles;  // n planes, tiles = tz*ty
   int units;                      // n * tiles
   int Ny, Nz, Nzp;
   i64 plane;  // Ny*Nzp
};
struct AirSeg {
   int xa, cnt, z0, y0, next;  // next = first unit after the segment
};
template <int TZ, int TY>
__device__ __forceinline__ AirSeg air_segment(const AirJob &jb, int u, int u_end) {
   const int per_chunk = jb.XC * jb.tiles;
   const int k = u / per_chunk;
   const int len = min(jb.XC, jb.n - k * jb.XC);
   const int up = u - k * per_chunk;
   const int t = up / len, p = up - t * len;
   AirSeg s;
   s.next = min(u_end, k * per_chunk + (t + 1) * len);
   s.cnt = s.next - u;
   s.xa = jb.x_begin + k * jb.XC + p;
   s.z0 = (t % jb.tz) * TZ;
   s.y0 = 1 + (t / jb.tz) * TY;
   return s;
}

// Fused extras of the Cartesian step (all optional, `fuse` = 0 gives the plain masked air update):
//  * the absorbing shell (cpu_engine.h:225-229): a node with Q = #{axes on which its index is 1 or
//    N-2} > 0 becomes (v + lQ*u0_old)/(1.0 + lQ), v = the fresh air value (or the untouched old value
//    of a masked node, exactly what the reference's ABC loop sees); lQ = l*Q in Real and
//    den = 1.0 + lQ in DOUBLE are precomputed on the host with the reference's expressions;
//  * the halo mirrors (cpu_engine.h:145-172) are applied when a value is WRITTEN instead of before it is
//    read: whoever stores index 2 (N-3) of an axis also stores it to index 0 (N-1).  The new state then
//    leaves the kernel with its face halos complete and the next step needs no mirror pass.
template <typename Real>
struct AirEdge {
   int fuse, x_lo, x_hi, Nx;
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

// ---------------------------------------------------------------- the kernel (7-point Cartesian)
// full[s] flips when all bytes of a plane have landed in stage s, empty[s] when all NW consumer warps
// are done with it.  Loads are numbered consecutively over all segments of the CTA; load i uses stage
// i % S.  A segment of cnt planes loads planes xa-1 .. xa+cnt: the first and the last only as u1.
template <typename Real, int RPT, int NW, int S, int MAXR>
__global__ void __maxnreg__(MAXR)
    k_air_tma_cart(const __grid_constant__ CUtensorMap map_u1, const __grid_constant__ CUtensorMap map_u0,
                   const __grid_constant__ CUtensorMap map_mk, Real *__restrict__ u0g, const AirJob jb, const Real a1, const Real a2,
                   const AirEdge<Real> eg) {
   typedef AirCfg<Real, RPT, NW, S> C;
   typedef Ops<Real> O;
   constexpr int VEC = C::VEC, BZ = C::BZ, TZ = C::TZ;
   constexpr uint32_t VMASK = (1u << VEC) - 1u;
   extern __shared__ unsigned char smem_raw[];
   unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
   uint64_t *full = (uint64_t *)(smem + S * C::STAGE_PITCH);
   uint64_t *empty = full + S;

   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const int u_begin = (int)((i64)jb.units * blockIdx.x / gridDim.x), u_end = (int)((i64)jb.units * (blockIdx.x + 1) / gridDim.x);
   if (u_end <= u_begin) return;

   if (tid == 0) {
      for (int s = 0; s < S; s++) {
         mbar_init(&full[s], 1);
         mbar_init(&empty[s], NW);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
   }
   __syncthreads();

   if (w == NW) {
      // ---------------- producer
      if (lane == 0) {
         int i = 0;
         for (int u = u_begin; u < u_end;) {
            const AirSeg sg = air_segment<C::TZ, C::TY>(jb, u, u_end);
            for (int q = 0; q < sg.cnt + 2; q++, i++) {  // planes xa-1 .. xa+cnt
               const int s = i % S;
               unsigned char *st = smem + s * C::STAGE_PITCH;
               const bool centre = q >= 1 && q <= sg.cnt;
               if (i >= S) mbar_wait(&empty[s], (uint32_t)(((i / S) - 1) & 1));
               mbar_expect_tx(&full[s], centre ? C::U1_BYTES + C::U0_BYTES + C::MK_BYTES : C::U1_BYTES);
               const int x = sg.xa - 1 + q;
               tma_load_3d(st, &map_u1, &full[s], sg.z0 - VEC, sg.y0 - 1, x);
               if (centre) {
                  tma_load_3d(st + C::U0_OFF, &map_u0, &full[s], sg.z0, sg.y0, x);
                  tma_load_3d(st + C::MK_OFF, &map_mk, &full[s], (sg.z0 >> 7) << 2, sg.y0, x);  // box start must be 16-byte aligned
               }
            }
            u = sg.next;
         }
      }
      return;
   }

   // ---------------- consumers
   auto stage = [&](int i) -> const unsigned char * { return smem + (i % S) * C::STAGE_PITCH; };
   auto wait_full = [&](int i) { mbar_wait(&full[i % S], (uint32_t)((i / S) & 1)); };
   auto release = [&](int i) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[i % S]);
   };
   const int soff = (w * RPT + 1) * BZ + VEC + VEC * lane;      // strip row 0 inside the u1 box (box row 0 is y0-1)
   const int u0off = C::U0_OFF + ((w * RPT) * TZ + VEC * lane) * (int)sizeof(Real);
   const int mshift = (VEC * lane) & 31;
   const int Ny = jb.Ny, Nz = jb.Nz, Nzp = jb.Nzp;

   int base = 0;  // load index of the segment's first plane
   for (int u = u_begin; u < u_end;) {
      const AirSeg sg = air_segment<C::TZ, C::TY>(jb, u, u_end);
      u = sg.next;
      // this thread's strip: rows y0 + w*RPT + r, columns z0 + VEC*lane .. +VEC-1
      const int zv = sg.z0 + VEC * lane;
      const int ybase = sg.y0 + w * RPT;
      const int mkoff = C::MK_OFF + ((w * RPT) * C::MKW + ((zv & 127) >> 5)) * 4;  // the stage holds the words of z0 & ~127 ..
      int nrow = 0;  // active rows of the strip; vectors entirely in the far halo/padding are never touched
#pragma unroll
      for (int r = 0; r < RPT; r++) nrow += (zv < Nz - 1 && (ybase + r) <= Ny - 2) ? 1 : 0;
      // rows / lanes that need the slow path of the fused extras (details are recomputed there)
      const bool fuse = eg.fuse != 0;
      const bool zthread = fuse && (zv <= 2 || zv + VEC > Nz - 3);  // holds a z-shell node or a z-mirror source
      unsigned yspec = 0;
#pragma unroll
      for (int r = 0; r < RPT; r++) {
         const int y = ybase + r;
         yspec |= (fuse && (y <= 2 || y >= Ny - 3)) ? (1u << r) : 0u;
      }

      Real um[RPT][VEC], uc[RPT][VEC], up[RPT][VEC];
      wait_full(base);
      {
         const Real *s0 = (const Real *)stage(base) + soff;
#pragma unroll
         for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(s0 + r * BZ, um[r]);
      }
      release(base);
      wait_full(base + 1);
      {
         const Real *s1 = (const Real *)stage(base + 1) + soff;
#pragma unroll
         for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(s1 + r * BZ, uc[r]);
      }
      Real *u0p = u0g + ((i64)sg.xa * Ny + ybase) * Nzp + zv;

      for (int j = 0; j < sg.cnt; j++) {
         const int x = sg.xa + j;
         const bool xspec = fuse && ((eg.x_lo && x <= 2) || (eg.x_hi && x >= eg.Nx - 3));
         wait_full(base + j + 2);
         const unsigned char *stc = stage(base + j + 1);
         const Real *sc = (const Real *)stc + soff;
         const Real *su = (const Real *)stage(base + j + 2) + soff;
         Real rowm[VEC], rowp[VEC], zl[RPT], zr[RPT];
#pragma unroll
         for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(su + r * BZ, up[r]);
         ld_vec<Real, VEC>(sc - BZ, rowm);
         ld_vec<Real, VEC>(sc + RPT * BZ, rowp);
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            zl[r] = sc[r * BZ - 1];
            zr[r] = sc[r * BZ + VEC];
         }
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if (r < nrow) {
               Real u0v[VEC];
               ld_vec<Real, VEC>((const Real *)(stc + u0off) + r * TZ, u0v);
               const uint32_t m = (*(const uint32_t *)(stc + mkoff + r * C::MKW * 4) >> mshift) & VMASK;
               Real o[VEC];
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
               Real *dst = u0p + (i64)r * Nzp;
               if (!(zthread || xspec || ((yspec >> r) & 1u))) {
                  if (m != VMASK) st_vec<Real, VEC>(dst, o);  // the common case
               } else if (!(xspec || ((yspec >> r) & 1u))) {
                  // only the z ends are special: this thread holds z=1 / z=Nz-2 (shell, Q = 1) and/or
                  // z=2 / z=Nz-3 (mirror sources).  Kept lean: one lane of every edge warp comes through here.
                  const int kl = 2 - zv, kh = Nz - 3 - zv;
                  bool keep = m != VMASK;
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     if (zv + k == 1 || zv + k == Nz - 2) {
                        o[k] = abc_apply<Real>(o[k], u0v[k], eg.lQ1, eg.den1, eg.rden1);
                        keep = true;
                     }
                  }
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     if (k >= 2 && k == kl) o[k - 2] = o[k];
                     if (k + 2 < VEC && k == kh) o[k + 2] = o[k];
                  }
                  if (keep) {
                     st_vec<Real, VEC>(dst, o);
#pragma unroll
                     for (int k = 0; k < VEC; k++) {
                        if (k < 2 && k == kl) dst[k - 2] = o[k];
                        if (k + 2 >= VEC && k == kh) dst[k + 2] = o[k];
                     }
                  }
               } else {
                  // rows / planes on the shell or next to a y / x halo (a vanishing share of the grid)
                  const int y = ybase + r;
                  const int kl = 2 - zv, kh = Nz - 3 - zv;
                  const int qrow = (((eg.x_lo && x == 1) || (eg.x_hi && x == eg.Nx - 2)) ? 1 : 0) + ((y == 1 || y == Ny - 2) ? 1 : 0);
                  const bool xmlo = eg.x_lo && x == 2, xmhi = eg.x_hi && x == eg.Nx - 3;
                  bool keep = m != VMASK;
#pragma unroll 1
                  for (int k = 0; k < VEC; k++) {
                     const int z = zv + k;
                     const int Q = qrow + ((z == 1 || z == Nz - 2) ? 1 : 0);
                     if (Q > 0 && z >= 1 && z <= Nz - 2) {
                        const Real lQ = Q == 1 ? eg.lQ1 : (Q == 2 ? eg.lQ2 : eg.lQ3);
                        const double den = Q == 1 ? eg.den1 : (Q == 2 ? eg.den2 : eg.den3);
                        const double rden = Q == 1 ? eg.rden1 : (Q == 2 ? eg.rden2 : eg.rden3);
                        Real v = o[0], old = u0v[0];
#pragma unroll
                        for (int kk = 1; kk < VEC; kk++)
                           if (kk == k) v = o[kk], old = u0v[kk];
                        v = abc_apply<Real>(v, old, lQ, den, rden);
#pragma unroll
                        for (int kk = 0; kk < VEC; kk++)
                           if (kk == k) o[kk] = v;
                        keep = true;
                     }
                  }
#pragma unroll
                  for (int k = 0; k < VEC; k++) {
                     if (k >= 2 && k == kl) o[k - 2] = o[k];
                     if (k + 2 < VEC && k == kh) o[k + 2] = o[k];
                  }
                  // the row itself, then the same row into the y / x halos it is the mirror source of
#pragma unroll 1
                  for (int t = 0; t < 5; t++) {
                     const bool on = t == 0 ? keep : t == 1 ? (y == 2) : t == 2 ? (y == Ny - 3) : t == 3 ? xmlo : xmhi;
                     if (on) {
                        Real *d = dst + (t == 1 ? -2 * (i64)Nzp : t == 2 ? 2 * (i64)Nzp : t == 3 ? -2 * jb.plane : t == 4 ? 2 * jb.plane : 0);
                        st_vec<Real, VEC>(d, o);
#pragma unroll
                        for (int k = 0; k < VEC; k++) {
                           if (k < 2 && k == kl) d[k - 2] = o[k];
                           if (k + 2 >= VEC && k == kh) d[k + 2] = o[k];
                        }
                     }
                  }
               }
            }
         }
         release(base + j + 1);  // plane x's stage may be refilled; x-1 and x+1 live in registers / the next stage
         u0p += jb.plane;
#pragma unroll
         for (int r = 0; r < RPT; r++) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
               um[r][k] = uc[r][k];
               uc[r][k] = up[r][k];
            }
         }
      }
      release(base + sg.cnt + 1);  // the last plane was only ever an "x+1" plane
      base += sg.cnt + 2;
   }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tile configurations (id, rows per thread, consumer warps, stages, register cap); cfg 0 is the default
#define PF_AIR_CONFIGS(X) \
   X(0, 2, 8, 4, 72)      \
   X(1, 4, 8, 3, 112)     \
   X(2, 1, 8, 4, 56)      \
   X(3, 1, 16, 4, 56)     \
   X(4, 2, 4, 4, 80)      \
   X(5, 2, 8, 3, 72)      \
   X(6, 2, 8, 5, 72)      \
   X(7, 2, 16, 3, 56)     \
   X(8, 2, 7, 4, 80)      \
   X(9, 2, 7, 4, 64)      \
   X(10, 1, 15, 4, 64)    \
   X(11, 2, 7, 3, 80)
#define PF_AIR_NCFG 12

template <typename Real>
static int air_tma_attr(int cfg) {
   cudaError_t rc = cudaErrorInvalidValue;
#define X(id, RPT, NW, S, MAXR)                                                                                     \
   if (cfg == id)                                                                                                   \
      rc = cudaFuncSetAttribute(k_air_tma_cart<Real, RPT, NW, S, MAXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                AirCfg<Real, RPT, NW, S>::SMEM_BYTES);
   PF_AIR_CONFIGS(X)
#undef X
   return (int)rc;
}

static void air_cfg_shape(int cfg, int *rpt, int *nw) {
   *rpt = 4, *nw = 8;
#define X(id, RPT, NW, S, MAXR) \
   if (cfg == id) *rpt = RPT, *nw = NW;
   PF_AIR_CONFIGS(X)
#undef X
}

static int air_tma_setup(AirTma *t, int precision, int fcc, i64 Nx, i64 Ny, i64 Nz, i64 Nzp, i64 mwpr, void *u_a, void *u_b, void *mask,
                         int cfg = 0) {
   t->ok = false;
   t->precision = precision, t->fcc = fcc, t->Nx = Nx, t->Ny = Ny, t->Nz = Nz, t->Nzp = Nzp, t->mwpr = mwpr;
   t->base[0] = u_a, t->base[1] = u_b, t->mask = mask;
   if (cfg < 0 || cfg >= PF_AIR_NCFG) {
      t->why = "no such tile configuration";
      return 1;
   }
   t->cfg = cfg;
   t->slots = 0;
   if (fcc != 0) {
      t->why = "13-point FCC runs on the generic kernel";
      return 1;
   }
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
   int rpt, nw;
   air_cfg_shape(cfg, &rpt, &nw);
   const CUtensorMapDataType dt = precision == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
   const cuuint64_t gdim[3] = {(cuuint64_t)Nzp, (cuuint64_t)Ny, (cuuint64_t)Nx};
   const cuuint64_t gstr[2] = {(cuuint64_t)Nzp * rs, (cuuint64_t)Ny * Nzp * rs};
   const cuuint32_t box1[3] = {(cuuint32_t)(32 * VEC + 2 * VEC), (cuuint32_t)(nw * rpt + 2), 1};
   const cuuint32_t box0[3] = {(cuuint32_t)(32 * VEC), (cuuint32_t)(nw * rpt), 1};
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
      const cuuint32_t mbox[3] = {4, (cuuint32_t)(nw * rpt), 1};
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

template <typename Real, int RPT, int NW, int S, int MAXR>
static int air_tma_launch_cfg(AirTma *t, int cur, Real *u0, i64 xb, i64 xe, Real a1, Real a2, const AirEdge<Real> &eg, cudaStream_t s) {
   typedef AirCfg<Real, RPT, NW, S> C;
   auto kern = k_air_tma_cart<Real, RPT, NW, S, MAXR>;
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
   jb.XC = std::min(jb.n, t->xc > 0 ? t->xc : 64);
   const i64 units = (i64)jb.n * jb.tiles;
   if (units > 0x7fffffff) return (int)cudaErrorInvalidValue;
   jb.units = (int)units;
   jb.Ny = (int)t->Ny, jb.Nz = (int)t->Nz, jb.Nzp = (int)t->Nzp;
   jb.plane = t->Ny * t->Nzp;
   // one CTA per resident slot, but never less than ~8 tile-planes per CTA
   const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(t->slots, units / 8));
   kern<<<grid, C::THREADS, C::SMEM_BYTES, s>>>(t->map_u1[cur], t->map_u0[cur ^ 1], t->map_mk, u0, jb, a1, a2, eg);
   return (int)cudaGetLastError();
}

// planes [xb, xe) of the slab; `cur` = index of the grid that currently is u1 (u0 = the other one)
template <typename Real>
static int air_tma_launch(AirTma *t, int cur, Real *u0, i64 xb, i64 xe, Real a1, Real a2, const AirEdge<Real> &eg, cudaStream_t s) {
#define X(id, RPT, NW, S, MAXR) \
   if (t->cfg == id) return air_tma_launch_cfg<Real, RPT, NW, S, MAXR>(t, cur, u0, xb, xe, a1, a2, eg, s);
   PF_AIR_CONFIGS(X)
#undef X
   return (int)cudaErrorInvalidValue;
}

}  // namespace pf
