// air_tma.cuh -- the interior air update as a 2.5-D blocked sweep for sm_100a.
//
// Replaces KernelAirCart of the reference (c_cuda/gpu_engine.h:220-242; one thread per node, seven
// scalar loads through L1) with a tile kernel built for Blackwell:
//   * a CTA owns a (TY x TZ) tile of the y-z plane and sweeps a chunk of x-planes;
//   * each u1 plane tile (+1-node halo) is brought into shared memory ONCE by TMA
//     (cp.async.bulk.tensor.3d, mbarrier complete_tx), S planes deep, so the loads of planes
//     x+2..x+S-1 are in flight while plane x is computed; out-of-grid parts of a box are zero-filled
//     by the TMA unit, so ragged tiles need no special code;
//   * a thread owns RPT rows x one 16-byte vector of z (4 fp32 / 2 fp64) and keeps the x-1, x, x+1
//     values of its own columns in registers, rotating them along the sweep: the +-x taps never
//     touch memory again, +-y taps of inner rows come from the thread's own registers, only the two
//     rows next to the thread's strip and the +-z end taps are read from shared memory;
//   * u0 is read and written straight from/to HBM with 128-bit accesses, prefetched one plane
//     ahead; the "do not write" mask is consumed as bits of one 32-bit word per row.
// Arithmetic is the reference CPU engine's (cpu_engine.h:182-189): a1*u1 - u0, then six separately
// rounded a2*u1[nb] products added in the order +x -x +y -y +z -z.  Masked lanes store the old value
// back, so every store is a full aligned vector.
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
   i64 Nx = 0, Ny = 0, Nz = 0, Nzp = 0;
   CUtensorMap map[2];  // over u[0], u[1]
   void *base[2] = {nullptr, nullptr};
   int cfg = 0;         // tile configuration, see PF_AIR_CONFIGS
   int xc = 0;          // planes per CTA chunk, 0 = automatic
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
   static constexpr int BZ = TZ + 2 * VEC;  // box starts one vector left of the tile: 16-byte aligned columns
   static constexpr int ROWS = TY + 2;
   static constexpr int STAGE_BYTES = ROWS * BZ * (int)sizeof(Real);
   static constexpr int STAGE_PITCH = (STAGE_BYTES + 127) / 128 * 128;
   static constexpr int SMEM_BYTES = S * STAGE_PITCH + 2 * S * 8 + 128;
   static constexpr int THREADS = (NW + 1) * 32;  // NW consumer warps + 1 TMA producer warp
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- work decomposition
// The job "planes [x_begin, x_end) x all y-z tiles" is cut into units of one tile-plane, ordered
// (x-chunk of XC planes, tile, plane in chunk), and every CTA of a persistent grid (one CTA per
// resident slot) takes an equal contiguous range of units: all CTAs finish together (no tail), and CTAs
// running at the same time work on the same x-chunk of neighbouring tiles, so tile halos hit in L2.
// A CTA's range is a few "segments" (one tile, consecutive planes); each costs two extra plane loads.
struct AirJob {
   int x_begin, n, XC, tz, tiles;  // n planes, tiles = tz*ty
   i64 units;                      // n * tiles
};
struct AirSeg {
   int xa, cnt, z0, y0;
   i64 next;  // first unit after the segment
};
template <int TZ, int TY>
__device__ __forceinline__ AirSeg air_segment(const AirJob &jb, i64 u, i64 u_end) {
   const i64 per_chunk = (i64)jb.XC * jb.tiles;
   const int k = (int)(u / per_chunk);
   const int len = min(jb.XC, jb.n - k * jb.XC);
   const i64 up = u - (i64)k * per_chunk;
   const int t = (int)(up / len), p = (int)(up - (i64)t * len);
   AirSeg s;
   s.next = min(u_end, (i64)k * per_chunk + (i64)(t + 1) * len);
   s.cnt = (int)(s.next - u);
   s.xa = jb.x_begin + k * jb.XC + p;
   s.z0 = (t % jb.tz) * TZ;
   s.y0 = 1 + (t / jb.tz) * TY;
   return s;
}

// ---------------------------------------------------------------- the kernel (7-point Cartesian)
// Warp-specialised: warp NW is the TMA producer (one lane), warps 0..NW-1 compute.  full[s] flips when
// the bytes of a plane have landed in stage s, empty[s] when all NW consumer warps are done with it.
// Loads are numbered consecutively over all segments of the CTA, load i uses stage i % S.
template <typename Real, int RPT, int NW, int S, int MAXR>
__global__ void __maxnreg__(MAXR)
    k_air_tma_cart(const __grid_constant__ CUtensorMap map_u1, Real *__restrict__ u0g, const uint32_t *__restrict__ mask, i64 Ny,
                   i64 Nz, i64 Nzp, AirJob jb, Real a1, Real a2) {
   typedef AirCfg<Real, RPT, NW, S> C;
   typedef Ops<Real> O;
   constexpr int VEC = C::VEC, BZ = C::BZ;
   constexpr uint32_t VMASK = (1u << VEC) - 1u;
   extern __shared__ unsigned char smem_raw[];
   unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
   uint64_t *full = (uint64_t *)(smem + S * C::STAGE_PITCH);
   uint64_t *empty = full + S;

   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const i64 u_begin = jb.units * blockIdx.x / gridDim.x, u_end = jb.units * (blockIdx.x + 1) / gridDim.x;
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
         for (i64 u = u_begin; u < u_end;) {
            const AirSeg sg = air_segment<C::TZ, C::TY>(jb, u, u_end);
            for (int q = 0; q < sg.cnt + 2; q++, i++) {  // planes xa-1 .. xa+cnt
               const int s = i % S;
               if (i >= S) mbar_wait(&empty[s], (uint32_t)(((i / S) - 1) & 1));
               mbar_expect_tx(&full[s], C::STAGE_BYTES);
               tma_load_3d(smem + s * C::STAGE_PITCH, &map_u1, &full[s], sg.z0 - VEC, sg.y0 - 1, sg.xa - 1 + q);
            }
            u = sg.next;
         }
      }
      return;
   }

   // ---------------- consumers
   auto stage = [&](int i) -> const Real * { return (const Real *)(smem + (i % S) * C::STAGE_PITCH); };
   auto wait_full = [&](int i) { mbar_wait(&full[i % S], (uint32_t)((i / S) & 1)); };
   auto release = [&](int i) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[i % S]);
   };
   const int soff = (w * RPT + 1) * BZ + VEC + VEC * lane;  // strip row 0 inside a stage (box row 0 is y0-1)
   const i64 u0_plane = Ny * Nzp, mk_plane = Ny * (Nzp >> 5), mk_row = Nzp >> 5;

   int base = 0;  // load index of the segment's first plane
   for (i64 u = u_begin; u < u_end;) {
      const AirSeg sg = air_segment<C::TZ, C::TY>(jb, u, u_end);
      u = sg.next;
      // this thread's strip: rows y0 + w*RPT + r, columns z0 + VEC*lane .. +VEC-1
      const int zv = sg.z0 + VEC * lane;
      const int ybase = sg.y0 + w * RPT;
      const bool zact = zv < Nz - 1;  // vectors entirely in the far halo/padding are never touched
      const int mshift = zv & 31;
      int nrow = 0;  // active rows of the strip
#pragma unroll
      for (int r = 0; r < RPT; r++) nrow += (zact && (ybase + r) <= Ny - 2) ? 1 : 0;

      Real um[RPT][VEC], uc[RPT][VEC], up[RPT][VEC];
      wait_full(base);
      {
         const Real *s0 = stage(base) + soff;
#pragma unroll
         for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(s0 + r * BZ, um[r]);
      }
      release(base);
      wait_full(base + 1);
      {
         const Real *s1 = stage(base + 1) + soff;
#pragma unroll
         for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(s1 + r * BZ, uc[r]);
      }
      // u0 / mask of the first plane; afterwards each row's registers are refilled for the next plane as
      // soon as the row has been stored, so the HBM loads of plane x+1 fly while plane x is computed
      Real *u0p = u0g + ((i64)sg.xa * Ny + ybase) * Nzp + zv;
      const uint32_t *mkp = mask + ((i64)sg.xa * Ny + ybase) * mk_row + (zv >> 5);
      Real u0v[RPT][VEC];
      uint32_t mk[RPT];
#pragma unroll
      for (int r = 0; r < RPT; r++) {
         if (r < nrow) {
            ld_vec<Real, VEC>(u0p + r * Nzp, u0v[r]);
            mk[r] = __ldg(mkp + r * mk_row);
         }
      }

      for (int j = 0; j < sg.cnt; j++) {
         wait_full(base + j + 2);
         const Real *sc = stage(base + j + 1) + soff;
         const Real *su = stage(base + j + 2) + soff;
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
         release(base + j + 1);  // plane x's stage may be refilled; x-1 and x+1 live in registers / the next stage
         const bool more = j + 1 < sg.cnt;
#pragma unroll
         for (int r = 0; r < RPT; r++) {
            if (r < nrow) {
               const uint32_t m = (mk[r] >> mshift) & VMASK;
               Real o[VEC];
#pragma unroll
               for (int k = 0; k < VEC; k++) {
                  const Real yp = (r < RPT - 1) ? uc[r + 1][k] : rowp[k];
                  const Real ym = (r > 0) ? uc[r - 1][k] : rowm[k];
                  const Real zp = (k < VEC - 1) ? uc[r][k + 1] : zr[r];
                  const Real zm = (k > 0) ? uc[r][k - 1] : zl[r];
                  Real p = O::sub(O::mul(a1, uc[r][k]), u0v[r][k]);
                  p = O::add(p, O::mul(a2, up[r][k]));
                  p = O::add(p, O::mul(a2, um[r][k]));
                  p = O::add(p, O::mul(a2, yp));
                  p = O::add(p, O::mul(a2, ym));
                  p = O::add(p, O::mul(a2, zp));
                  p = O::add(p, O::mul(a2, zm));
                  o[k] = ((m >> k) & 1u) ? u0v[r][k] : p;
               }
               if (m != VMASK) st_vec<Real, VEC>(u0p + r * Nzp, o);
               if (more) {
                  ld_vec<Real, VEC>(u0p + u0_plane + r * Nzp, u0v[r]);
                  mk[r] = __ldg(mkp + mk_plane + r * mk_row);
               }
            }
         }
         u0p += u0_plane;
         mkp += mk_plane;
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

// tile configurations (rows per thread, consumer warps, planes in flight, register cap); cfg 0 is the default
#define PF_AIR_CONFIGS(X) \
   X(0, 2, 8, 4, 72)      \
   X(1, 4, 8, 4, 112)     \
   X(2, 1, 8, 4, 56)      \
   X(3, 1, 16, 4, 56)     \
   X(4, 2, 4, 4, 80)      \
   X(5, 2, 8, 3, 72)      \
   X(6, 2, 8, 5, 72)      \
   X(7, 2, 16, 4, 56)
#define PF_AIR_NCFG 8

template <typename Real>
static int air_tma_attr(int cfg) {
   cudaError_t rc = cudaErrorInvalidValue;
#define X(id, RPT, NW, S, MAXR)                                                                                              \
   if (cfg == id)                                                                                                            \
      rc = cudaFuncSetAttribute(k_air_tma_cart<Real, RPT, NW, S, MAXR>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
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

static int air_tma_setup(AirTma *t, int precision, int fcc, i64 Nx, i64 Ny, i64 Nz, i64 Nzp, void *u_a, void *u_b, int cfg = 0) {
   t->ok = false;
   t->precision = precision, t->fcc = fcc, t->Nx = Nx, t->Ny = Ny, t->Nz = Nz, t->Nzp = Nzp;
   t->base[0] = u_a, t->base[1] = u_b;
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
   const cuuint64_t gdim[3] = {(cuuint64_t)Nzp, (cuuint64_t)Ny, (cuuint64_t)Nx};
   const cuuint64_t gstr[2] = {(cuuint64_t)Nzp * rs, (cuuint64_t)Ny * Nzp * rs};
   const cuuint32_t box[3] = {(cuuint32_t)(32 * VEC + 2 * VEC), (cuuint32_t)(nw * rpt + 2), 1};
   const cuuint32_t estr[3] = {1, 1, 1};
   for (int k = 0; k < 2; k++) {
      CUresult r = encode(&t->map[k], precision == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, t->base[k],
                          gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
         t->why = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
         return 1;
      }
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
static int air_tma_launch_cfg(AirTma *t, int cur, Real *u0, const uint32_t *mask, i64 xb, i64 xe, Real a1, Real a2, cudaStream_t s) {
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
   jb.units = (i64)jb.n * jb.tiles;
   // one CTA per resident slot, but never less than ~8 tile-planes per CTA
   const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(t->slots, jb.units / 8));
   kern<<<grid, C::THREADS, C::SMEM_BYTES, s>>>(t->map[cur], u0, mask, t->Ny, t->Nz, t->Nzp, jb, a1, a2);
   return (int)cudaGetLastError();
}

// planes [xb, xe) of the slab; `cur` = index of the grid that currently is u1
template <typename Real>
static int air_tma_launch(AirTma *t, int cur, const Real *u1, Real *u0, const uint32_t *mask, i64 xb, i64 xe, Real a1, Real a2,
                          cudaStream_t s) {
   (void)u1;
#define X(id, RPT, NW, S, MAXR) \
   if (t->cfg == id) return air_tma_launch_cfg<Real, RPT, NW, S, MAXR>(t, cur, u0, mask, xb, xe, a1, a2, s);
   PF_AIR_CONFIGS(X)
#undef X
   return (int)cudaErrorInvalidValue;
}

}  // namespace pf
