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
   int xc = 0;          // planes per CTA chunk, 0 = automatic
   int sm_count = 148;
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
   static constexpr int SMEM_BYTES = S * STAGE_PITCH + S * 8 + 128;
};

// ---------------------------------------------------------------- the kernel (7-point Cartesian)
template <typename Real, int RPT, int NW, int S>
__global__ void __launch_bounds__(NW * 32, 2)
    k_air_tma_cart(const __grid_constant__ CUtensorMap map_u1, Real *__restrict__ u0g, const uint32_t *__restrict__ mask, i64 Ny,
                   i64 Nz, i64 Nzp, int x_begin, int x_end, int XC, Real a1, Real a2) {
   typedef AirCfg<Real, RPT, NW, S> C;
   typedef Ops<Real> O;
   constexpr int VEC = C::VEC, BZ = C::BZ;
   extern __shared__ unsigned char smem_raw[];
   unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
   uint64_t *full = (uint64_t *)(smem + S * C::STAGE_PITCH);

   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const int z0 = blockIdx.x * C::TZ;
   const int y0 = 1 + blockIdx.y * C::TY;
   const int xa = x_begin + blockIdx.z * XC;
   const int xe = min(x_end, xa + XC);
   const int nsteps = xe - xa;
   if (nsteps <= 0) return;
   const int L = nsteps + 2;  // planes xa-1 .. xe

   auto stage = [&](int i) -> Real * { return (Real *)(smem + (i % S) * C::STAGE_PITCH); };
   auto issue = [&](int i) {
      uint64_t *bar = &full[i % S];
      mbar_expect_tx(bar, C::STAGE_BYTES);
      tma_load_3d(stage(i), &map_u1, bar, z0 - VEC, y0 - 1, xa - 1 + i);
   };
   auto wait = [&](int i) { mbar_wait(&full[i % S], (uint32_t)((i / S) & 1)); };

   if (tid == 0) {
      for (int s = 0; s < S; s++) mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
   }
   __syncthreads();
   if (tid == 0) {
      for (int i = 0; i < S && i < L; i++) issue(i);
   }

   // this thread's strip: rows y0 + w*RPT + r, columns z0 + VEC*lane .. +VEC-1
   const int zv = z0 + VEC * lane;
   const int ybase = y0 + w * RPT;
   const bool zact = zv < Nz - 1;  // vectors entirely in the far halo/padding are never touched
   const int srow0 = w * RPT + 1;  // shared-memory row of strip row 0 (box row 0 is y0-1)
   const int scol = VEC + VEC * lane;
   // mask bits of this thread's vector inside the row's 32-node word
   const int mword = zv >> 5, mshift = zv & 31;

   Real um[RPT][VEC], uc[RPT][VEC], up[RPT][VEC];
   wait(0);
   {
      const Real *s0 = stage(0);
#pragma unroll
      for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(s0 + (srow0 + r) * BZ + scol, um[r]);
   }
   wait(1);
   {
      const Real *s1 = stage(1);
#pragma unroll
      for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(s1 + (srow0 + r) * BZ + scol, uc[r]);
   }
   __syncthreads();
   if (tid == 0 && S < L) issue(S);  // stage 0 (plane xa-1) now lives in registers

   // u0 / mask prefetch for the first plane
   Real u0n[RPT][VEC];
   uint32_t mkn[RPT];
   bool ract[RPT];
#pragma unroll
   for (int r = 0; r < RPT; r++) ract[r] = zact && (ybase + r) <= Ny - 2;
   auto fetch = [&](int x) {
#pragma unroll
      for (int r = 0; r < RPT; r++) {
         if (ract[r]) {
            const i64 row = (i64)x * Ny + (ybase + r);
            ld_vec<Real, VEC>(u0g + row * Nzp + zv, u0n[r]);
            mkn[r] = __ldg(mask + row * (Nzp >> 5) + mword) >> mshift;
         }
      }
   };
   fetch(xa);

   for (int j = 0; j < nsteps; j++) {
      const int x = xa + j;
      wait(j + 2);
      const Real *sc = stage(j + 1);
      const Real *su = stage(j + 2);
      Real rowm[VEC], rowp[VEC], zl[RPT], zr[RPT];
#pragma unroll
      for (int r = 0; r < RPT; r++) ld_vec<Real, VEC>(su + (srow0 + r) * BZ + scol, up[r]);
      ld_vec<Real, VEC>(sc + (srow0 - 1) * BZ + scol, rowm);
      ld_vec<Real, VEC>(sc + (srow0 + RPT) * BZ + scol, rowp);
#pragma unroll
      for (int r = 0; r < RPT; r++) {
         zl[r] = sc[(srow0 + r) * BZ + scol - 1];
         zr[r] = sc[(srow0 + r) * BZ + scol + VEC];
      }
      Real u0c[RPT][VEC];
      uint32_t mk[RPT];
#pragma unroll
      for (int r = 0; r < RPT; r++) {
         mk[r] = mkn[r];
#pragma unroll
         for (int k = 0; k < VEC; k++) u0c[r][k] = u0n[r][k];
      }
      if (j + 1 < nsteps) fetch(x + 1);
#pragma unroll
      for (int r = 0; r < RPT; r++) {
         if (ract[r]) {
            Real o[VEC];
#pragma unroll
            for (int k = 0; k < VEC; k++) {
               const Real yp = (r < RPT - 1) ? uc[r + 1][k] : rowp[k];
               const Real ym = (r > 0) ? uc[r - 1][k] : rowm[k];
               const Real zp = (k < VEC - 1) ? uc[r][k + 1] : zr[r];
               const Real zm = (k > 0) ? uc[r][k - 1] : zl[r];
               Real p = O::sub(O::mul(a1, uc[r][k]), u0c[r][k]);
               p = O::add(p, O::mul(a2, up[r][k]));
               p = O::add(p, O::mul(a2, um[r][k]));
               p = O::add(p, O::mul(a2, yp));
               p = O::add(p, O::mul(a2, ym));
               p = O::add(p, O::mul(a2, zp));
               p = O::add(p, O::mul(a2, zm));
               o[k] = ((mk[r] >> k) & 1u) ? u0c[r][k] : p;
            }
            if ((mk[r] & ((1u << VEC) - 1u)) != ((1u << VEC) - 1u)) {
               const i64 row = (i64)x * Ny + (ybase + r);
               st_vec<Real, VEC>(u0g + row * Nzp + zv, o);
            }
         }
      }
#pragma unroll
      for (int r = 0; r < RPT; r++) {
#pragma unroll
         for (int k = 0; k < VEC; k++) {
            um[r][k] = uc[r][k];
            uc[r][k] = up[r][k];
         }
      }
      __syncthreads();  // everyone is done with plane x's stage
      if (tid == 0 && j + 1 + S < L) issue(j + 1 + S);
   }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tile shape: 8 warps x 4 rows, 4 planes in flight
#define PF_AIR_RPT 4
#define PF_AIR_NW 8
#define PF_AIR_S 4

template <typename Real>
static int air_tma_attr() {
   typedef AirCfg<Real, PF_AIR_RPT, PF_AIR_NW, PF_AIR_S> C;
   return (int)cudaFuncSetAttribute(k_air_tma_cart<Real, PF_AIR_RPT, PF_AIR_NW, PF_AIR_S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES);
}

static int air_tma_setup(AirTma *t, int precision, int fcc, i64 Nx, i64 Ny, i64 Nz, i64 Nzp, void *u_a, void *u_b) {
   t->ok = false;
   t->precision = precision, t->fcc = fcc, t->Nx = Nx, t->Ny = Ny, t->Nz = Nz, t->Nzp = Nzp;
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
   const cuuint64_t gdim[3] = {(cuuint64_t)Nzp, (cuuint64_t)Ny, (cuuint64_t)Nx};
   const cuuint64_t gstr[2] = {(cuuint64_t)Nzp * rs, (cuuint64_t)Ny * Nzp * rs};
   const cuuint32_t box[3] = {(cuuint32_t)(32 * VEC + 2 * VEC), (cuuint32_t)(PF_AIR_NW * PF_AIR_RPT + 2), 1};
   const cuuint32_t estr[3] = {1, 1, 1};
   void *bases[2] = {u_a, u_b};
   for (int k = 0; k < 2; k++) {
      CUresult r = encode(&t->map[k], precision == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, bases[k],
                          gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
         t->why = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
         return 1;
      }
   }
   int rc = precision == 1 ? air_tma_attr<float>() : air_tma_attr<double>();
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

// planes [xb, xe) of the slab; `cur` = index of the grid that currently is u1
template <typename Real>
static int air_tma_launch(AirTma *t, int cur, const Real *u1, Real *u0, const uint32_t *mask, i64 xb, i64 xe, Real a1, Real a2,
                          cudaStream_t s) {
   typedef AirCfg<Real, PF_AIR_RPT, PF_AIR_NW, PF_AIR_S> C;
   (void)u1;
   const int n = (int)(xe - xb);
   const unsigned tz = (unsigned)((t->Nz - 1 + C::TZ - 1) / C::TZ);  // vectors starting at z >= Nz-1 hold no interior node
   const unsigned ty = (unsigned)((t->Ny - 2 + C::TY - 1) / C::TY);
   int xc = t->xc;
   if (xc <= 0) {
      // enough CTAs for ~4 rounds of 2 resident CTAs per SM, but chunks no shorter than 8 planes (2 extra
      // plane loads per chunk)
      const i64 tiles = (i64)tz * ty;
      i64 chunks = (8LL * t->sm_count + tiles - 1) / tiles;
      chunks = std::max<i64>(1, std::min<i64>(chunks, std::max(1, n / 8)));
      xc = (int)((n + chunks - 1) / chunks);
   }
   const unsigned nch = (unsigned)((n + xc - 1) / xc);
   dim3 grd(tz, ty, nch);
   k_air_tma_cart<Real, PF_AIR_RPT, PF_AIR_NW, PF_AIR_S><<<grd, PF_AIR_NW * 32, C::SMEM_BYTES, s>>>(
       t->map[cur], u0, mask, t->Ny, t->Nz, t->Nzp, (int)xb, (int)xe, xc, a1, a2);
   return (int)cudaGetLastError();
}

}  // namespace pf
