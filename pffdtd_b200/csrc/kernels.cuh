// kernels.cuh -- device kernels of the simulation step, everything except the tiled TMA air kernel
// (air_tma.cuh).  One time step follows the reference C CPU engine (c_cuda/cpu_engine.h:129-325,
// SURVEY.md App. B) operation for operation: every product and sum below is an explicitly rounded
// IEEE operation (__fmul_rn/__fadd_rn/... never contract into FMAs), in the reference's order, so fp32
// and fp64 results are bit-identical to the CPU engine.
//
// Device layout: grids are [Nx][Ny][Nzp], z contiguous, Nzp = Nz rounded up to 32 elements so that
// every row starts on a 128-byte (fp32) / 256-byte (fp64) line and 16-byte vectors / TMA strides are
// legal.  All node lists are re-linearised to that pitch at create time.  `mask` has one bit per
// padded node, 32 nodes per word, `mwpr` words per row (Nzp/32 rounded up to 4 words so that a row of the
// mask is a legal TMA stride) (LSB first, the reference's convention fdtd_data.h:567-572 extended):
// a set bit means "the air update must not write this node": boundary nodes, the outer halo layer,
// the z padding and -- for the checkerboard FCC layout (fcc_flag 1) -- the unused odd-parity nodes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pf {

typedef long long i64;

template <typename Real> struct Ops;
template <> struct Ops<float> {
   static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
   static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
   static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
   static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Ops<double> {
   static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
   static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
   static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
   static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// neighbour offsets in adjacency-bit order (cpu_engine.h:249-254 Cartesian, :273-284 FCC)
struct Offsets {
   i64 o[12];
};

// ------------------------------------------------------------------------------------------------
// mask construction (create time)
// ------------------------------------------------------------------------------------------------
// one thread per 32-node word: halo layer, z padding, odd parity for fcc_flag==1
__global__ void k_mask_init(uint32_t *mask, i64 Nx, i64 Ny, i64 Nz, i64 mwpr, int fcc_flag, i64 ix0) {
   const i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (w >= Nx * Ny * mwpr) return;
   const i64 row = w / mwpr;
   const i64 z0 = (w - row * mwpr) << 5;
   const i64 ix = row / Ny, iy = row - ix * Ny;
   uint32_t bits = 0;
   const bool edge_row = (ix == 0) || (ix == Nx - 1) || (iy == 0) || (iy == Ny - 1);
   for (int b = 0; b < 32; b++) {
      const i64 iz = z0 + b;
      bool m = edge_row || iz == 0 || iz >= Nz - 1;
      if (fcc_flag == 1 && ((ix0 + ix + iy + iz) & 1)) m = true;
      bits |= (m ? 1u : 0u) << b;
   }
   mask[w] = bits;
}

__global__ void k_mask_nodes(uint32_t *mask, const i64 *bn, i64 Nb, i64 Nzp, i64 mwpr) {
   const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= Nb) return;
   const i64 c = bn[i];
   const i64 row = c / Nzp, iz = c - row * Nzp;
   atomicOr(&mask[row * mwpr + (iz >> 5)], 1u << (iz & 31));
}

// ------------------------------------------------------------------------------------------------
// step 1: u2ba[i] = u0[bna[i]]                                        (cpu_engine.h:131-134)
// ------------------------------------------------------------------------------------------------
template <typename Real>
__global__ void k_gather(const Real *__restrict__ u, const i64 *__restrict__ idx, Real *__restrict__ out, i64 n) {
   const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) out[i] = u[idx[i]];
}

// ------------------------------------------------------------------------------------------------
// step 2+3: folded-FCC seam row and halo mirrors                      (cpu_engine.h:135-172)
// The reference applies z, then y, then x mirrors, each over the FULL face, so edges and corners
// pick up already-mirrored values.  Three launches in stream order keep exactly that.
// ------------------------------------------------------------------------------------------------
// rows (ix,iy): optional seam copy is a separate kernel because it must precede the z mirror of row Ny-1
template <typename Real>
__global__ void k_fold_seam(Real *u1, i64 Nx, i64 Ny, i64 Nz, i64 Nzp) {
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (iz >= Nz) return;
   for (i64 ix = blockIdx.y; ix < Nx; ix += gridDim.y) {  // gridDim.y is capped at 65535
      Real *pl = u1 + ix * Ny * Nzp;
      pl[(Ny - 1) * Nzp + iz] = pl[(Ny - 2) * Nzp + iz];
   }
}

template <typename Real>
__global__ void k_flip_z(Real *u1, i64 nrows, i64 Nz, i64 Nzp) {
   const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= nrows) return;
   Real *row = u1 + r * Nzp;
   row[0] = row[2];
   row[Nz - 1] = row[Nz - 3];
}

template <typename Real>
__global__ void k_flip_y(Real *u1, i64 Nx, i64 Ny, i64 Nz, i64 Nzp, int do_yend) {
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (iz >= Nz) return;
   for (i64 ix = blockIdx.y; ix < Nx; ix += gridDim.y) {
      Real *pl = u1 + ix * Ny * Nzp;
      pl[iz] = pl[2 * Nzp + iz];
      if (do_yend) pl[(Ny - 1) * Nzp + iz] = pl[(Ny - 3) * Nzp + iz];
   }
}

template <typename Real>
__global__ void k_flip_x(Real *u1, i64 Nx, i64 Ny, i64 Nz, i64 Nzp, int lo, int hi) {
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (iz >= Nz) return;
   const i64 P = Ny * Nzp;
   for (i64 iy = blockIdx.y; iy < Ny; iy += gridDim.y) {
      const i64 r = iy * Nzp + iz;
      if (lo) u1[r] = u1[2 * P + r];
      if (hi) u1[(Nx - 1) * P + r] = u1[(Nx - 3) * P + r];
   }
}

// ------------------------------------------------------------------------------------------------
// step 4 (generic air kernel, one thread per node)                    (cpu_engine.h:175-223)
//   p = a1*u1[c] - u0[c];  p += a2*u1[c+off_j] for j in bit order;  u0[c] = p   unless masked
// ------------------------------------------------------------------------------------------------
template <typename Real, int NN>
__global__ void __launch_bounds__(256) k_air_generic(const Real *__restrict__ u1, Real *__restrict__ u0,
                                                      const uint32_t *__restrict__ mask, i64 Ny, i64 Nzp, i64 mwpr, i64 x_begin,
                                                      Real a1, Real a2, Offsets off) {
   typedef Ops<Real> O;
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   const i64 iy = (i64)blockIdx.y * blockDim.y + threadIdx.y;
   const i64 ix = x_begin + blockIdx.z;
   if (iz >= Nzp || iy >= Ny) return;
   const i64 c = (ix * Ny + iy) * Nzp + iz;
   if ((mask[(ix * Ny + iy) * mwpr + (iz >> 5)] >> (iz & 31)) & 1u) return;
   Real p = O::sub(O::mul(a1, u1[c]), u0[c]);
#pragma unroll
   for (int j = 0; j < NN; j++) p = O::add(p, O::mul(a2, u1[c + off.o[j]]));
   u0[c] = p;
}

// ------------------------------------------------------------------------------------------------
// step 5: absorbing shell                                             (cpu_engine.h:225-229)
//   lQ = l*Q (Real);  u0 = (u0 + lQ*u2ba)/(1.0 + lQ): the literal 1.0 makes the division a DOUBLE
//   division of a Real numerator in the reference, also when Real is float.
// ------------------------------------------------------------------------------------------------
// `zf` (optional): the shell's z faces also write the z halos of the NEW state, which the next step's mirror pass would otherwise
// fetch one 32-byte sector at a time (k_flip_z: 4 % of a 13-point step) -- here the sector is in hand.  Per node: bit 0 / 1 = the
// node is z = 1 / Nz-2 of a row off the x / y shell: copy its neighbour z = 2 / Nz-3 (an air node, final since the air kernel) to the
// halo z = 0 / Nz-1; bit 2 / 3 = the node is z = 2 / Nz-3 of a row ON the x / y shell (every node of such a row is in this list and
// gets its shell update from its own thread): it writes its own new value to the halo.  The engine only passes `zf` when no
// boundary / source node sits at z = 2 / Nz-3 (nothing after this kernel changes the copied values).
template <typename Real>
__global__ void k_abc(Real *__restrict__ u0, const i64 *__restrict__ bna, const int8_t *__restrict__ Q,
                      const Real *__restrict__ u2ba, i64 i0, i64 n, Real l, const uint8_t *__restrict__ zf) {
   typedef Ops<Real> O;
   const i64 i = i0 + (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= i0 + n) return;
   const Real lQ = O::mul(l, (Real)Q[i]);
   const i64 ib = bna[i];
   const Real num = O::add(u0[ib], O::mul(lQ, u2ba[i]));
   const double den = __dadd_rn(1.0, (double)lQ);
   const Real v = (Real)__ddiv_rn((double)num, den);
   u0[ib] = v;
   if (zf) {
      const unsigned f = zf[i];
      if (f & 1u) u0[ib - 1] = u0[ib + 1];
      if (f & 2u) u0[ib + 1] = u0[ib - 1];
      if (f & 4u) u0[ib - 2] = v;
      if (f & 8u) u0[ib + 2] = v;
   }
}

// ------------------------------------------------------------------------------------------------
// step 6: rigid boundary nodes                                        (cpu_engine.h:234-287)
//   b1 = 2 - sl2*K;  p = b1*u1[c] - u0[c];  p += (b2*(Real)bit_j)*u1[c+off_j]
// ------------------------------------------------------------------------------------------------
template <typename Real, int NN>
__global__ void k_rigid(const Real *__restrict__ u1, Real *__restrict__ u0, const i64 *__restrict__ bn,
                        const uint16_t *__restrict__ adj_bn, i64 i0, i64 n, Real sl2, Real a2, Offsets off) {
   typedef Ops<Real> O;
   // descending order: the air kernel swept x upwards, so the lines it touched last are the ones still in L2
   const i64 i = i0 + n - 1 - ((i64)blockIdx.x * blockDim.x + threadIdx.x);
   if (i < i0) return;
   const i64 c = bn[i];
   const unsigned adj = adj_bn[i];
   const Real K = (Real)__popc(adj);
   const Real b1 = O::sub((Real)2.0, O::mul(sl2, K));
   Real p = O::sub(O::mul(b1, u1[c]), u0[c]);
#pragma unroll
   for (int j = 0; j < NN; j++) {
      const Real bit = (Real)((adj >> j) & 1u);
      p = O::add(p, O::mul(O::mul(a2, bit), u1[c + off.o[j]]));
   }
   u0[c] = p;
}

// ------------------------------------------------------------------------------------------------
// step 7: frequency-dependent (RLC branch) boundary nodes      (cpu_engine.h:290-301, 363-405)
// State vh1/gh1 is stored branch-major [m][Nbl] so that a warp's accesses coalesce.
// hist = the node's value two steps back on entry (u2b), this step's value on exit (u0b -> u2b of n+2).
// The per-node constants of the reference's loop head,
//     lo2Kbg = lo2*ssaf*beta[k]        fac = 2*lo2*ssaf / (1 + lo2Kbg)
// never change, so k_fd_prep evaluates them once (same operations, same order, same bits) and packs
// (material, Mb) into one 16-bit word: the hot kernel then starts its 2*Mb state loads after ONE
// dependent load instead of three.
// ------------------------------------------------------------------------------------------------
// Branch state layout: groups of 32 consecutive lossy nodes; within a group, for each branch m the 32 v values then the 32 g values
// ([group][m][v|g][32]).  A warp's 2*Mb loads are still one aligned 128-byte line each, but they are NEIGHBOURING lines: the warp
// streams one contiguous block of up to 3 KB (fp32) instead of touching 24 arrays megabytes apart, which DRAM serves in long bursts.
// st_idx(m, i) is the v value; the g value sits 32 elements further.
template <int MMB>
__host__ __device__ __forceinline__ i64 st_idx(int m, i64 i) {
   return ((i >> 5) * (2 * MMB) + 2 * m) * 32 + (i & 31);
}

struct MatTable {
   const void *quads;  // Real [Nm][MMB][4] = b, bd, bDh, bFh
   const void *beta;   // Real [Nm]
   const int8_t *Mb;   // [Nm]
};

template <typename Real>
__global__ void k_fd_prep(const int8_t *__restrict__ mat_bnl, const Real *__restrict__ ssaf_bnl, i64 Nbl, Real lo2, MatTable mt,
                          Real *__restrict__ lo2Kbg_out, Real *__restrict__ fac_out, uint16_t *__restrict__ matmb) {
   typedef Ops<Real> O;
   const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= Nbl) return;
   const Real one = (Real)1.0, two = (Real)2.0;
   const int k = mat_bnl[i];
   const Real ssaf = ssaf_bnl[i];
   const Real lo2Kbg = O::mul(O::mul(lo2, ssaf), ((const Real *)mt.beta)[k]);
   lo2Kbg_out[i] = lo2Kbg;
   fac_out[i] = O::div(O::mul(O::mul(two, lo2), ssaf), O::add(one, lo2Kbg));
   matmb[i] = (uint16_t)((unsigned)k | ((unsigned)mt.Mb[k] << 8));
}

// `Nbl` below is the PITCH of the branch-major state arrays (the node count rounded up to 32 elements, so that a warp's
// 32 consecutive nodes are one aligned 128-byte line for every branch).
// x/2 is written x*0.5 (the same correctly rounded value; twelve of round 1's thirteen IEEE divisions per node were divisions by
// two) and the block stages (2*bDh, bFh, b, bd, 2*bFh) per (material, branch) in shared memory once (2*x is exact).  A version
// with the branch count fixed at compile time (11, every shipped material) was tried: 584 instead of 952 instructions, but the
// compiler wants 128 registers for it and spills at every cap that keeps the occupancy -- measured slower (c2: 0.291 vs 0.249 ms
// per step), dropped.
template <typename Real, int MMB>
__global__ void __launch_bounds__(128) k_fd(Real *__restrict__ u0, const i64 *__restrict__ bnl, const uint16_t *__restrict__ matmb,
                                            const Real *__restrict__ lo2Kbg_bnl, const Real *__restrict__ fac_bnl,
                                            Real *__restrict__ hist0, Real *__restrict__ hist1, Real *__restrict__ vh1,
                                            Real *__restrict__ gh1, i64 i0, i64 n, i64 Nbl, const Real *__restrict__ quads, int nquads,
                                            const i64 *__restrict__ d_n, int mb_max) {
   constexpr int MB = 0;
   typedef Ops<Real> O;
   extern __shared__ __align__(16) unsigned char fd_smem[];
   Real *qs = reinterpret_cast<Real *>(fd_smem);  // [Nm][MMB][5] = 2*bDh, bFh, b, bd, 2*bFh
   const Real one = (Real)1.0, two = (Real)2.0, half = (Real)0.5;
   const i64 iraw = i0 + n - 1 - ((i64)blockIdx.x * blockDim.x + threadIdx.x);  // descending, see k_rigid
   const bool active = iraw >= i0;
   const i64 i = active ? iraw : i0;  // (idle threads of the last block shadow node i0: they only help to stage the table)
   // Everything this node needs from memory is requested up front, before the block stages the coefficient table and meets at
   // the barrier (round 1 and the first round-2 version loaded the table, synchronised, and only then issued the node's loads:
   // a memory latency per block with nothing in flight).  The state loads are predicated on the LARGEST branch count of the
   // problem's materials (a kernel argument; the buffer holds MMB branches for every node, unused ones zero), not on this
   // node's own count, which is itself a load.  The two dependent loads (step parity -> history buffer, index -> u0) go last.
   const i64 dn = *d_n;
   const i64 c = bnl[i];
   const unsigned mm = matmb[i];
   constexpr int NB = MB > 0 ? MB : MMB;
   Real v1[NB], g1[NB];
   Real *pv = vh1 + st_idx<MMB>(0, i);  // (gh1 = vh1 + 32 in the grouped layout; Nbl is unused by it)
#pragma unroll
   for (int m = 0; m < NB; m++) {
      if (m < mb_max) {
         v1[m] = pv[64 * m];
         g1[m] = pv[64 * m + 32];
      }
   }
   const Real lo2Kbg = lo2Kbg_bnl[i], fac = fac_bnl[i];
   for (int t = threadIdx.x; t < nquads / 4; t += blockDim.x) {
      const Real b = quads[4 * t + 0], bd = quads[4 * t + 1], bDh = quads[4 * t + 2], bFh = quads[4 * t + 3];
      qs[5 * t + 0] = O::mul(two, bDh), qs[5 * t + 1] = bFh, qs[5 * t + 2] = b, qs[5 * t + 3] = bd, qs[5 * t + 4] = O::mul(two, bFh);
   }
   Real *hist = (dn & 1) ? hist1 : hist0;  // the value two steps back lives in the buffer of the step's parity
   const Real u2 = hist[i];
   const Real u0c = u0[c];
   __syncthreads();
   if (!active) return;
   const int Mb = MB > 0 ? MB : (int)(mm >> 8);
   const Real *q = qs + (mm & 0xffu) * (MMB * 5);
   const Real den = O::add(one, lo2Kbg);
   Real u = O::div(O::add(u0c, O::mul(lo2Kbg, u2)), den);
#pragma unroll
   for (int m = 0; m < NB; m++) {
      if (MB > 0 || m < Mb) u = O::sub(u, O::mul(fac, O::sub(O::mul(q[5 * m + 0], v1[m]), O::mul(q[5 * m + 1], g1[m]))));
   }
   const Real du = O::sub(u, u2);
   hist[i] = u;
   u0[c] = u;
#pragma unroll
   for (int m = 0; m < NB; m++) {
      if (MB > 0 || m < Mb) {
         const Real v0 = O::sub(O::add(O::mul(q[5 * m + 2], du), O::mul(q[5 * m + 3], v1[m])), O::mul(q[5 * m + 4], g1[m]));
         pv[64 * m + 32] = O::add(g1[m], O::mul(O::add(v0, v1[m]), half));
         pv[64 * m] = v0;
      }
   }
}

// The same update with the branch state moved by the TMA unit: a block owns 4 groups of 32 nodes, whose state is one contiguous
// run of lines per group in the grouped layout; one lane issues a bulk copy global -> shared per group (the first 2*mb_max lines),
// the threads read and rewrite their 2*Mb values in shared memory, and one lane issues the bulk stores back.  No state value passes
// through a register on its way in or out and the LSU queue (ncu: lg_throttle, 64 registers, 45 % occupancy in k_fd) is out of the
// picture.  Arithmetic and order are k_fd's.  Measured on B200 (c2): 53.6 us against k_fd's 50.9 -- the same 177 MB of DRAM reads, now
// waiting on shared-memory traffic (mio_throttle) instead of the LSU queue; option "fd_bulk", not the default.  What both pay for is
// the gather / scatter of u0 at one node per row on the z walls (a DRAM row activation per 4 useful bytes), not the state stream.
// Whole groups are copied in and out, also where a group straddles the end of the launch's range [i0, i0+n): the nodes outside
// it are written back unchanged (launches that share a state buffer run on one stream, in order).
namespace fdbulk {
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, int count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t *bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
   uint32_t ok;
   do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(s32(bar)), "r"(parity)
          : "memory");
   } while (!ok);
}
__device__ __forceinline__ void load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
                "r"(bytes), "r"(s32(bar))
                : "memory");
}
__device__ __forceinline__ void store(void *dst, const void *src, uint32_t bytes) {
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}
}  // namespace fdbulk

template <typename Real, int MMB>
__global__ void __launch_bounds__(128) k_fd_bulk(Real *__restrict__ u0, const i64 *__restrict__ bnl, const uint16_t *__restrict__ matmb,
                                                 const Real *__restrict__ lo2Kbg_bnl, const Real *__restrict__ fac_bnl,
                                                 Real *__restrict__ hist0, Real *__restrict__ hist1, Real *__restrict__ state, i64 i0, i64 n,
                                                 const Real *__restrict__ quads, int nquads, const i64 *__restrict__ d_n, int mb_max) {
   typedef Ops<Real> O;
   constexpr int GL = 2 * MMB * 32;  // elements of one group's block in the state buffer
   extern __shared__ __align__(128) unsigned char fdb_smem[];
   Real *st = reinterpret_cast<Real *>(fdb_smem);                         // [4][2*MMB][32]
   Real *qs = st + 4 * GL;                                                // [Nm][MMB][5] = 2*bDh, bFh, b, bd, 2*bFh
   uint64_t *bar = reinterpret_cast<uint64_t *>(qs + (nquads / 4 * 5 + 1) / 2 * 2);
   const Real one = (Real)1.0, two = (Real)2.0, half = (Real)0.5;
   // blocks walk the groups downwards from the end of the range (the lines the air kernel touched last are still in L2)
   const i64 gend = (i0 + n + 31) >> 5;                  // one past the last group of the range
   const i64 g0 = gend - 4 * ((i64)blockIdx.x + 1);      // first of this block's 4 groups (may be < the range's first group)
   const i64 gfirst = i0 >> 5;
   const int tid = threadIdx.x, gl = tid >> 5;
   const i64 i = (g0 + gl) * 32 + (tid & 31);
   const bool active = i >= i0 && i < i0 + n;
   const uint32_t gbytes = (uint32_t)(2 * mb_max * 32 * sizeof(Real));
   if (tid == 0) {
      fdbulk::bar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t total = 0;
      for (int g = 0; g < 4; g++)
         if (g0 + g >= gfirst && g0 + g < gend) total += gbytes;
      fdbulk::bar_expect(bar, total);
      for (int g = 0; g < 4; g++)
         if (g0 + g >= gfirst && g0 + g < gend) fdbulk::load(st + g * GL, state + (g0 + g) * GL, gbytes, bar);
   }
   // the node's own small loads go out while the state is on its way
   const i64 dn = *d_n;
   const i64 ii = active ? i : i0;
   const i64 c = bnl[ii];
   const unsigned mm = matmb[ii];
   const Real lo2Kbg = lo2Kbg_bnl[ii], fac = fac_bnl[ii];
   for (int t = tid; t < nquads / 4; t += blockDim.x) {
      const Real b = quads[4 * t + 0], bd = quads[4 * t + 1], bDh = quads[4 * t + 2], bFh = quads[4 * t + 3];
      qs[5 * t + 0] = O::mul(two, bDh), qs[5 * t + 1] = bFh, qs[5 * t + 2] = b, qs[5 * t + 3] = bd, qs[5 * t + 4] = O::mul(two, bFh);
   }
   Real *hist = (dn & 1) ? hist1 : hist0;  // the value two steps back lives in the buffer of the step's parity
   const Real u2 = hist[ii];
   const Real u0c = u0[c];
   __syncthreads();  // table staged, barrier initialised
   fdbulk::bar_wait(bar, 0);
   if (active) {
      const int Mb = (int)(mm >> 8);
      const Real *q = qs + (mm & 0xffu) * (MMB * 5);
      Real *sv = st + gl * GL + (tid & 31);  // v of branch m at sv[64*m], g at sv[64*m + 32]
      const Real den = O::add(one, lo2Kbg);
      Real u = O::div(O::add(u0c, O::mul(lo2Kbg, u2)), den);
#pragma unroll
      for (int m = 0; m < MMB; m++) {
         if (m < Mb) u = O::sub(u, O::mul(fac, O::sub(O::mul(q[5 * m + 0], sv[64 * m]), O::mul(q[5 * m + 1], sv[64 * m + 32]))));
      }
      const Real du = O::sub(u, u2);
      hist[i] = u;
      u0[c] = u;
#pragma unroll
      for (int m = 0; m < MMB; m++) {
         if (m < Mb) {
            const Real v1 = sv[64 * m], g1 = sv[64 * m + 32];
            const Real v0 = O::sub(O::add(O::mul(q[5 * m + 2], du), O::mul(q[5 * m + 3], v1)), O::mul(q[5 * m + 4], g1));
            sv[64 * m + 32] = O::add(g1, O::mul(O::add(v0, v1), half));
            sv[64 * m] = v0;
         }
      }
   }
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy writes above, before the bulk stores read them
   __syncthreads();
   if (tid == 0) {
      for (int g = 0; g < 4; g++)
         if (g0 + g >= gfirst && g0 + g < gend) fdbulk::store(state + (g0 + g) * GL, st + g * GL, gbytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory must outlive the stores' reads; writes done before exit
   }
}

// ------------------------------------------------------------------------------------------------
// step 9: sources are added to the NEW state u0 (cpu_engine.h:310-313); step 8, the receivers
// reading the CURRENT state u1 (cpu_engine.h:304-307), is k_gather.
// `serial_src` keeps the reference's serial order when two source entries name the same node.
// ------------------------------------------------------------------------------------------------
template <typename Real>
__global__ void k_src(Real *__restrict__ u0, const i64 *__restrict__ in_ixyz, const Real *__restrict__ in_row, i64 s0, i64 ns,
                      int serial_src) {
   typedef Ops<Real> O;
   const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (serial_src) {
      if (i == 0)
         for (i64 s = s0; s < s0 + ns; s++) u0[in_ixyz[s]] = O::add(u0[in_ixyz[s]], in_row[s]);
   } else if (i < ns) {
      u0[in_ixyz[s0 + i]] = O::add(u0[in_ixyz[s0 + i]], in_row[s0 + i]);
   }
}

// steps 8+9 in one launch: receivers (threads [0,Nr)) and sources (threads [0,ns)) touch different grids.
// The time index comes from a device counter (so that a captured step can be replayed as a CUDA graph); `nr` = 0
// for the parts of a step that do not read the receivers.  Host-driven steps (pffdtd_step_host) pass fixed staging
// buffers: `in_stage` holds this step's source samples (also filed into insig), `out_stage` receives a copy of the
// receiver samples -- fixed addresses, so the host<->device copies can be nodes of the same replayed graph.
template <typename Real>
__global__ void k_io(const Real *__restrict__ u1, Real *__restrict__ u0, const i64 *__restrict__ out_ixyz, Real *__restrict__ uout,
                     i64 Nr, i64 nr_all, const i64 *__restrict__ in_ixyz, Real *__restrict__ insig, i64 Ns_all, i64 s0, i64 ns,
                     int serial_src, i64 *__restrict__ d_n, const Real *__restrict__ in_stage, Real *__restrict__ out_stage, int tick) {
   typedef Ops<Real> O;
   const i64 n = *d_n;
   if (tick) {  // single-block launch that closes the step: advance the device step counter here (saves the k_tick launch)
      __syncthreads();
      if (threadIdx.x == 0) *d_n = n + 1;
   }
   Real *out_row = uout + n * nr_all;
   Real *in_row = insig + n * Ns_all;
   const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < Nr) {
      const Real v = u1[out_ixyz[i]];
      out_row[i] = v;
      if (out_stage) out_stage[i] = v;
   }
   if (serial_src) {
      if (i == 0)
         for (i64 s = s0; s < s0 + ns; s++) {
            if (in_stage) in_row[s] = in_stage[s];
            u0[in_ixyz[s]] = O::add(u0[in_ixyz[s]], in_row[s]);
         }
   } else if (i < ns) {
      if (in_stage) in_row[s0 + i] = in_stage[s0 + i];
      u0[in_ixyz[s0 + i]] = O::add(u0[in_ixyz[s0 + i]], in_row[s0 + i]);
   }
}

// Halo exchange between PROCESSES over peer memory (engine.cu "p2p").  Every slab counts its steps in a device word `seq`.
// A step starts with k_p2p_wait: it spins until each neighbour's flag has reached the number of steps done so far (the
// neighbour has delivered the plane of every earlier step), then counts the step.  After the step's edge planes are final they
// are copied into the neighbour's halo planes through a CUDA IPC mapping of its grid, and the same stream then copies `seq`
// into the neighbour's flag word -- two copy-engine transfers, no kernel: the persistent air kernel holds every SM, so a
// signalling kernel (or NCCL's) would only run once it drains, which is why round 1's "overlapped" NCCL exchange never overlapped.
// The counters live on the device, so the step replays from graphs.  A wait that lasts longer than ~20 s (a dead neighbour)
// gives up and raises the error word instead of hanging the GPU.
__global__ void k_p2p_wait(volatile long long *flags, int need_lo, int need_hi, long long *seq) {
   const long long done = *seq;
   const long long t0 = clock64();
   for (int k = 0; k < 2; k++) {
      if (!(k == 0 ? need_lo : need_hi)) continue;
      while (flags[k] < done) {
         __nanosleep(200);
         if (clock64() - t0 > 40000000000ll) {
            flags[3] = 1;  // error word, read by pffdtd_sync
            break;
         }
      }
   }
   __threadfence_system();
   *seq = done + 1;
}

// device step counter: set at the start of a batch of steps, advanced at the end of every step
__global__ void k_set_n(i64 *d_n, i64 v) { *d_n = v; }
__global__ void k_tick(i64 *d_n) { *d_n += 1; }

// halo mirrors of nodes that were written AFTER the fused air kernel (boundary and source nodes sitting
// at index 2 / N-3 of an axis): u[dst] = u[src] for a precomputed list; usually empty
template <typename Real>
__global__ void k_pairs(Real *__restrict__ u, const i64 *__restrict__ src, const i64 *__restrict__ dst, i64 i0, i64 n) {
   const i64 i = i0 + (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < i0 + n) u[dst[i]] = u[src[i]];
}

// ------------------------------------------------------------------------------------------------
// layout conversion for grid read-back / initialisation (reference layout <-> padded, double <-> Real)
// ------------------------------------------------------------------------------------------------
template <typename Real>
__global__ void k_unpad(const Real *__restrict__ u, double *__restrict__ out, i64 nrows, i64 Nz, i64 Nzp) {
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   const i64 r = blockIdx.y + (i64)blockIdx.z * gridDim.y;
   if (iz < Nz && r < nrows) out[r * Nz + iz] = (double)u[r * Nzp + iz];
}
template <typename Real>
__global__ void k_pad(Real *__restrict__ u, const double *__restrict__ in, i64 nrows, i64 Nz, i64 Nzp) {
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   const i64 r = blockIdx.y + (i64)blockIdx.z * gridDim.y;
   if (iz < Nz && r < nrows) u[r * Nzp + iz] = (Real)in[r * Nz + iz];
}

}  // namespace pf
