// energy.cuh -- the reference Python engine's energy balance (python/fdtd/sim_fdtd.py:587-620, kernels
// :706-770, 838-856) evaluated on the device, so that the invariant can be checked on
// grids far beyond what the CPU engines reach (BASELINE configs[3], 1024^3 fp64).
//
// Same data flow as the reference: a third grid Lu holds the boundary-aware Laplacian of the previous
// state (nb_stencil_air_* + nb_stencil_bn_*, written after the halo mirrors of each step), and per step
//     H_tot[n]    =  V*h/2 * sum_interior[(u1-u2)^2/l2 - u1*Lu2]
//                  - V*h/2 * sum_ABC (1-2^-Q)[(u1-u2)^2/l2 - u1*Lu2]
//                  + V*c/(2 l2) * sum_lossy ssaf sum_m [vh1^2 D + (Ts gh1)^2 F]
//     E_lost[n+1] = E_lost[n] + V*h/(4 l) * sum_lossy ssaf sum_m (vh0+vh1)^2 E
//                             + V*h/(2 l) * sum_ABC 2^-Q Q (u0_new - u2ba)^2
//     E_in[n+1]   = E_in[n]   + V*h/(2 l2) * sum_src (u0_new[in] - u2in) * in_sigs[:,n]
// with V = 1 (Cartesian) or 2 (FCC).  All sums are accumulated in double in a FIXED order (per-block
// partials, then one block adds the partials), so the numbers are reproducible run to run.
#pragma once
#include "kernels.cuh"

namespace pf {

constexpr int EN_BLOCKS = 1184;  // 148 SMs x 8
constexpr int EN_THREADS = 256;

// block-wide sum in a fixed order; thread 0 returns the total
__device__ __forceinline__ double en_block_sum(double v) {
   __shared__ double sh[32];
#pragma unroll
   for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
   __syncthreads();  // sh may still be read by a previous call
   if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
   __syncthreads();
   double w = 0.0;
   if (threadIdx.x < 32) {
      w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
      for (int o = 16; o; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
   }
   return w;
}

// Lu = lfac*(-NN*u + sum of neighbours) on every unmasked interior node     (sim_fdtd.py:699-735)
template <typename Real, int NN>
__global__ void __launch_bounds__(256) k_lap_air(const Real *__restrict__ u1, Real *__restrict__ Lu, const uint32_t *__restrict__ mask,
                                                 i64 Ny, i64 Nzp, i64 mwpr, i64 x_begin, double lfac, Offsets off) {
   const i64 iz = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   const i64 iy = (i64)blockIdx.y * blockDim.y + threadIdx.y;
   const i64 ix = x_begin + blockIdx.z;
   if (iz >= Nzp || iy >= Ny) return;
   const i64 c = (ix * Ny + iy) * Nzp + iz;
   if ((mask[(ix * Ny + iy) * mwpr + (iz >> 5)] >> (iz & 31)) & 1u) return;
   double s = -(double)NN * (double)u1[c];
#pragma unroll
   for (int j = 0; j < NN; j++) s += (double)u1[c + off.o[j]];
   Lu[c] = (Real)(lfac * s);
}

// Lu = lfac*(-K*u + sum of reachable neighbours) on the boundary nodes       (sim_fdtd.py:737-770)
template <typename Real, int NN>
__global__ void k_lap_bn(const Real *__restrict__ u1, Real *__restrict__ Lu, const i64 *__restrict__ bn, const uint16_t *__restrict__ adj_bn,
                         i64 n, double lfac, Offsets off) {
   const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const i64 c = bn[i];
   const unsigned adj = adj_bn[i];
   double s = -(double)__popc(adj) * (double)u1[c];
#pragma unroll
   for (int j = 0; j < NN; j++)
      if ((adj >> j) & 1u) s += (double)u1[c + off.o[j]];
   Lu[c] = (Real)(lfac * s);
}

// partial[b] = sum over interior nodes of planes [1, Nx-2] of (u1-u2)^2/l2 - u1*Lu2    (sim_fdtd.py:838-841)
template <typename Real>
__global__ void __launch_bounds__(EN_THREADS) k_energy_int(const Real *__restrict__ u1, const Real *__restrict__ u2, const Real *__restrict__ Lu,
                                                           i64 Nx, i64 Ny, i64 Nz, i64 Nzp, double l2, double *__restrict__ partial) {
   double acc = 0.0;
   const i64 nrows = (Nx - 2) * (Ny - 2);
   for (i64 r = blockIdx.x; r < nrows; r += gridDim.x) {
      const i64 ix = 1 + r / (Ny - 2), iy = 1 + r % (Ny - 2);
      const i64 base = (ix * Ny + iy) * Nzp;
      for (i64 iz = 1 + threadIdx.x; iz < Nz - 1; iz += blockDim.x) {
         const double a = (double)u1[base + iz], b = (double)u2[base + iz];
         const double d = a - b;
         acc += (d * d) / l2 - a * (double)Lu[base + iz];
      }
   }
   const double t = en_block_sum(acc);
   if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// partial[b] = sum over the absorbing-shell nodes of (1-2^-Q)[(u1-u2)^2/l2 - u1*Lu2]     (sim_fdtd.py:592)
template <typename Real>
__global__ void __launch_bounds__(EN_THREADS) k_energy_abc_corr(const Real *__restrict__ u1, const Real *__restrict__ u2,
                                                                const Real *__restrict__ Lu, const i64 *__restrict__ bna,
                                                                const int8_t *__restrict__ Q, i64 n, double l2, double *__restrict__ partial) {
   double acc = 0.0;
   for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
      const i64 c = bna[i];
      const double a = (double)u1[c], b = (double)u2[c], d = a - b;
      acc += (1.0 - exp2(-(double)Q[i])) * ((d * d) / l2 - a * (double)Lu[c]);
   }
   const double t = en_block_sum(acc);
   if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// stored (which = 0): sum_i ssaf_i sum_m [v^2 D + (Ts g)^2 F]  with v = vh1, g = gh1        (sim_fdtd.py:848-849)
// lost   (which = 1): sum_i ssaf_i sum_m (v + vold)^2 E        with v = new vh1 (the reference's vh0)  (:852-853)
// state layout: st_idx (kernels.cuh); `v` = the state buffer, `g_or_vold` = the same buffer + 32 elements (g) or the copy of the
// pre-step buffer (its v); def = double [Nm][MMB][3] = D, E, F
template <typename Real, int MMB>
__global__ void __launch_bounds__(EN_THREADS) k_energy_branches(const Real *__restrict__ ssaf, const uint16_t *__restrict__ matmb,
                                                                const Real *__restrict__ v, const Real *__restrict__ g_or_vold, i64 Nbl,
                                                                i64 pitch, const double *__restrict__ def, double Ts, int which,
                                                                double *__restrict__ partial) {
   double acc = 0.0;
   for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < Nbl; i += (i64)gridDim.x * blockDim.x) {
      const unsigned mm = matmb[i];
      const int Mb = (int)(mm >> 8);
      const double *d = def + (i64)(mm & 0xffu) * MMB * 3;
      double s = 0.0;
      for (int m = 0; m < Mb; m++) {
         const double a = (double)v[st_idx<MMB>(m, i)], b = (double)g_or_vold[st_idx<MMB>(m, i)];
         if (which == 0) {
            const double tg = Ts * b;
            s += (a * a) * d[3 * m + 0] + (tg * tg) * d[3 * m + 2];
         } else {
            const double t = a + b;
            s += (t * t) * d[3 * m + 1];
         }
      }
      acc += (double)ssaf[i] * s;
   }
   const double t = en_block_sum(acc);
   if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// partial[b] = sum over the shell of (2^-Q * Q) (u0_new - u2ba)^2                            (sim_fdtd.py:614)
template <typename Real>
__global__ void __launch_bounds__(EN_THREADS) k_energy_abc_loss(const Real *__restrict__ u0, const Real *__restrict__ u2ba,
                                                                const i64 *__restrict__ bna, const int8_t *__restrict__ Q, i64 n,
                                                                double *__restrict__ partial) {
   double acc = 0.0;
   for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
      const double q = (double)Q[i], d = (double)u0[bna[i]] - (double)u2ba[i];
      acc += (exp2(-q) * q) * (d * d);
   }
   const double t = en_block_sum(acc);
   if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// one block: out = sum_s (u0_new[in_s] - u2in_s) * in_sigs[s, n]                              (sim_fdtd.py:618)
template <typename Real>
__global__ void __launch_bounds__(EN_THREADS) k_energy_in(const Real *__restrict__ u0, const Real *__restrict__ u2in, const i64 *__restrict__ in_ixyz,
                                                          const Real *__restrict__ in_row, i64 Ns, double *__restrict__ out) {
   double acc = 0.0;
   for (i64 s = threadIdx.x; s < Ns; s += blockDim.x) acc += ((double)u0[in_ixyz[s]] - (double)u2in[s]) * (double)in_row[s];
   const double t = en_block_sum(acc);
   if (threadIdx.x == 0) out[0] = t;
}

// one block: total of a partial array in a fixed order
__device__ __forceinline__ double en_total(const double *partial, int n) {
   double acc = 0.0;
   for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
   return en_block_sum(acc);
}

struct EnergyCoef {
   double V, h, c, l, l2;
};

// H_tot[n] from the three partial arrays, the reference's expression order (sim_fdtd.py:591-595)
__global__ void __launch_bounds__(EN_THREADS) k_energy_finish_H(const double *pA, const double *pB, const double *pC, int nA, int nB, int nC,
                                                                EnergyCoef k, double *__restrict__ H, i64 n) {
   const double A = en_total(pA, nA), B = en_total(pB, nB), C = en_total(pC, nC);
   if (threadIdx.x == 0) {
      double v = k.V * 0.5 * k.h * A;
      v -= k.V * 0.5 * k.h * B;
      v += k.V * 0.5 * k.c / k.l2 * C;
      H[n] = v;
   }
}

// E_lost[n+1], E_in[n+1] (sim_fdtd.py:612-618)
__global__ void __launch_bounds__(EN_THREADS) k_energy_finish_E(const double *pD, const double *pE, const double *pF, int nD, int nE,
                                                                EnergyCoef k, double *__restrict__ lost, double *__restrict__ ein, i64 n) {
   const double D = en_total(pD, nD), E = en_total(pE, nE);
   if (threadIdx.x == 0) {
      double v = lost[n] + k.V * 0.25 * k.h / k.l * D;
      v += 0.5 * k.V * k.h / k.l * E;
      lost[n + 1] = v;
      ein[n + 1] = ein[n] + (k.V * k.h / k.l2) * 0.5 * pF[0];
   }
}

}  // namespace pf
