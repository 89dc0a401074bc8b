// vox.cuh -- the voxeliser's hot stage (VoxScene.calc_adj, python/voxelizer/vox_scene.py:139-279) on the GPU.
//
// One thread block per non-empty voxel of the reference's voxel grid; the block's threads stride over the voxel's grid points.
// Triangles are taken in the voxel's own order and directions in order, as the reference does: the choice of the nearest triangle
// (strict "<": the first one wins a tie) and the per-voxel early-outs depend on it.  Two of those early-outs couple the points of a
// voxel -- a triangle is skipped when NO point of the voxel is near its plane, and a (triangle, direction) pair is applied only when
// SOME point of the voxel has a hit within hf (vox_scene.py:193/201, :224) -- they are block-wide votes (__syncthreads_or).  The hit
// distance of a point is recomputed after the vote instead of being parked in memory (about 80 double-precision operations against
// 16 bytes of traffic).  All arithmetic is vox_core.h's, shared with the host checker (oracle/vox_host.cpp).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "vox_core.h"

namespace pf {

struct VoxArgs {
   VoxConst c;
   int NN, fcc;
   long long Ny, Nz;
   const double *xv, *yv, *zv, *vvh, *ray_un;
   const long long *vox_start, *vox_shape, *vox_tri_off, *pt_off;
   const int *vox_tri;
   const double *unor, *cent, *bmin, *bmax, *v, *eab, *ebc, *eca;
   // per point of every voxel (halo layer included): nearest hit, its triangle, cut links, flags (bit 0 boundary point,
   // bit 1 lies on the surface, bit 2 scratch: near the current triangle's plane)
   double *ndist;
   int *tidx;
   unsigned short *cut;
   unsigned char *fl;
   // compaction (vox_scene.py:246-279, 343-366): the boundary points of every voxel's interior, voxel by voxel, ascending in a voxel
   long long *count;      // [Nvox]
   const long long *off;  // [Nvox + 1]
   long long *o_bn;
   unsigned char *o_adj;  // [Nb][NN], 1 = link open
   int *o_tidx;
   double *o_ndist;
};

// cut links of point p of a voxel if it is a boundary point of the voxel's interior, else 0; a point lying on the surface has every
// link cut (vox_scene.py:244)
__device__ __forceinline__ unsigned vox_selected(const int NN, const int p, const long long sx, const long long sy, const long long sz,
                                                 const unsigned short *cut, const unsigned char *fl) {
   const long long iz = p % sz, iy = (p / sz) % sy, ix = p / (sz * sy);
   if (ix < 1 || ix > sx - 2 || iy < 1 || iy > sy - 2 || iz < 1 || iz > sz - 2) return 0u;
   return (fl[p] & 2u) ? (1u << NN) - 1u : (unsigned)cut[p];
}

__global__ void __launch_bounds__(256) k_vox_calc_adj(const VoxArgs a) {
   const long long vi = blockIdx.x;
   const long long sx = a.vox_shape[3 * vi], sy = a.vox_shape[3 * vi + 1], sz = a.vox_shape[3 * vi + 2];
   const long long gx0 = a.vox_start[3 * vi], gy0 = a.vox_start[3 * vi + 1], gz0 = a.vox_start[3 * vi + 2];
   const int np = (int)(sx * sy * sz);
   const long long base = a.pt_off[vi];
   double *ndist = a.ndist + base;
   int *tidx = a.tidx + base;
   unsigned short *cut = a.cut + base;
   unsigned char *fl = a.fl + base;
   __shared__ VoxTri t;
   __shared__ int tri_id;
   for (int p = threadIdx.x; p < np; p += blockDim.x) ndist[p] = (double)INFINITY, tidx[p] = -1, cut[p] = 0, fl[p] = 0;
   const VoxConst c = a.c;
   for (long long q = a.vox_tri_off[vi]; q < a.vox_tri_off[vi + 1]; q++) {
      __syncthreads();  // everybody is done with the previous triangle
      if (threadIdx.x < 3) {
         const int ti = a.vox_tri[q], j = threadIdx.x;
         t.unor[j] = a.unor[3 * ti + j], t.cent[j] = a.cent[3 * ti + j], t.bmin[j] = a.bmin[3 * ti + j], t.bmax[j] = a.bmax[3 * ti + j];
         t.eab[j] = a.eab[3 * ti + j], t.ebc[j] = a.ebc[3 * ti + j], t.eca[j] = a.eca[3 * ti + j];
         for (int w = 0; w < 3; w++) t.v[w][j] = a.v[9 * ti + 3 * w + j];
         if (j == 0) tri_id = ti;
      }
      __syncthreads();
      int any1 = 0;
      for (int p = threadIdx.x; p < np; p += blockDim.x) {
         const long long iz = p % sz, iy = (p / sz) % sy, ix = p / (sz * sy);
         const long long gx = gx0 + ix, gy = gy0 + iy, gz = gz0 + iz;
         const bool par = !a.fcc || (((gx + gy + gz) & 1) == 0);
         const bool m = par && pfv_point_near_plane(t, c, a.xv[gx], a.yv[gy], a.zv[gz]);
         fl[p] = (unsigned char)((fl[p] & 3u) | (m ? 4u : 0u));
         any1 |= m ? 1 : 0;
      }
      if (!__syncthreads_or(any1)) continue;  // vox_scene.py:193 / :201
      const int ti = tri_id;
      for (int k = 0; k < a.NN; k++) {
         const double vx = a.vvh[3 * k], vy = a.vvh[3 * k + 1], vz = a.vvh[3 * k + 2];
         const double ru[3] = {a.ray_un[3 * k], a.ray_un[3 * k + 1], a.ray_un[3 * k + 2]};
         int anyk = 0;
         for (int p = threadIdx.x; p < np; p += blockDim.x) {
            if (!(fl[p] & 4u)) continue;
            const long long iz = p % sz, iy = (p / sz) % sy, ix = p / (sz * sy);
            const double tt = pfv_ray_hit(t, c, PFV_SUB(a.xv[gx0 + ix], vx), PFV_SUB(a.yv[gy0 + iy], vy), PFV_SUB(a.zv[gz0 + iz], vz), ru);
            bool near;
            const double hd = pfv_hit_dist(c, tt, &near);
            if (near) fl[p] |= 2u;
            anyk |= hd <= c.hf ? 1 : 0;
         }
         if (!__syncthreads_or(anyk)) continue;  // vox_scene.py:224: decided for the whole voxel
         for (int p = threadIdx.x; p < np; p += blockDim.x) {
            if (!(fl[p] & 4u)) continue;
            const long long iz = p % sz, iy = (p / sz) % sy, ix = p / (sz * sy);
            const double tt = pfv_ray_hit(t, c, PFV_SUB(a.xv[gx0 + ix], vx), PFV_SUB(a.yv[gy0 + iy], vy), PFV_SUB(a.zv[gz0 + iz], vz), ru);
            bool near;
            const double hd = pfv_hit_dist(c, tt, &near);
            if (!(hd <= c.c_far)) continue;
            cut[p] |= (unsigned short)(1u << k);
            fl[p] |= 1u;
            if (hd < ndist[p]) ndist[p] = hd, tidx[p] = ti;
         }
      }
   }
   // how many boundary points this voxel contributes
   __shared__ int s_cnt;
   if (threadIdx.x == 0) s_cnt = 0;
   __syncthreads();
   int cnt = 0;
   for (int p = threadIdx.x; p < np; p += blockDim.x) cnt += vox_selected(a.NN, p, sx, sy, sz, cut, fl) != 0u;
   if (cnt) atomicAdd(&s_cnt, cnt);
   __syncthreads();
   if (threadIdx.x == 0) a.count[vi] = s_cnt;
}

// second pass: the boundary points of voxel vi go to rows off[vi].. of the result, in ascending point order
__global__ void __launch_bounds__(256) k_vox_emit(const VoxArgs a) {
   const long long vi = blockIdx.x;
   const long long sx = a.vox_shape[3 * vi], sy = a.vox_shape[3 * vi + 1], sz = a.vox_shape[3 * vi + 2];
   const long long gx0 = a.vox_start[3 * vi], gy0 = a.vox_start[3 * vi + 1], gz0 = a.vox_start[3 * vi + 2];
   const int np = (int)(sx * sy * sz);
   const long long base = a.pt_off[vi];
   if (a.off[vi + 1] == a.off[vi]) return;
   __shared__ int wsum[8];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   long long run = a.off[vi];
   for (int p0 = 0; p0 < np; p0 += 256) {
      const int p = p0 + threadIdx.x;
      const unsigned cu = p < np ? vox_selected(a.NN, p, sx, sy, sz, a.cut + base, a.fl + base) : 0u;
      const unsigned m = __ballot_sync(0xffffffffu, cu != 0u);
      if (lane == 0) wsum[w] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
      for (int j = 0; j < 8; j++) {
         before += j < w ? wsum[j] : 0;
         total += wsum[j];
      }
      if (cu) {
         const long long r = run + before + __popc(m & ((1u << lane) - 1u));
         const long long iz = p % sz, iy = (p / sz) % sy, ix = p / (sz * sy);
         a.o_bn[r] = ((gx0 + ix) * a.Ny + (gy0 + iy)) * a.Nz + (gz0 + iz);
         for (int k = 0; k < a.NN; k++) a.o_adj[r * a.NN + k] = (cu >> k) & 1u ? 0 : 1;
         a.o_tidx[r] = a.tidx[base + p];
         a.o_ndist[r] = a.ndist[base + p];
      }
      run += total;
      __syncthreads();
   }
}

// VoxGridBase.fill (vox_grid_base.py:67-176): which triangles meet which voxel.  One warp per voxel; its lanes take 32 consecutive
// triangles at a time, so that the hits of a voxel come out in ascending triangle order (the order calc_adj later depends on) from a
// ballot and a population count.  Run twice: WRITE = false counts, WRITE = true fills the lists at the offsets the counts gave.
struct VoxFillArgs {
   long long Nvox, Ntris;
   const double *vbmin, *vbmax, *v, *nor, *cent, *bmin, *bmax;
   long long *count;      // [Nvox]
   const long long *off;  // [Nvox + 1]
   int *tri;
};

template <bool WRITE>
__global__ void __launch_bounds__(256) k_vox_fill(const VoxFillArgs a) {
   const long long vi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   if (vi >= a.Nvox) return;
   const int lane = threadIdx.x & 31;
   double lo[3], hi[3];
   for (int j = 0; j < 3; j++) lo[j] = a.vbmin[3 * vi + j], hi[j] = a.vbmax[3 * vi + j];
   long long run = WRITE ? a.off[vi] : 0;
   for (long long t0 = 0; t0 < a.Ntris; t0 += 32) {
      const long long ti = t0 + lane;
      const bool hit = ti < a.Ntris && pfv_tri_box(lo, hi, a.v + 9 * ti, a.nor + 3 * ti, a.cent + 3 * ti, a.bmin + 3 * ti, a.bmax + 3 * ti);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (WRITE && hit) a.tri[run + __popc(m & ((1u << lane) - 1u))] = (int)ti;
      run += __popc(m);
   }
   if (!WRITE && lane == 0) a.count[vi] = run;
}

}  // namespace pf
