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
};

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
