// TEST INFRASTRUCTURE ONLY -- a sequential host restatement of the voxeliser's hot stage (VoxScene.calc_adj,
// python/voxelizer/vox_scene.py:139-279) over the arithmetic shared with the CUDA kernel (pffdtd_b200/csrc/vox_core.h).  It lets
// the CPU test suite pin that arithmetic and the sequencing rules against golden vectors written by the unmodified reference
// (tests/golden/make_vox_fixtures.py) where no GPU is available; the product path is pffdtd_vox_run in libpffdtd_b200.so.
// Built by oracle/Makefile with -ffp-contract=off.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "pffdtd_b200.h"
#include "../pffdtd_b200/csrc/vox_core.h"

struct HostResult {
   std::vector<int64_t> bn;
   std::vector<uint8_t> adj;
   std::vector<int32_t> tidx;
   std::vector<double> ndist;
   int NN;
};

static VoxTri load_tri(const pffdtd_vox_desc *d, int32_t ti) {
   VoxTri t;
   for (int j = 0; j < 3; j++) {
      t.unor[j] = d->unor[3 * ti + j], t.cent[j] = d->cent[3 * ti + j], t.bmin[j] = d->bmin[3 * ti + j], t.bmax[j] = d->bmax[3 * ti + j];
      t.eab[j] = d->eab[3 * ti + j], t.ebc[j] = d->ebc[3 * ti + j], t.eca[j] = d->eca[3 * ti + j];
      for (int q = 0; q < 3; q++) t.v[q][j] = d->v[9 * ti + 3 * q + j];
   }
   return t;
}

extern "C" void *voxhost_run(const pffdtd_vox_desc *d) {
   if (!d || d->struct_size != (int32_t)sizeof(pffdtd_vox_desc)) return nullptr;
   HostResult *R = new HostResult();
   R->NN = d->NN;
   const VoxConst c{d->hf, d->c_bb, d->c_near, d->c_far, d->d_eps, d->cp_eps, -2.220446049250313e-16};
   const int NN = d->NN;
   for (int64_t vi = 0; vi < d->Nvox; vi++) {
      const int64_t *st = d->vox_start + 3 * vi, *sh = d->vox_shape + 3 * vi;
      const int64_t np = sh[0] * sh[1] * sh[2];
      std::vector<double> ndist((size_t)np, INFINITY), hd((size_t)np);
      std::vector<int32_t> tidx((size_t)np, -1);
      std::vector<uint16_t> cut((size_t)np, 0);  // bit k: link k is cut
      std::vector<uint8_t> bp((size_t)np, 0), nb((size_t)np, 0), m1((size_t)np);
      for (int64_t q = d->vox_tri_off[vi]; q < d->vox_tri_off[vi + 1]; q++) {
         const int32_t ti = d->vox_tri[q];
         const VoxTri t = load_tri(d, ti);
         bool any1 = false;
         for (int64_t p = 0; p < np; p++) {
            const int64_t iz = p % sh[2], iy = (p / sh[2]) % sh[1], ix = p / (sh[2] * sh[1]);
            const int64_t gx = st[0] + ix, gy = st[1] + iy, gz = st[2] + iz;
            const bool par = !d->fcc || ((gx + gy + gz) % 2 == 0);
            m1[(size_t)p] = par && pfv_point_near_plane(t, c, d->xv[gx], d->yv[gy], d->zv[gz]);
            any1 = any1 || m1[(size_t)p];
         }
         if (!any1) continue;  // vox_scene.py:193 / :201
         for (int k = 0; k < NN; k++) {
            bool anyk = false;
            for (int64_t p = 0; p < np; p++) {
               hd[(size_t)p] = INFINITY;
               if (!m1[(size_t)p]) continue;
               const int64_t iz = p % sh[2], iy = (p / sh[2]) % sh[1], ix = p / (sh[2] * sh[1]);
               const double x = d->xv[st[0] + ix], y = d->yv[st[1] + iy], z = d->zv[st[2] + iz];
               const double tt = pfv_ray_hit(t, c, PFV_SUB(x, d->vvh[3 * k]), PFV_SUB(y, d->vvh[3 * k + 1]), PFV_SUB(z, d->vvh[3 * k + 2]), d->ray_un + 3 * k);
               bool near;
               hd[(size_t)p] = pfv_hit_dist(c, tt, &near);
               if (near) nb[(size_t)p] = 1;
               anyk = anyk || hd[(size_t)p] <= c.hf;
            }
            if (!anyk) continue;  // vox_scene.py:224: decided for the whole voxel
            for (int64_t p = 0; p < np; p++) {
               const double h = hd[(size_t)p];
               if (!(h <= c.c_far)) continue;
               cut[(size_t)p] |= (uint16_t)(1u << k);
               bp[(size_t)p] = 1;
               if (h < ndist[(size_t)p]) ndist[(size_t)p] = h, tidx[(size_t)p] = ti;
            }
         }
      }
      for (int64_t p = 0; p < np; p++) {
         const int64_t iz = p % sh[2], iy = (p / sh[2]) % sh[1], ix = p / (sh[2] * sh[1]);
         const bool in = ix >= 1 && ix <= sh[0] - 2 && iy >= 1 && iy <= sh[1] - 2 && iz >= 1 && iz <= sh[2] - 2;
         if (!in) continue;
         uint16_t cu = cut[(size_t)p];
         if (nb[(size_t)p]) cu = (uint16_t)((1u << NN) - 1u);  // a point on the surface: every link cut (vox_scene.py:244)
         if (!cu) continue;
         R->bn.push_back(((st[0] + ix) * d->Ny + (st[1] + iy)) * d->Nz + (st[2] + iz));
         for (int k = 0; k < NN; k++) R->adj.push_back((cu >> k) & 1u ? 0 : 1);
         R->tidx.push_back(tidx[(size_t)p]);
         R->ndist.push_back(ndist[(size_t)p]);
      }
   }
   return R;
}
extern "C" int64_t voxhost_count(const void *h) { return h ? (int64_t)((const HostResult *)h)->bn.size() : -1; }
extern "C" int voxhost_read(const void *h, int64_t *bn, uint8_t *adj, int32_t *tidx, double *ndist) {
   const HostResult *R = (const HostResult *)h;
   if (!R) return -1;
   if (R->bn.empty()) return 0;
   memcpy(bn, R->bn.data(), R->bn.size() * 8);
   memcpy(adj, R->adj.data(), R->adj.size());
   memcpy(tidx, R->tidx.data(), R->tidx.size() * 4);
   memcpy(ndist, R->ndist.data(), R->ndist.size() * 8);
   return 0;
}
extern "C" void voxhost_free(void *h) { delete (HostResult *)h; }

// VoxGridBase.fill (python/voxelizer/vox_grid_base.py:67-176) sequentially: every voxel against every triangle, ascending
struct HostFill {
   std::vector<int64_t> off;
   std::vector<int32_t> tri;
};
extern "C" void *voxhost_fill(const pffdtd_voxfill_desc *d) {
   if (!d || d->struct_size != (int32_t)sizeof(pffdtd_voxfill_desc)) return nullptr;
   HostFill *R = new HostFill();
   R->off.assign((size_t)d->Nvox + 1, 0);
   for (int64_t vi = 0; vi < d->Nvox; vi++) {
      for (int64_t ti = 0; ti < d->Ntris; ti++)
         if (pfv_tri_box(d->vbmin + 3 * vi, d->vbmax + 3 * vi, d->v + 9 * ti, d->nor + 3 * ti, d->cent + 3 * ti, d->bmin + 3 * ti, d->bmax + 3 * ti))
            R->tri.push_back((int32_t)ti);
      R->off[(size_t)vi + 1] = (int64_t)R->tri.size();
   }
   return R;
}
extern "C" int64_t voxhost_fill_count(const void *h) { return h ? (int64_t)((const HostFill *)h)->tri.size() : -1; }
extern "C" int voxhost_fill_read(const void *h, int64_t *off, int32_t *tri) {
   const HostFill *R = (const HostFill *)h;
   if (!R) return -1;
   memcpy(off, R->off.data(), R->off.size() * 8);
   if (!R->tri.empty()) memcpy(tri, R->tri.data(), R->tri.size() * 4);
   return 0;
}
extern "C" void voxhost_fill_free(void *h) { delete (HostFill *)h; }
