// vox_core.h -- arithmetic of the voxeliser's hot stage, shared by the CUDA kernel (vox.cuh) and the host checker
// (oracle/vox_host.cpp).  Restates what VoxScene.calc_adj does for ONE grid point, ONE triangle, ONE direction
// (python/voxelizer/vox_scene.py:181-244 with common/tri_ray_intersection.py:67-104), operation for operation in double and
// without contraction, so that the boundary nodes, adjacencies and nearest-triangle choices come out identical:
//   * the point takes part when it lies in the triangle's bounding box grown by hf*(1+R_EPS) (and, on the FCC grid, has even
//     parity) and within hf*(1+R_EPS) of the triangle's plane (vox_scene.py:188-200);
//   * a ray from (point - vvh[k]) along ray_un[k] = normalise(uvv[k]) meets the triangle at distance t
//     (tri_ray_intersection_vec: coplanarity test on beta, t = unor.(cent - o)/beta, t >= 0, point-on-plane inside the three
//     outward edge functions up to d_eps = 1e-3 h), the hit distance is t - hf, "behind the point" beyond R_EPS*hf is no hit,
//     |t - hf| <= R_EPS*hf marks the point as lying ON the surface ("near boundary": all its links are cut);
//   * sums of three products are ((p0 + p1) + p2), as numpy's sum over the last axis of a 3-vector.
// and, for the stage before it (VoxGridBase.fill, python/voxelizer/vox_grid_base.py:67-176: which triangles meet which voxel),
// the Schwarz-Seidel triangle / box overlap test of common/tri_box_intersection.py:84-120 for ONE box and ONE triangle
// (pfv_tri_box).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PFV_HD __device__ __forceinline__
#else
#define PFV_HD static inline
#endif

struct VoxTri {
   double unor[3], cent[3], bmin[3], bmax[3], v[3][3], eab[3], ebc[3], eca[3];
};
struct VoxConst {
   double hf, c_bb /* hf*(1+R_EPS) */, c_near /* R_EPS*hf */, c_far /* (1+R_EPS)*hf */, d_eps /* 1e-3*h */, cp_eps /* 1e-6 */, neg_eps /* -DBL_EPSILON */;
};

#if defined(__CUDACC__)
#define PFV_MUL(a, b) __dmul_rn(a, b)
#define PFV_ADD(a, b) __dadd_rn(a, b)
#define PFV_SUB(a, b) __dsub_rn(a, b)
#define PFV_DIV(a, b) __ddiv_rn(a, b)
#else
// (the host checker is compiled with -ffp-contract=off)
#define PFV_MUL(a, b) ((a) * (b))
#define PFV_ADD(a, b) ((a) + (b))
#define PFV_SUB(a, b) ((a) - (b))
#define PFV_DIV(a, b) ((a) / (b))
#endif

PFV_HD double pfv_dot3(const double a0, const double a1, const double a2, const double b0, const double b1, const double b2) {
   return PFV_ADD(PFV_ADD(PFV_MUL(a0, b0), PFV_MUL(a1, b1)), PFV_MUL(a2, b2));
}

// vox_scene.py:188-200: bounding-box mask, then distance to the triangle's plane
PFV_HD bool pfv_point_near_plane(const VoxTri &t, const VoxConst &c, const double x, const double y, const double z) {
   if (!(x >= PFV_SUB(t.bmin[0], c.c_bb) && y >= PFV_SUB(t.bmin[1], c.c_bb) && z >= PFV_SUB(t.bmin[2], c.c_bb))) return false;
   if (!(x <= PFV_ADD(t.bmax[0], c.c_bb) && y <= PFV_ADD(t.bmax[1], c.c_bb) && z <= PFV_ADD(t.bmax[2], c.c_bb))) return false;
   const double dtp = pfv_dot3(t.unor[0], t.unor[1], t.unor[2], PFV_SUB(t.cent[0], x), PFV_SUB(t.cent[1], y), PFV_SUB(t.cent[2], z));
   return fabs(dtp) <= c.c_bb;
}

// tri_ray_intersection_vec for one ray: origin o = point - vvh[k], unit direction ru = normalise(uvv[k]); returns t or +inf
PFV_HD double pfv_ray_hit(const VoxTri &t, const VoxConst &c, const double ox, const double oy, const double oz, const double *ru) {
   double beta = pfv_dot3(ru[0], ru[1], ru[2], t.unor[0], t.unor[1], t.unor[2]);
   bool fail = fabs(beta) < c.cp_eps;
   if (fail) beta = c.neg_eps;
   const double tt = PFV_DIV(pfv_dot3(t.unor[0], t.unor[1], t.unor[2], PFV_SUB(t.cent[0], ox), PFV_SUB(t.cent[1], oy), PFV_SUB(t.cent[2], oz)), beta);
   fail = fail || tt < 0.0;
   const double px = PFV_ADD(ox, PFV_MUL(ru[0], tt)), py = PFV_ADD(oy, PFV_MUL(ru[1], tt)), pz = PFV_ADD(oz, PFV_MUL(ru[2], tt));
   // inside the three outward edge functions, measured from the edge midpoints
   const double (*v)[3] = t.v;
   fail = fail || pfv_dot3(PFV_SUB(px, PFV_MUL(0.5, PFV_ADD(v[0][0], v[1][0]))), PFV_SUB(py, PFV_MUL(0.5, PFV_ADD(v[0][1], v[1][1]))),
                           PFV_SUB(pz, PFV_MUL(0.5, PFV_ADD(v[0][2], v[1][2]))), t.eab[0], t.eab[1], t.eab[2]) > c.d_eps;
   fail = fail || pfv_dot3(PFV_SUB(px, PFV_MUL(0.5, PFV_ADD(v[1][0], v[2][0]))), PFV_SUB(py, PFV_MUL(0.5, PFV_ADD(v[1][1], v[2][1]))),
                           PFV_SUB(pz, PFV_MUL(0.5, PFV_ADD(v[1][2], v[2][2]))), t.ebc[0], t.ebc[1], t.ebc[2]) > c.d_eps;
   fail = fail || pfv_dot3(PFV_SUB(px, PFV_MUL(0.5, PFV_ADD(v[2][0], v[0][0]))), PFV_SUB(py, PFV_MUL(0.5, PFV_ADD(v[2][1], v[0][1]))),
                           PFV_SUB(pz, PFV_MUL(0.5, PFV_ADD(v[2][2], v[0][2]))), t.eca[0], t.eca[1], t.eca[2]) > c.d_eps;
   return fail ? (double)INFINITY : tt;
}

// vox_scene.py:216-222: hit distance of the point itself; *near = the point lies on the surface
PFV_HD double pfv_hit_dist(const VoxConst &c, const double t, bool *near) {
   double hd = PFV_SUB(t, c.hf);       // (+inf stays +inf)
   if (hd < -c.c_near) hd = (double)INFINITY;  // hits behind the point
   *near = fabs(hd) <= c.c_near;
   return *near ? fabs(hd) : hd;
}

// common/tri_box_intersection.py:84-120 (tri_box_intersection_vec) for one box [bbmin, bbmax] and one triangle given by its rows
// of tris_precompute (v: 3 corners, nor: the area-scaled normal, cent, and its bounding box).  The bounding-box rejection is also
// the candidate mask of vox_grid_base.py:110.
PFV_HD bool pfv_tri_box(const double *bbmin, const double *bbmax, const double *v /* [3][3] */, const double *nor, const double *cent,
                        const double *tbmin, const double *tbmax) {
   for (int j = 0; j < 3; j++)
      if (tbmin[j] > bbmax[j] || bbmin[j] > tbmax[j]) return false;
   double dp[3], c1[3], c2[3];
   for (int j = 0; j < 3; j++) {
      dp[j] = PFV_SUB(bbmax[j], bbmin[j]);
      const double c = nor[j] > 0.0 ? dp[j] : 0.0;
      c1[j] = PFV_SUB(c, cent[j]);
      c2[j] = PFV_SUB(PFV_SUB(dp[j], c), cent[j]);
   }
   // the triangle's plane passes through the box
   const double d1 = pfv_dot3(nor[0], nor[1], nor[2], c1[0], c1[1], c1[2]), d2 = pfv_dot3(nor[0], nor[1], nor[2], c2[0], c2[1], c2[2]);
   const double np = pfv_dot3(nor[0], nor[1], nor[2], bbmin[0], bbmin[1], bbmin[2]);
   if (PFV_MUL(PFV_ADD(np, d1), PFV_ADD(np, d2)) > 0.0) return false;
   // the three axis-aligned projections overlap
   for (int q = 0; q < 3; q++) {
      const int xq = q, yq = (q + 1) % 3, zq = (q + 2) % 3;
      for (int i = 0; i < 3; i++) {
         const double *va = v + 3 * i, *vb = v + 3 * ((i + 1) % 3);
         const double ex = PFV_SUB(vb[xq], va[xq]), ey = PFV_SUB(vb[yq], va[yq]);
         const double mx = PFV_MUL(0.5, PFV_ADD(vb[xq], va[xq])), my = PFV_MUL(0.5, PFV_ADD(vb[yq], va[yq]));
         double n0 = -ey, n1 = ex;
         if (nor[zq] < 0.0) n0 = -n0, n1 = -n1;
         const double dpx = PFV_MUL(dp[xq], n0), dpy = PFV_MUL(dp[yq], n1);
         const double de = PFV_ADD(PFV_ADD(-PFV_ADD(PFV_MUL(n0, mx), PFV_MUL(n1, my)), dpx > 0.0 ? dpx : 0.0), dpy > 0.0 ? dpy : 0.0);
         if (PFV_ADD(PFV_ADD(PFV_MUL(n0, bbmin[xq]), PFV_MUL(n1, bbmin[yq])), de) < 0.0) return false;
      }
   }
   return true;
}
