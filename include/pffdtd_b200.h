/* pffdtd_b200.h -- C ABI of libpffdtd_b200.so, the B200-native replacement for the simulation
 * step of bsxfun/pffdtd.
 *
 * The reference has no plugin/FFI interface; its in-process contract for this path is
 *     double run_sim(const struct SimData *sd)        c_cuda/gpu_engine.h:665  (CPU: cpu_engine.h:52)
 * fed by load_sim_data()/scale_input() (c_cuda/fdtd_data.h:99,879) and drained by
 * rescale_output()/write_outputs() (fdtd_data.h:912,928); the Python twin is
 * SimEngine.run_steps(nstart,nsteps) (python/fdtd/sim_fdtd.py:529).  The entry points below are
 * what a binding of that contract would call.  Plain pointers and sizes only; every function
 * returns 0 on success or a negative PFFDTD_E* code (the reference aborts via assert/exit,
 * gpu_engine.h:192-200; here the message is kept in pffdtd_last_error()).
 *
 * Conventions (all the reference's, SURVEY.md "Index conventions"):
 *   grid Nx x Ny x Nz, z contiguous, linear index ii = ix*Ny*Nz + iy*Nz + iz;
 *   outermost layer on every axis is a halo; two pressure grids u1 (state n), u0 (state n-1,
 *   overwritten with n+1); receivers read u1; sources are added to the new state.
 *   Step order and arithmetic follow the C CPU engine exactly in both
 *   precisions, so fp64 AND fp32 results are bit-identical to cpu_engine.h.
 */
#ifndef PFFDTD_B200_H
#define PFFDTD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFFDTD_MMB 12 /* max RLC branches per material   (fdtd_data.h:33 MMb) */
#define PFFDTD_MNM 64 /* max number of materials         (fdtd_data.h:35 MNm) */

#define PFFDTD_OK 0
#define PFFDTD_EINVAL (-1)  /* bad argument / inconsistent description */
#define PFFDTD_ECUDA (-2)   /* CUDA runtime / driver error */
#define PFFDTD_ENCCL (-3)   /* NCCL error */
#define PFFDTD_ESTATE (-4)  /* call not valid in the current state */

/* Problem description handed to the engine: the fields of the reference's `struct SimData`
 * (fdtd_data.h:38-76) after load_sim_data()+scale_input(), for ONE slab of the grid.
 * Arrays are borrowed for the duration of pffdtd_create() only (copied to the device).
 * "Real" quantities (a1,a2,sl2,lo2, ssaf_bnl, mat_beta, mat_quads) travel as doubles that
 * already hold the value rounded to the working precision, so one layout serves fp32 and fp64. */
typedef struct pffdtd_desc {
   int32_t struct_size;  /* = sizeof(pffdtd_desc), ABI check */
   int32_t precision;    /* 1 = fp32, 2 = fp64                      (Makefile -DPRECISION, fdtd_common.h:43-71) */
   int32_t fcc_flag;     /* 0 Cartesian, 1 FCC checkerboard, 2 FCC folded (fdtd_data.h:162-165) */
   int32_t Nm;           /* number of materials                       (SimData.Nm) */
   int64_t Nx, Ny, Nz;   /* slab dims INCLUDING halo planes            (gpuHostData.Nxh, gpu_engine.h:755-760) */
   int64_t Nb, Nbl, Nba; /* boundary / lossy-boundary / ABC node counts in this slab */
   int64_t Ns, Nr, Nt;   /* source nodes, receiver nodes (this slab), time steps */
   double l, l2;         /* Courant number and its square            (SimData.l, l2) */
   double a1, a2;        /* stencil coefficients                     (fdtd_data.h:186-194) */
   double sl2, lo2;      /* (1+EPS)*lfac*l2 and l/2 */
   /* slab placement (gpu_engine.h:532-543, 784-823): global x index of local plane 0, and
    * whether local plane 0 / Nx-1 is a GLOBAL halo plane (mirror there) or a neighbour's plane */
   int64_t ix0;
   int32_t x_lo_edge, x_hi_edge;
   const int64_t *bn_ixyz;   /* [Nb]  slab-local linear indices, ascending not required */
   const uint16_t *adj_bn;   /* [Nb]  adjacency bits, bit j = neighbour j reachable (fdtd_data.h:532-538) */
   const int64_t *bnl_ixyz;  /* [Nbl] */
   const int8_t *mat_bnl;    /* [Nbl] material id */
   const double *ssaf_bnl;   /* [Nbl] */
   const int64_t *bna_ixyz;  /* [Nba] */
   const int8_t *Q_bna;      /* [Nba] 1 face, 2 edge, 3 corner */
   const int64_t *in_ixyz;   /* [Ns] */
   const int64_t *out_ixyz;  /* [Nr] */
   const double *in_sigs;    /* [Ns*Nt] row per source node, already scale_input()-scaled */
   const int8_t *Mb;         /* [Nm] branches per material */
   const double *mat_beta;   /* [Nm] */
   const double *mat_quads;  /* [Nm*PFFDTD_MMB*4] (b, bd, bDh, bFh) per (material, branch) (fdtd_data.h:79-84) */
} pffdtd_desc;

typedef struct pffdtd_engine pffdtd_engine; /* opaque */

/* last error message of the calling thread's most recent failing call */
const char *pffdtd_last_error(void);
/* library/version string, e.g. "pffdtd_b200 0.1 sm_100a" */
const char *pffdtd_version(void);

/* Allocate all simulation state on CUDA device `device` and upload the description.
 * Replaces the allocation half of run_sim (gpu_engine.h:739-974). */
int pffdtd_create(const pffdtd_desc *desc, int device, pffdtd_engine **out);
int pffdtd_destroy(pffdtd_engine *e);

/* Multi-GPU (one process per GPU): rank 0 obtains an id, the host shares it, every rank joins.
 * Neighbour ranks exchange one Ny*Nz plane per side per step (replaces the
 * cudaMemcpyPeerAsync waves of gpu_engine.h:1086-1126) with ncclSend/ncclRecv, overlapped
 * with the interior update. */
int pffdtd_comm_unique_id(void *id128 /* 128 bytes out */);
int pffdtd_comm_init(pffdtd_engine *e, const void *id128, int rank, int nranks);
/* The same exchange over PEER MEMORY instead of NCCL (GPUs of one node): every rank exports a blob (CUDA IPC handles of its two
 * grids and of two flag words), the host hands each rank the blobs of its lower / upper neighbour (NULL at the ends of the grid),
 * and from then on a step copies its edge planes straight into the neighbours' halo planes, raises their flag from the device,
 * and starts by waiting for its own flags on the device.  No library call in the step, and the whole step replays from CUDA
 * graphs again.  Falls back to the communicator of pffdtd_comm_init when not connected. */
#define PFFDTD_PEER_BLOB 256
int pffdtd_peer_export(pffdtd_engine *e, void *blob /* PFFDTD_PEER_BLOB bytes out */);
int pffdtd_peer_connect(pffdtd_engine *e, const void *blob_lo, const void *blob_hi);

/* Options: "air_kernel" (0 generic one-thread-per-node, 1 tiled TMA sweep [default for Cartesian grids]),
 * "fuse" (1 = the tiled kernel mirrors the halos on write and the absorbing shell is finished from stashed values, -1 = the layout's
 * default [7-point: on; 13-point: off]),
 * "overlap" (1 = edge planes first, halo exchange overlapped with the interior [default]),
 * "air_cfg" (tile configuration), "air_xc" (x-chunk length), "profile_air" (CUDA events around every air launch),
 * "manual_halo" (allow stepping a slab without a communicator; the caller moves the halo planes),
 * "use_graph" (1 = replay captured steps as CUDA graphs [default]), "svc" (1 = the air kernel's service warp finishes the sparse
 * rigid-boundary nodes and the z faces of the absorbing shell from shared memory, -1 = the layout's default [7-point: on where the
 * grid allows it; 13-point: off]), "fd_bulk" (1 = the branch kernel moves its state with TMA bulk copies [0]), "svc_cap"
 * (tile-planes with more boundary nodes than this leave them to the list kernel [64, at most 192 minus two per tile row]), "abc_overlap" (1 = the absorbing-shell kernel runs beside the boundary kernels when no
 * boundary / source node lies on the shell [default]), "zflip" (1 = in the unfused step the absorbing-shell kernel also writes the z
 * halos of the new state, so the next step's mirror pass needs no z kernel; only where no boundary / source node sits at z = 2 or
 * Nz-3 and the shell list is the canonical one [default]). */
int pffdtd_set_option(pffdtd_engine *e, const char *key, int64_t value);
/* Counters/timers: "launches", "steps", "air_ms" / "air_launches_timed" (CUDA-event time of the air launches
 * since reset, with profile_air), "timer_start" / "timer_stop_ms" (device stopwatch on the engine's stream),
 * "fused", "mirror_pairs", "air_kernel", "air_cfg" / "air_lanes_z" (tile configuration in use: 32, 16 or 8 lanes of a warp along z),
 * "abc_disjoint", "Nzp", "energy", "svc" (service-warp lists in use), "svc_entries", "nb_left" (boundary nodes left to k_rigid), "zflip" (the unfused step writes the z halos in k_abc). */
int pffdtd_get_stat(pffdtd_engine *e, const char *key, double *out);
int pffdtd_reset_stats(pffdtd_engine *e);

/* Advance time steps nstart .. nstart+nsteps-1 (the hot loop, gpu_engine.h:993-1170 /
 * sim_fdtd.py:529).  Asynchronous on the engine's streams; source samples come from the
 * uploaded in_sigs, receiver samples accumulate on the device. */
int pffdtd_run_steps(pffdtd_engine *e, int64_t nstart, int64_t nsteps);
/* One step with HOST buffers: in_samples[Ns] (this step's source samples, NULL = use uploaded
 * in_sigs) are copied to the device, the step runs, and the step's receiver samples
 * out_samples[Nr] are copied back; blocks until done (the per-step D2H of gpu_engine.h:1059-1075).
 * On one GPU the two copies and the step's kernels replay as ONE captured CUDA graph (fixed staging buffers). */
int pffdtd_step_host(pffdtd_engine *e, int64_t n, const double *in_samples, double *out_samples);
/* Block until all queued work is complete. */
int pffdtd_sync(pffdtd_engine *e);

/* Receiver traces for steps [n0,n1): u_out[nr*(n1-n0) + (n-n0)], device (sorted) receiver order,
 * widened to double as the reference does (gpu_engine.h:1072). */
int pffdtd_read_outputs(pffdtd_engine *e, int64_t n0, int64_t n1, double *u_out);
/* Copy a whole pressure grid to the host in the reference's unpadded layout
 * (which: 1 = u1 current state, 0 = u0), widened to double.  For energy checks and plots
 * (sim_fdtd.py:640-658 gather_slice). */
int pffdtd_read_grid(pffdtd_engine *e, int which, double *out /* Nx*Ny*Nz */);
/* Overwrite a pressure grid from the host (initial conditions for energy tests). */
int pffdtd_write_grid(pffdtd_engine *e, int which, const double *in /* Nx*Ny*Nz */);
/* Boundary ODE state (vh1, gh1: [Nbl*PFFDTD_MMB], reference CPU layout nb*MMb+m) for energy checks. */
int pffdtd_read_boundary_state(pffdtd_engine *e, double *vh1, double *gh1);

/* Energy balance of the reference's Python engine (python/fdtd/sim_fdtd.py:587-620 with --energy; SURVEY.md
 * App. F), evaluated on the device.  Enabling it (before the first step) allocates a third grid for the
 * boundary-aware Laplacian of the previous state and makes every step also accumulate
 *   H_tot[n] (stored energy at step n), E_lost[n+1] (cumulative boundary + absorbing-shell losses),
 *   E_in[n+1] (cumulative energy injected by the sources),
 * all in double, in a fixed summation order.  The invariant is H_tot[n] + E_lost[n] == E_in[n] to rounding.
 * With slabs each rank holds the sums over its own planes; the host adds the ranks.  Energy steps use the
 * unfused kernels (same receiver traces bit for bit).  Not available for folded FCC grids (fcc_flag 2), which
 * the reference's Python engine cannot load either. */
typedef struct pffdtd_energy_desc {
   int32_t struct_size;   /* = sizeof(pffdtd_energy_desc) */
   int32_t reserved;
   double h, c, Ts;       /* grid spacing, speed of sound, time step (sim_consts.h5: h, c, Ts) */
   const double *mat_DEF; /* [Nm*PFFDTD_MMB*3] raw (D, E, F) per (material, branch), zero padded (sim_mats.h5 mat_XX_DEF) */
} pffdtd_energy_desc;
int pffdtd_energy_enable(pffdtd_engine *e, const pffdtd_energy_desc *d);
/* H_tot[Nt], E_lost[Nt+1], E_in[Nt+1]; entries of steps not run yet are 0 */
int pffdtd_read_energy(pffdtd_engine *e, double *H_tot, double *E_lost, double *E_in);

/* Device self-test: the fused absorbing-shell update divides by the constants 1 + l*Q with a
 * reciprocal + one exact-residual correction instead of a general division; this compares the two on
 * `count` pseudo-random numerators per divisor and returns the number of differing results (must be 0). */
int pffdtd_selftest(int device, double l, int precision, int64_t count, int64_t *mismatches);

/* Diagnostic (no device needed): the x-chunk plan the tiled air kernel's work queue uses for a job of `n_planes` planes with
 * chunk length `xc`: writes the nch+1 chunk bounds (bounds[0] = 0 ... bounds[nch] = n_planes) and returns nch, or a negative
 * PFFDTD_E* code.  The plan is arithmetic (any number of planes) with a guided tail of halving chunk lengths. */
int64_t pffdtd_air_chunk_plan(int64_t n_planes, int xc, int64_t *bounds, int64_t max_bounds);

/* Whole-run convenience with the reference's run_sim() shape: create, run Nt steps, write
 * u_out[Nr*Nt], destroy; returns elapsed seconds of the loop through *elapsed_s. */
int pffdtd_run_sim(const pffdtd_desc *desc, int device, double *u_out, double *elapsed_s);

/* ---- Multi-GPU from ONE host thread, the reference's own model: run_sim counts the visible devices (gpu_engine.h:679-691;
 * CUDA_VISIBLE_DEVICES selects them), split_data cuts the grid into x-slabs (:516-662), the step loop walks the devices (:994) and
 * the halo planes move by peer copies (:1086-1126).  Here: one engine per slab, the edge planes of a slab are computed first and pushed
 * into the neighbours' halo planes (cudaMemcpyPeerAsync over NVLink) on a second stream while the interior runs; the neighbour's next
 * step waits on an event -- no host synchronisation in the loop.  `desc` describes the WHOLE grid with sorted node lists (what the
 * reference requires of a multi-GPU folder, :497-513).  nslabs <= 0: one slab per visible device.  devices = NULL: slab r on device
 * r % visible (several slabs may share a device).  balance != 0: slabs of about equal cost (air nodes + weighted boundary / lossy /
 * shell nodes per plane) instead of the reference's equal-plane split; the results are the same bits either way. */
typedef struct pffdtd_multi pffdtd_multi; /* opaque */
int pffdtd_multi_create(const pffdtd_desc *desc, int nslabs, const int *devices, int balance, pffdtd_multi **out);
int pffdtd_multi_destroy(pffdtd_multi *m);
/* Diagnostic (no device needed): first owned plane and number of owned planes of every slab as pffdtd_multi_create would cut `desc` */
int pffdtd_slab_plan(const pffdtd_desc *desc, int nslabs, int balance, int64_t *starts, int64_t *sizes);
/* number of slabs; planes[r] = owned planes of slab r (up to `max` entries) */
int pffdtd_multi_slabs(pffdtd_multi *m, int64_t *planes, int max);
/* the engine of one slab, for pffdtd_set_option / pffdtd_get_stat / pffdtd_read_grid (do not step or destroy it directly) */
int pffdtd_multi_engine(pffdtd_multi *m, int slab, pffdtd_engine **e);
int pffdtd_multi_run_steps(pffdtd_multi *m, int64_t nstart, int64_t nsteps);
int pffdtd_multi_sync(pffdtd_multi *m);
/* receiver traces of the whole grid, sorted receiver order (rows of the slabs in slab order): u_out[nr*(n1-n0) + (n-n0)] */
int pffdtd_multi_read_outputs(pffdtd_multi *m, int64_t n0, int64_t n1, double *u_out);
/* run_sim() over every visible device (nslabs <= 0) or `nslabs` slabs: create, run Nt steps, write u_out[Nr*Nt], destroy */
int pffdtd_run_sim_multi(const pffdtd_desc *desc, int nslabs, const int *devices, double *u_out, double *elapsed_s);

/* ---- the voxeliser's hot stage on the GPU.  VoxScene.calc_adj (python/voxelizer/vox_scene.py:95-440) casts, for
 * every grid point near a triangle, a ray towards each of its 6 (12) neighbours and cuts the link where the ray meets the surface within
 * one grid step; the reference does it voxel by voxel in numpy over a process pool.  Here one thread block takes one voxel of the
 * reference's voxel grid (same voxels, same per-voxel triangle lists, same order of triangles and directions, the same arithmetic in
 * double -- pffdtd_b200/csrc/vox_core.h), so bn_ixyz / adj_bn and the nearest triangle of every boundary node come out identical.
 * Everything calc_adj reads travels in the descriptor (host arrays, borrowed for the call). */
typedef struct pffdtd_vox_desc {
   int32_t struct_size; /* = sizeof(pffdtd_vox_desc) */
   int32_t NN;          /* 6 Cartesian, 12 FCC */
   int32_t fcc;         /* FCC: only points of even parity take part (vox_scene.py:176-179) */
   int32_t reserved;
   int64_t Nx, Ny, Nz;
   const double *xv, *yv, *zv;     /* grid coordinates (cart_grid) */
   double hf, c_bb, c_near, c_far; /* hf; hf*(1+R_EPS); R_EPS*hf; (1+R_EPS)*hf   (vox_scene.py:60, 188-228) */
   double d_eps, cp_eps;           /* 1e-3*h; 1e-6                              (vox_scene.py:213, tri_ray_intersection.py:67) */
   const double *vvh;              /* [NN][3] h * direction                     (vox_scene.py:85) */
   const double *ray_un;           /* [NN][3] normalise(uvv[k])                 (tri_ray_intersection.py:77) */
   int64_t Nvox;                   /* non-empty voxels, in the reference's order (vox_grid.nonempty_idx) */
   const int64_t *vox_start;       /* [Nvox][3] ixyz_start */
   const int64_t *vox_shape;       /* [Nvox][3] Nhxyz (points, halo layer included) */
   const int64_t *vox_tri_off;     /* [Nvox+1] into vox_tri */
   const int32_t *vox_tri;         /* triangle indices of every voxel, in the voxel's own order */
   int64_t Ntris;
   const double *unor, *cent, *bmin, *bmax; /* [Ntris][3]    (tris_precompute.py) */
   const double *v;                         /* [Ntris][3][3] */
   const double *eab, *ebc, *eca;           /* [Ntris][3] outward unit edge normals */
} pffdtd_vox_desc;
typedef struct pffdtd_vox pffdtd_vox; /* opaque: the result of one run */
int pffdtd_vox_run(const pffdtd_vox_desc *d, int device, pffdtd_vox **out); /* env PFFDTD_VOX_TIMING=1: phase times on stderr */
int64_t pffdtd_vox_count(const pffdtd_vox *r); /* boundary nodes found */
/* bn_ixyz[Nb] (voxel by voxel in the reference's order, ascending inside a voxel), adj[Nb][NN] (1 = link open), tidx[Nb] (nearest
 * triangle), ndist[Nb] (its hit distance) */
int pffdtd_vox_read(const pffdtd_vox *r, int64_t *bn_ixyz, uint8_t *adj, int32_t *tidx, double *ndist);
int pffdtd_vox_free(pffdtd_vox *r);

/* The stage before it: VoxGridBase.fill (python/voxelizer/vox_grid_base.py:67-176) decides which triangles meet which voxel with the
 * Schwarz-Seidel triangle / box test of common/tri_box_intersection.py:84-120, voxel by voxel in numpy (minutes at production size).
 * Here one warp takes one voxel and scans the triangle table; the lists come out in ascending triangle order, as the reference's. */
typedef struct pffdtd_voxfill_desc {
   int32_t struct_size; /* = sizeof(pffdtd_voxfill_desc) */
   int32_t reserved;
   int64_t Nvox;                /* every voxel of the grid (vox_grid.voxels) */
   const double *vbmin, *vbmax; /* [Nvox][3] the voxels' boxes (vox_grid.py:128-129) */
   int64_t Ntris;
   const double *v;                        /* [Ntris][3][3]  (tris_precompute.py) */
   const double *nor, *cent, *bmin, *bmax; /* [Ntris][3] area-scaled normal, centroid, bounding box */
} pffdtd_voxfill_desc;
typedef struct pffdtd_voxfill pffdtd_voxfill;
int pffdtd_voxfill_run(const pffdtd_voxfill_desc *d, int device, pffdtd_voxfill **out);
int64_t pffdtd_voxfill_count(const pffdtd_voxfill *r); /* total length of the lists */
int pffdtd_voxfill_read(const pffdtd_voxfill *r, int64_t *off /* [Nvox+1] */, int32_t *tri /* [count] */);
int pffdtd_voxfill_free(pffdtd_voxfill *r);

#ifdef __cplusplus
}
#endif
#endif /* PFFDTD_B200_H */
