/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product path.
 *
 * Driver that compiles the UNMODIFIED reference CPU engine from where it lies under
 * /root/reference/c_cuda (helper_funcs.h, fdtd_common.h, fdtd_data.h, cpu_engine.h) into a
 * shared library, oracle/_ref/libpffdtd_ref_{f32,f64}.so (see oracle/Makefile).
 *
 * The reference loader (fdtd_data.h:99 load_sim_data) talks to libhdf5, which is absent in this
 * image.  The 13 H5* entry points it uses are implemented here over an in-memory dataset
 * registry that the Python harness fills (refdrv_put) from the very same .h5 files, read with
 * the repo's own HDF5 reader.  Everything after H5Dread -- coefficient derivation, bit packing,
 * lossy-node compaction, ABC list, scale_input, run_sim (cpu_engine.h:52), rescale_output,
 * write_outputs -- is the reference's code, untouched.
 *
 * Uses: (1) pin oracle/pffdtd_oracle.c (the restatement) bit-exactly; (2) validate the product's
 * host prep (pffdtd_b200/sim_data.py) field by field; (3) generate tests/golden/ fixtures;
 * (4) the CPU baseline arm of bench.py ("kind": "reference").
 */
#ifndef _DEFAULT_SOURCE
#define _DEFAULT_SOURCE
#endif
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <assert.h>
#include <stdbool.h>
#include <math.h>
#include <omp.h>

#include "hdf5.h" /* oracle/stub/hdf5.h */

/* ---- the unmodified reference sources ---- */
#include <helper_funcs.h>
#include <fdtd_common.h>
#include <fdtd_data.h>
#if defined(USING_B200) && USING_B200
#include <b200_engine.h> /* integration/b200_engine.h: the reference's main sequence with run_sim served by libpffdtd_b200.so (tests) */
#elif USING_CUDA
#include <gpu_engine.h> /* the reference's own CUDA engine, a PERFORMANCE COMPARATOR only (tests/diag/compare_reference_gpu_engine.py); built by nvcc -x cu */
#else
#include <cpu_engine.h>
#endif

/* nvcc compiles this file as C++ (the reference's Makefile does the same to fdtd_main.c): keep the entry points unmangled */
#ifdef __cplusplus
#define REFDRV_API extern "C"
#else
#define REFDRV_API
#endif

/* ------------------------------------------------------------------------------------------
 * in-memory dataset registry
 * ---------------------------------------------------------------------------------------- */
enum { DT_F64 = 0, DT_I64 = 1, DT_I8 = 2 };

struct dset_rec {
   char file[64];
   char name[64];
   int dtype;
   int ndims;
   hsize_t dims[4];
   uint64_t nbytes;
   void *data;
};

#define MAX_DSETS 512
static struct dset_rec g_dsets[MAX_DSETS];
static int g_ndsets = 0;

#define MAX_FILES 16
static char g_files[MAX_FILES][64];
static int g_nfiles = 0;

#define MAX_SPACES 64
static struct { int ndims; hsize_t dims[4]; } g_spaces[MAX_SPACES];
static int g_nspaces = 0;

static int dt_size(int dt) { return dt == DT_I8 ? 1 : 8; }

static struct dset_rec *find_dset(const char *file, const char *name) {
   for (int i = 0; i < g_ndsets; i++) {
      if (strcmp(g_dsets[i].file, file) == 0 && strcmp(g_dsets[i].name, name) == 0) return &g_dsets[i];
   }
   return NULL;
}

REFDRV_API void refdrv_clear(void) {
   for (int i = 0; i < g_ndsets; i++) free(g_dsets[i].data);
   g_ndsets = 0;
   g_nfiles = 0;
   g_nspaces = 0;
}

REFDRV_API int refdrv_put(const char *file, const char *name, int dtype, int ndims, const int64_t *dims, const void *data) {
   if (g_ndsets >= MAX_DSETS || ndims > 4) return -1;
   struct dset_rec *r = find_dset(file, name);
   if (r == NULL) r = &g_dsets[g_ndsets++];
   else free(r->data);
   snprintf(r->file, sizeof r->file, "%s", file);
   snprintf(r->name, sizeof r->name, "%s", name);
   r->dtype = dtype;
   r->ndims = ndims;
   uint64_t n = 1;
   for (int d = 0; d < ndims; d++) { r->dims[d] = (hsize_t)dims[d]; n *= (uint64_t)dims[d]; }
   r->nbytes = n * dt_size(dtype);
   r->data = malloc(r->nbytes ? r->nbytes : 1);
   memcpy(r->data, data, r->nbytes);
   return 0;
}

/* ---- the H5 shim ---- */
hid_t H5Fopen(const char *filename, unsigned flags, hid_t fapl) {
   (void)flags; (void)fapl;
   assert(g_nfiles < MAX_FILES);
   snprintf(g_files[g_nfiles], 64, "%s", filename);
   return 1000 + g_nfiles++;
}
hid_t H5Fcreate(const char *filename, unsigned flags, hid_t fcpl, hid_t fapl) {
   (void)fcpl;
   return H5Fopen(filename, flags, fapl);
}
herr_t H5Fclose(hid_t file) { (void)file; return 0; }

hid_t H5Dopen(hid_t file, const char *name, hid_t dapl) {
   (void)dapl;
   struct dset_rec *r = find_dset(g_files[file - 1000], name);
   if (r == NULL) { fprintf(stderr, "refdrv: no dataset %s in %s\n", name, g_files[file - 1000]); abort(); }
   return 2000 + (hid_t)(r - g_dsets);
}
hid_t H5Dcreate(hid_t file, const char *name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl) {
   (void)lcpl; (void)dcpl; (void)dapl;
   assert(type == H5T_NATIVE_DOUBLE);
   int64_t dims[4];
   int nd = g_spaces[space - 3000].ndims;
   uint64_t n = 1;
   for (int d = 0; d < nd; d++) { dims[d] = (int64_t)g_spaces[space - 3000].dims[d]; n *= dims[d]; }
   void *zeros = calloc(n ? n : 1, 8);
   refdrv_put(g_files[file - 1000], name, DT_F64, nd, dims, zeros);
   free(zeros);
   return 2000 + (hid_t)(find_dset(g_files[file - 1000], name) - g_dsets);
}
hid_t H5Dget_space(hid_t dset) {
   struct dset_rec *r = &g_dsets[dset - 2000];
   if (g_nspaces >= MAX_SPACES) g_nspaces = 0; /* spaces are never closed by the loader: recycle */
   g_spaces[g_nspaces].ndims = r->ndims;
   for (int d = 0; d < r->ndims; d++) g_spaces[g_nspaces].dims[d] = r->dims[d];
   return 3000 + g_nspaces++;
}
herr_t H5Dclose(hid_t dset) { (void)dset; return 0; }

static int mem_type_size(hid_t t) {
   if (t == H5T_NATIVE_DOUBLE || t == H5T_NATIVE_INT64) return 8;
   if (t == H5T_NATIVE_FLOAT) return 4;
   return 1;
}
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t ms, hid_t fs, hid_t xfer, void *buf) {
   (void)ms; (void)fs; (void)xfer;
   struct dset_rec *r = &g_dsets[dset - 2000];
   /* stored and requested types always coincide for the files of SURVEY App. A */
   if (mem_type_size(mem_type) != dt_size(r->dtype)) {
      fprintf(stderr, "refdrv: type mismatch reading %s/%s\n", r->file, r->name);
      return -1;
   }
   if ((mem_type == H5T_NATIVE_DOUBLE) != (r->dtype == DT_F64)) return -1;
   memcpy(buf, r->data, r->nbytes);
   return 0;
}
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t ms, hid_t fs, hid_t xfer, const void *buf) {
   (void)ms; (void)fs; (void)xfer;
   struct dset_rec *r = &g_dsets[dset - 2000];
   if (mem_type_size(mem_type) != dt_size(r->dtype)) return -1;
   memcpy(r->data, buf, r->nbytes);
   return 0;
}
hid_t H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *maxdims) {
   (void)maxdims;
   if (g_nspaces >= MAX_SPACES) g_nspaces = 0;
   g_spaces[g_nspaces].ndims = rank;
   for (int d = 0; d < rank; d++) g_spaces[g_nspaces].dims[d] = dims[d];
   return 3000 + g_nspaces++;
}
herr_t H5Sclose(hid_t space) { (void)space; return 0; }
int H5Sget_simple_extent_ndims(hid_t space) { return g_spaces[space - 3000].ndims; }
int H5Sget_simple_extent_dims(hid_t space, hsize_t *dims, hsize_t *maxdims) {
   (void)maxdims;
   for (int d = 0; d < g_spaces[space - 3000].ndims; d++) dims[d] = g_spaces[space - 3000].dims[d];
   return g_spaces[space - 3000].ndims;
}

/* ------------------------------------------------------------------------------------------
 * driver API (ctypes)
 * ---------------------------------------------------------------------------------------- */
static struct SimData g_sd;
static bool g_loaded = false;
static int g_quiet = 1;

/* the reference prints a 6-line progress block + ioctl per step; silence stdout around calls */
static int quiet_begin(void) {
   if (!g_quiet) return -1;
   fflush(stdout);
   int saved = dup(1);
   int devnull = open("/dev/null", O_WRONLY);
   dup2(devnull, 1);
   close(devnull);
   return saved;
}
static void quiet_end(int saved) {
   if (saved < 0) return;
   fflush(stdout);
   dup2(saved, 1);
   close(saved);
}

REFDRV_API void refdrv_set_quiet(int q) { g_quiet = q; }
REFDRV_API void refdrv_set_threads(int n) { omp_set_num_threads(n); }
REFDRV_API int refdrv_max_threads(void) { return omp_get_max_threads(); }
REFDRV_API int refdrv_precision(void) { return PRECISION; }

/* runs the reference load_sim_data(); `dir` must contain the four (real) .h5 files because the
 * loader stat()s them (fdtd_data.h:143) before "opening" them through the shim */
REFDRV_API int refdrv_load(const char *dir) {
   char cwd[4096];
   if (getcwd(cwd, sizeof cwd) == NULL) return -1;
   if (chdir(dir) != 0) return -2;
   int s = quiet_begin();
   if (g_loaded) { free_sim_data(&g_sd); g_loaded = false; }
   g_nfiles = 0;
   load_sim_data(&g_sd);
   quiet_end(s);
   g_loaded = true;
   if (chdir(cwd) != 0) return -3;
   return 0;
}
REFDRV_API void refdrv_scale_input(void) { int s = quiet_begin(); scale_input(&g_sd); quiet_end(s); }
REFDRV_API double refdrv_run_sim(void) { int s = quiet_begin(); double t = run_sim(&g_sd); quiet_end(s); return t; }
REFDRV_API void refdrv_rescale_output(void) { int s = quiet_begin(); rescale_output(&g_sd); quiet_end(s); }
/* reference write_outputs() -> lands in the registry as ("sim_outs.h5","u_out") */
REFDRV_API void refdrv_write_outputs(void) { int s = quiet_begin(); g_nfiles = 0; write_outputs(&g_sd); quiet_end(s); }
REFDRV_API void refdrv_free(void) { int s = quiet_begin(); if (g_loaded) free_sim_data(&g_sd); g_loaded = false; quiet_end(s); }

REFDRV_API int refdrv_get_dataset(const char *file, const char *name, void *out, uint64_t nbytes) {
   struct dset_rec *r = find_dset(file, name);
   if (r == NULL || r->nbytes != nbytes) return -1;
   memcpy(out, r->data, nbytes);
   return 0;
}

/* field access into the reference's struct SimData (fdtd_data.h:38-76) */
REFDRV_API int refdrv_field(const char *f, const void **ptr, int64_t *count, int *elsize) {
   const struct SimData *sd = &g_sd;
   int64_t Nbm = (sd->Npts - 1) / 8 + 1;
#define FIELD_ARR(nm, cnt) if (strcmp(f, #nm) == 0) { *ptr = sd->nm; *count = (cnt); *elsize = (int)sizeof(*sd->nm); return 0; }
#define FIELD_SCL(nm) if (strcmp(f, #nm) == 0) { *ptr = &sd->nm; *count = 1; *elsize = (int)sizeof(sd->nm); return 0; }
   FIELD_ARR(bn_ixyz, sd->Nb) FIELD_ARR(bnl_ixyz, sd->Nbl) FIELD_ARR(bna_ixyz, sd->Nba) FIELD_ARR(Q_bna, sd->Nba)
   FIELD_ARR(in_ixyz, sd->Ns) FIELD_ARR(out_ixyz, sd->Nr) FIELD_ARR(out_reorder, sd->Nr) FIELD_ARR(adj_bn, sd->Nb)
   FIELD_ARR(ssaf_bnl, sd->Nbl) FIELD_ARR(bn_mask, Nbm) FIELD_ARR(mat_bnl, sd->Nbl) FIELD_ARR(K_bn, sd->Nb)
   FIELD_ARR(in_sigs, sd->Ns * sd->Nt) FIELD_ARR(u_out, sd->Nr * sd->Nt) FIELD_ARR(Mb, sd->Nm)
   FIELD_ARR(mat_beta, sd->Nm)
   if (strcmp(f, "mat_quads") == 0) { *ptr = sd->mat_quads; *count = (int64_t)sd->Nm * MMb * 4; *elsize = (int)sizeof(Real); return 0; }
   FIELD_SCL(Ns) FIELD_SCL(Nr) FIELD_SCL(Nt) FIELD_SCL(Npts) FIELD_SCL(Nx) FIELD_SCL(Ny) FIELD_SCL(Nz)
   FIELD_SCL(Nb) FIELD_SCL(Nbl) FIELD_SCL(Nba) FIELD_SCL(l) FIELD_SCL(l2) FIELD_SCL(fcc_flag) FIELD_SCL(NN)
   FIELD_SCL(Nm) FIELD_SCL(infac) FIELD_SCL(sl2) FIELD_SCL(lo2) FIELD_SCL(a2) FIELD_SCL(a1)
#undef FIELD_ARR
#undef FIELD_SCL
   return -1;
}
