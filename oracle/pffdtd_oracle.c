/* TEST INFRASTRUCTURE ONLY.
 *
 * pffdtd_oracle.c -- CPU restatement of the reference simulation step (c_cuda/cpu_engine.h:52-405,
 * SURVEY.md App. B), in fp32 and fp64, taking the same `pffdtd_desc` the product's C ABI takes.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product
 * (pffdtd_b200/) never does.
 *
 * Parity status: PINNED.  oracle/_ref (the unmodified reference engine, built by oracle/Makefile
 * from /root/reference/c_cuda) is executed in tests/test_oracle_vs_ref.py and must agree with this
 * file bit for bit (fp32 and fp64, Cartesian / FCC flag 1 / FCC folded, rigid + lossy materials);
 * tests/golden/ holds receiver traces produced by oracle/_ref for use where /root/reference is absent.
 *
 * Beyond the reference it adds (a) slab support: x_lo_edge/x_hi_edge/ix0 of the desc, with halo planes
 * settable from outside, so the world_size-2 gloo tests can check the slab partition; (b) grid and
 * boundary-state read-back for the energy tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "pffdtd_b200.h"

static void *dupmem(const void *src, int64_t nbytes) {
   void *p = malloc(nbytes > 0 ? (size_t)nbytes : 1);
   if (nbytes > 0) memcpy(p, src, (size_t)nbytes);
   return p;
}

#define REAL float
#define NAME(x) x##_f32
#include "oracle_step.inc"
#undef REAL
#undef NAME
#define REAL double
#define NAME(x) x##_f64
#include "oracle_step.inc"
#undef REAL
#undef NAME

typedef struct oracle {
   int precision;
   ostate_f32 *s32;
   ostate_f64 *s64;
} oracle;

oracle *oracle_create(const pffdtd_desc *d) {
   if (d == NULL || d->struct_size != (int32_t)sizeof(pffdtd_desc)) return NULL;
   oracle *o = (oracle *)calloc(1, sizeof *o);
   o->precision = d->precision;
   if (d->precision == 1) o->s32 = ocreate_f32(d);
   else o->s64 = ocreate_f64(d);
   return o;
}

void oracle_destroy(oracle *o) {
   if (!o) return;
   ofree_f32(o->s32);
   ofree_f64(o->s64);
   free(o);
}

/* steps nstart..nstart+nsteps-1; u_out may be NULL, else [Nr][out_stride] with sample n at column n */
void oracle_run_steps(oracle *o, int64_t nstart, int64_t nsteps, double *u_out, int64_t out_stride) {
   for (int64_t n = nstart; n < nstart + nsteps; n++) {
      if (o->precision == 1) ostep_f32(o->s32, n, u_out, out_stride);
      else ostep_f64(o->s64, n, u_out, out_stride);
   }
}

static int64_t npts(const oracle *o) {
   const pffdtd_desc *d = o->precision == 1 ? &o->s32->d : &o->s64->d;
   return d->Nx * d->Ny * d->Nz;
}

/* which: 1 = u1 (current state), 0 = u0 */
void oracle_read_grid(oracle *o, int which, double *out) {
   const int64_t N = npts(o);
   if (o->precision == 1) { const float *g = which ? o->s32->u1 : o->s32->u0; for (int64_t i = 0; i < N; i++) out[i] = g[i]; }
   else { const double *g = which ? o->s64->u1 : o->s64->u0; memcpy(out, g, (size_t)N * 8); }
}

void oracle_write_grid(oracle *o, int which, const double *in) {
   const int64_t N = npts(o);
   if (o->precision == 1) { float *g = which ? o->s32->u1 : o->s32->u0; for (int64_t i = 0; i < N; i++) g[i] = (float)in[i]; }
   else { double *g = which ? o->s64->u1 : o->s64->u0; memcpy(g, in, (size_t)N * 8); }
}

/* one x-plane (Ny*Nz values) of u1, for the slab halo exchange in the gloo tests */
void oracle_read_plane(oracle *o, int64_t ix, double *out) {
   const pffdtd_desc *d = o->precision == 1 ? &o->s32->d : &o->s64->d;
   const int64_t P = d->Ny * d->Nz;
   if (o->precision == 1) { for (int64_t i = 0; i < P; i++) out[i] = o->s32->u1[ix * P + i]; }
   else memcpy(out, o->s64->u1 + ix * P, (size_t)P * 8);
}

void oracle_write_plane(oracle *o, int64_t ix, const double *in) {
   const pffdtd_desc *d = o->precision == 1 ? &o->s32->d : &o->s64->d;
   const int64_t P = d->Ny * d->Nz;
   if (o->precision == 1) { for (int64_t i = 0; i < P; i++) o->s32->u1[ix * P + i] = (float)in[i]; }
   else memcpy(o->s64->u1 + ix * P, in, (size_t)P * 8);
}

/* vh1, gh1: [Nbl*MMB] in the reference CPU layout nb*MMb+m */
void oracle_read_boundary_state(oracle *o, double *vh1, double *gh1) {
   if (o->precision == 1) {
      const int64_t N = o->s32->d.Nbl * PFFDTD_MMB;
      for (int64_t i = 0; i < N; i++) { vh1[i] = o->s32->vh1[i]; gh1[i] = o->s32->gh1[i]; }
   } else {
      const int64_t N = o->s64->d.Nbl * PFFDTD_MMB;
      memcpy(vh1, o->s64->vh1, (size_t)N * 8);
      memcpy(gh1, o->s64->gh1, (size_t)N * 8);
   }
}
