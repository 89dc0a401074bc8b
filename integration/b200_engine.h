/* b200_engine.h -- the reference-side binding of libpffdtd_b200.so: a third engine next to c_cuda/cpu_engine.h and
 * c_cuda/gpu_engine.h.  Dropped into c_cuda/ and selected in fdtd_main.c:29-33,
 *
 *     #if USING_B200
 *        #include <b200_engine.h>
 *     #elif USING_CUDA
 *        #include <gpu_engine.h>
 *     ...
 *
 * it gives the reference binaries (`make ... -DUSING_B200 -lpffdtd_b200`) the same `double run_sim(const struct SimData *sd)`
 * as the other two engines (gpu_engine.h:665, cpu_engine.h:52): load_sim_data / scale_input before it and rescale_output /
 * write_outputs / print_last_samples after it stay the reference's own code.  Plain C99; must be included after fdtd_data.h
 * (it needs struct SimData, Real, PRECISION, MMb).
 *
 * This file is this repository's code (an adapter over include/pffdtd_b200.h), not a copy of anything in the reference.
 */
#ifndef PFFDTD_B200_ENGINE_H
#define PFFDTD_B200_ENGINE_H

#include <stdio.h>
#include <stdlib.h>
#include <pffdtd_b200.h>

/* Real arrays travel through the C ABI as doubles that hold the Real-rounded values */
static double *b200_widen(const Real *src, int64_t n) {
   double *d = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
   if (!d) {
      fprintf(stderr, "b200 engine: out of memory\n");
      exit(EXIT_FAILURE);
   }
   for (int64_t i = 0; i < n; i++) d[i] = (double)src[i];
   return d;
}

double run_sim(const struct SimData *sd) {
   pffdtd_desc d;
   memset(&d, 0, sizeof d);
   d.struct_size = (int32_t)sizeof d;
   d.precision = PRECISION; /* 1 float, 2 double (fdtd_common.h:43-71) */
   d.fcc_flag = sd->fcc_flag;
   d.Nm = sd->Nm;
   d.Nx = sd->Nx, d.Ny = sd->Ny, d.Nz = sd->Nz;
   d.Nb = sd->Nb, d.Nbl = sd->Nbl, d.Nba = sd->Nba, d.Ns = sd->Ns, d.Nr = sd->Nr, d.Nt = sd->Nt;
   d.l = sd->l, d.l2 = sd->l2;
   d.a1 = (double)sd->a1, d.a2 = (double)sd->a2, d.sl2 = (double)sd->sl2, d.lo2 = (double)sd->lo2;
   d.ix0 = 0, d.x_lo_edge = 1, d.x_hi_edge = 1; /* the whole grid; the library cuts it into one slab per visible device */
   d.bn_ixyz = sd->bn_ixyz, d.adj_bn = sd->adj_bn;
   d.bnl_ixyz = sd->bnl_ixyz, d.mat_bnl = sd->mat_bnl;
   d.bna_ixyz = sd->bna_ixyz, d.Q_bna = sd->Q_bna;
   d.in_ixyz = sd->in_ixyz, d.out_ixyz = sd->out_ixyz;
   d.in_sigs = sd->in_sigs;
   d.Mb = sd->Mb;
   double *ssaf = b200_widen(sd->ssaf_bnl, sd->Nbl);
   double *beta = b200_widen(sd->mat_beta, sd->Nm);
   double *quads = b200_widen((const Real *)sd->mat_quads, (int64_t)sd->Nm * MMb * 4); /* struct MatQuad = 4 Reals */
   d.ssaf_bnl = ssaf, d.mat_beta = beta, d.mat_quads = quads;
   double seconds = 0.0;
   /* u_out comes back in the engine's receiver order, like the other engines'; write_outputs applies out_reorder */
   /* nslabs = 0: every device CUDA_VISIBLE_DEVICES shows, like the reference's own GPU engine (gpu_engine.h:679-691); more than one
    * device needs the sorted "gpu folder" the reference needs too (gpu_engine.h:497-513) */
   if (pffdtd_run_sim_multi(&d, /*nslabs*/ 0, /*devices*/ NULL, sd->u_out, &seconds) != PFFDTD_OK) {
      fprintf(stderr, "b200 engine: %s\n", pffdtd_last_error());
      exit(EXIT_FAILURE); /* the reference's error convention (gpu_engine.h:192-200) */
   }
   free(ssaf), free(beta), free(quads);
   printf("Combined (total): %.6fs, %.2f Mvox/s\n", seconds, sd->Npts * sd->Nt / 1e6 / seconds);
   return seconds;
}

#endif /* PFFDTD_B200_ENGINE_H */
