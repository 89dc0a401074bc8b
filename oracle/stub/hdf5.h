/* TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for <hdf5.h>: just the 3 typedefs, 8 constants and 13
 * functions that the reference's c_cuda/fdtd_data.h uses (fdtd_data.h:145-171,
 * 746-860, 951-977).  libhdf5 does not exist in this image, so the unmodified
 * reference loader/engine is compiled against this header and the in-memory
 * dataset registry implemented in oracle/ref_driver.c.
 */
#ifndef _HDF5_H
#define _HDF5_H
#include <stdint.h>

typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;

#define H5F_ACC_RDONLY 0u
#define H5F_ACC_TRUNC 2u
#define H5P_DEFAULT ((hid_t)0)
#define H5S_ALL ((hid_t)0)

/* type ids (plain constants here; globals in the real library) */
#define H5T_NATIVE_DOUBLE ((hid_t)101)
#define H5T_NATIVE_FLOAT ((hid_t)102)
#define H5T_NATIVE_INT64 ((hid_t)103)
#define H5T_NATIVE_INT8 ((hid_t)104)

hid_t H5Fopen(const char *filename, unsigned flags, hid_t fapl);
hid_t H5Fcreate(const char *filename, unsigned flags, hid_t fcpl, hid_t fapl);
herr_t H5Fclose(hid_t file);
hid_t H5Dopen(hid_t file, const char *name, hid_t dapl);
hid_t H5Dcreate(hid_t file, const char *name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
hid_t H5Dget_space(hid_t dset);
herr_t H5Dclose(hid_t dset);
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t xfer, void *buf);
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t xfer, const void *buf);
hid_t H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *maxdims);
herr_t H5Sclose(hid_t space);
int H5Sget_simple_extent_ndims(hid_t space);
int H5Sget_simple_extent_dims(hid_t space, hsize_t *dims, hsize_t *maxdims);

#endif
