/* TEST INFRASTRUCTURE (oracle build only) -- never part of the product library.
 *
 * Stub <hdf5.h>: libhdf5 is absent in this image. Only the reference's
 * save_mps/load_mps (src/state/mps.c:1219-1460) touch HDF5 and they are not on
 * the DMRG hot path; the stubs below let that file compile, and every call
 * fails at run time (see h5stub.c).
 */
#ifndef ORACLE_SHIM_HDF5_H
#define ORACLE_SHIM_HDF5_H
#include <stdint.h>
typedef int64_t  hid_t;
typedef int      herr_t;
typedef uint64_t hsize_t;
#define H5F_ACC_RDONLY 0u
#define H5F_ACC_TRUNC  2u
#define H5P_DEFAULT    ((hid_t)0)
#define H5T_NATIVE_INT ((hid_t)1)
#define H5T_STD_I32LE  ((hid_t)2)
hid_t  H5Fcreate(const char* filename, unsigned flags, hid_t fcpl, hid_t fapl);
hid_t  H5Fopen(const char* filename, unsigned flags, hid_t fapl);
herr_t H5Fclose(hid_t file);
hid_t  H5Dopen2(hid_t file, const char* name, hid_t dapl);
#define H5Dopen H5Dopen2
herr_t H5Dclose(hid_t dset);
hid_t  H5Dget_type(hid_t dset);
herr_t H5Tclose(hid_t dtype);
#endif
