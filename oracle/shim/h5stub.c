/* TEST INFRASTRUCTURE (oracle build only) -- never part of the product library.
 *
 * Failing stand-ins for libhdf5 and for the reference's src/util/hdf5_util.c
 * (declared in include/util/hdf5_util.h). HDF5 I/O is outside the DMRG hot
 * path; fixtures are read by oracle/hdf5_lite.py instead.
 */
#include <stdio.h>
#include "hdf5_util.h"

static herr_t fail(const char* what) { fprintf(stderr, "oracle/_ref: HDF5 is stubbed out (%s)\n", what); return -1; }

hid_t  H5Fcreate(const char* f, unsigned fl, hid_t a, hid_t b) { (void)f; (void)fl; (void)a; (void)b; return fail("H5Fcreate"); }
hid_t  H5Fopen(const char* f, unsigned fl, hid_t a) { (void)f; (void)fl; (void)a; return fail("H5Fopen"); }
herr_t H5Fclose(hid_t f) { (void)f; return -1; }
hid_t  H5Dopen2(hid_t f, const char* n, hid_t d) { (void)f; (void)n; (void)d; return fail("H5Dopen2"); }
herr_t H5Dclose(hid_t d) { (void)d; return -1; }
hid_t  H5Dget_type(hid_t d) { (void)d; return -1; }
herr_t H5Tclose(hid_t t) { (void)t; return -1; }

herr_t get_hdf5_dataset_ndims(const hid_t file, const char* name, int* ndims) { (void)file; (void)name; (void)ndims; return fail("get_hdf5_dataset_ndims"); }
herr_t get_hdf5_dataset_dims(const hid_t file, const char* name, hsize_t* dims) { (void)file; (void)name; (void)dims; return fail("get_hdf5_dataset_dims"); }
herr_t read_hdf5_dataset(const hid_t file, const char* name, hid_t mem_type, void* data) { (void)file; (void)name; (void)mem_type; (void)data; return fail("read_hdf5_dataset"); }
herr_t write_hdf5_dataset(const hid_t file, const char* name, int degree, const hsize_t dims[], hid_t s, hid_t i, const void* data) { (void)file; (void)name; (void)degree; (void)dims; (void)s; (void)i; (void)data; return fail("write_hdf5_dataset"); }
herr_t get_hdf5_attribute_dims(const hid_t file, const char* name, hsize_t* dims) { (void)file; (void)name; (void)dims; return fail("get_hdf5_attribute_dims"); }
herr_t read_hdf5_attribute(const hid_t file, const char* name, hid_t mem_type, void* data) { (void)file; (void)name; (void)mem_type; (void)data; return fail("read_hdf5_attribute"); }
herr_t write_hdf5_scalar_attribute(const hid_t file, const char* name, hid_t s, hid_t i, const void* data) { (void)file; (void)name; (void)s; (void)i; (void)data; return fail("write_hdf5_scalar_attribute"); }
herr_t write_hdf5_vector_attribute(const hid_t file, const char* name, hid_t s, hid_t i, const ct_long length, const void* data) { (void)file; (void)name; (void)s; (void)i; (void)length; (void)data; return fail("write_hdf5_vector_attribute"); }
hid_t construct_hdf5_single_complex_dtype(const bool storage) { (void)storage; return -1; }
hid_t construct_hdf5_double_complex_dtype(const bool storage) { (void)storage; return -1; }
enum numeric_type hdf5_to_numeric_dtype(const hid_t dtype) { (void)dtype; return CT_DOUBLE_REAL; }
hid_t numeric_to_hdf5_dtype(const enum numeric_type dtype, const bool storage) { (void)dtype; (void)storage; return -1; }
