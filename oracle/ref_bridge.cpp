// TEST INFRASTRUCTURE -- C bridge onto the REFERENCE's own free functions.
//
// oracle/_ref/libogl_ref.so = this file + the unmodified reference source
// /root/reference/HostMatrix/HostMatrixFreeFunctions.C compiled where it lies
// (see Makefile; only the include of HostMatrix.H is redirected to
// oracle/shim/HostMatrix.H because the real header needs OpenFOAM + Ginkgo).
// Used by tests/ to validate the restatement in assembly.cpp on random meshes.
#include "HostMatrix.H"

extern "C" {

void ref_init_local_sparsity(int nrows, int upper_nnz, int is_symmetric,
                             const int *upper, const int *lower, int *rows,
                             int *cols, int *permute)
{
    Foam::init_local_sparsity(nrows, upper_nnz, is_symmetric != 0, upper, lower,
                              rows, cols, permute);
}

void ref_symmetric_update(int total_nnz, int upper_nnz, const int *permute,
                          double scale, const double *diag, const double *upper,
                          double *out)
{
    Foam::symmetric_update(total_nnz, upper_nnz, permute, scale, diag, upper, out);
}

void ref_non_symmetric_update(int total_nnz, int upper_nnz, const int *permute,
                              double scale, const double *diag,
                              const double *upper, const double *lower,
                              double *out)
{
    Foam::non_symmetric_update(total_nnz, upper_nnz, permute, scale, diag, upper,
                               lower, out);
}

void ref_symmetric_update_w_interface(int total_nnz, int diag_nnz, int upper_nnz,
                                      const int *permute, double scale,
                                      const double *diag, const double *upper,
                                      const double *iface, double *out)
{
    Foam::symmetric_update_w_interface(total_nnz, diag_nnz, upper_nnz, permute,
                                       scale, diag, upper, iface, out);
}

void ref_non_symmetric_update_w_interface(int total_nnz, int diag_nnz,
                                          int upper_nnz, const int *permute,
                                          double scale, const double *diag,
                                          const double *upper,
                                          const double *lower,
                                          const double *iface, double *out)
{
    Foam::non_symmetric_update_w_interface(total_nnz, diag_nnz, upper_nnz,
                                           permute, scale, diag, upper, lower,
                                           iface, out);
}

}  // extern "C"
