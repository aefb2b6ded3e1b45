// Shared between the CSR kernels (spmv.cu) and the ELL-family kernels (ell.cu).
#pragma once

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

struct SpmvK {
    const label *row_ptrs;
    const label *cols;
    const double *vals;
    const double *x;
    const double *y_in;   // advanced: y = alpha*A*x + beta*y_in
    double *y;
    label n;
    label n_row_blocks;   // stream kernel: ceil(n / kRowsPerBlock)
    int blocked;          // 1: each CTA walks a CONTIGUOUS range of tiles (x reuse in L1)
    unsigned long long mat_policy;   // L2 cache policy of the (column, value) stream, see l2_policy()
    double alpha, beta;
    const double *dot_with;
    double *partials;
    unsigned int *ticket;
    SolveState *state;
    int epi, inline_epi, guard_done;
    EpiArgs ea;
};

__device__ __forceinline__ double prod_of(double v, double xv, double alpha, bool adv)
{
    // reference kernels: `alpha * val * b` (advanced) or `val * b`
    return adv ? __dmul_rn(__dmul_rn(alpha, v), xv) : __dmul_rn(v, xv);
}


// (column, value) stream of the pipelined kernel: read-only path, no L1
// allocation, L2 priority from the policy the host picked -- evict-first for a
// matrix much larger than L2, evict-last (for all or an address-hashed fraction
// of the lines) when that share of the matrix can stay L2-resident from one
// Krylov iteration to the next.
__device__ __forceinline__ label ld_mat(const label *p, unsigned long long pol)
{
    label r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ld_mat(const double *p, unsigned long long pol)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}


// ell.cu: variant 7 launchers
int spmv_ell(Context *ctx, SpmvK &k, const SpmvArgs &sa, bool ghosted);

}  // namespace ogl
