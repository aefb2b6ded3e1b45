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


// TMA (cp.async.bulk) + mbarrier helpers shared by the TMA-fed kernels (spmv.cu:k_spmv_tma,
// ell.cu:k_spmv_ell_tma)
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes,
                                          uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

}  // namespace tma

// ell.cu: variant 7 launchers
int spmv_ell(Context *ctx, SpmvK &k, const SpmvArgs &sa, bool ghosted);
// spmv_merge.cu: variant 8
int spmv_merge(Context *ctx, const SpmvK &k, const SpmvArgs &sa);

}  // namespace ogl
