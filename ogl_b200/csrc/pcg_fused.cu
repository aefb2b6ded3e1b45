// ogl_b200 -- the whole preconditioned CG loop as ONE persistent cooperative kernel.
//
// A PCG iteration on a 1 M-cell system moves 191 MB and takes ~39 us as three
// kernels per iteration; the device-side timeline (tools/trace_iter.py) shows
// ~11 us of that in kernel boundaries and in the serial tails of the two
// reductions (last CTA -> partial sums -> scalar epilogue -> next launch).
// Here the grid stays resident (5 CTAs per SM, the SpMV's budget) and walks
// the phases of Ginkgo's cg.cpp loop itself:
//
//   P  p' = z + (rho/rho_prev) p        (out of place; ghost entries on several GPUs)
//      -- grid barrier --               (the next SpMV tile's matrix entries already in flight)
//   S  q = A p', <p',q>                 (software-pipelined 256-row tiles, as k_spmv_pipe)
//      -- grid barrier + reduction: last CTA adds the partials in CTA order,
//         all-reduces over the ranks, computes alpha --
//   X  x += alpha p', r' = r - alpha q, z = M^-1 r', <r',z>, |r'|_1
//      -- grid barrier + reduction + OGL's criterion (StoppingCriterion.C:71-151) --
//
// A barrier is one acq_rel ticket per CTA and one released generation word;
// the last CTA to arrive runs the scalar work on a shared-memory copy of the
// SolveState.  The loop runs until the criterion fires: one launch per solve.
// Arithmetic, operation order and the criterion are those of the three-kernel
// path (solver.cu); only the grouping of the partial sums differs (one partial
// per persistent CTA in every phase).
//
// Coherence inside the launch: vectors written by other CTAs (p', q) are read
// with plain ld.global after the barrier's acquire (which invalidates L1);
// never through the non-coherent path.  Matrix and inv-diagonal are read-only.
//
// Several GPUs (peer-memory windows, "ghost p" mode of solver.cu): phase X
// first pushes the new boundary z of the cells into slot 2 of the neighbours'
// windows (r is ping-ponged, so the push reads the old r without a race) as
// self-validating stamped words; phase P updates the ghost entries of p' from them.
#include <cstddef>
#include <cstring>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

namespace {

constexpr int kT = 256;            // threads per CTA
constexpr int kRows = 256;         // rows per SpMV tile
constexpr int kBatchF = 7;         // (column, value) pairs a thread keeps in flight
constexpr int kCtasPerSM = 5;      // <= 51 registers per thread
constexpr int kGenWord = 32;       // barrier: arrivals at [0], generation one 128-byte line further
static_assert(kStateWords <= kT, "one thread per state word");

struct PcgK {
    // CSR (ghosted on several GPUs: columns >= n address the ghost part of p)
    const label *row_ptrs, *cols;
    const double *vals;
    const double *inv_diag;
    label n, n_row_blocks;
    unsigned long long mat_policy;
    // vectors: r and p are ping-ponged (0 = current at launch)
    double *x, *z, *q, *r0, *r1, *p0, *p1;
    SolveState *state;
    double *partials;
    unsigned int *bar;       // [0] arrivals, [kGenWord] generation
    EpiArgs ea;
    int max_iters;
    // ghost-p mode
    int ghost;
    label n_ghost;
    const label *send_idx;
    double *const *push_dst;
};

// ---- memory-model helpers -------------------------------------------------------
__device__ __forceinline__ unsigned int atom_add_acq_rel(unsigned int *p, unsigned int v)
{
    unsigned int r;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
    unsigned int r;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p)
{
    unsigned int r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// coherent loads (L1 allocating; made current by the barrier's acquire)
__device__ __forceinline__ double ld_coh(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_coh2(const double *p)
{
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_l2(const double *p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
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

__device__ __forceinline__ void tr(const PcgK &a, int tag)
{
    if (!a.ea.trace) return;
    EpiArgs e = a.ea;
    e.trace_tag = tag;
    trace_event(e, 0);
}

// scalars every thread needs after a barrier, fetched once per CTA
struct Scalars {
    double coef_p, coef_x, beta;
    int done, p_is_z;
};

struct Shared {
    double sm[2 * 32];
    double state[kStateWords];
    Scalars sc;
    unsigned int gen;
    int is_last;
    int timed_out;
};

// Grid barrier; with NRED > 0 also the deterministic reduction of v over the
// grid, the all-reduce over the ranks and the scalar epilogue, run by the last
// CTA to arrive before it releases the others.  Returns false on a timeout.
template <int NRED>
__device__ __forceinline__ void grid_sync(double (&v)[NRED > 0 ? NRED : 1], const PcgK &a, Shared &sh,
                                          int epi, int ar_count, int tag)
{
    const int tid = threadIdx.x;
    if (NRED > 0) block_sum<(NRED > 0 ? NRED : 1)>(v, sh.sm);
    else __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j) a.partials[(size_t)blockIdx.x * (NRED > 0 ? NRED : 1) + j] = v[j];
        sh.is_last = atom_add_acq_rel(&a.bar[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (sh.is_last) {
        if (NRED > 0 || epi != EPI_NONE) {
            if (tid == 0) tr(a, tag + 1);
            if (tid < kStateWords) sh.state[tid] = ld_l2(reinterpret_cast<const double *>(a.state) + tid);
            double acc[NRED > 0 ? NRED : 1];
#pragma unroll
            for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) acc[j] = 0.0;
            if (NRED > 0) {
#pragma unroll 4
                for (unsigned int b = tid; b < gridDim.x; b += kT) {
#pragma unroll
                    for (int j = 0; j < NRED; ++j) acc[j] += __ldcg(&a.partials[(size_t)b * NRED + j]);
                }
                __syncthreads();   // sh.sm reuse
                block_sum<(NRED > 0 ? NRED : 1)>(acc, sh.sm);
            }
            __syncthreads();       // sh.state complete
            SolveState *s = reinterpret_cast<SolveState *>(sh.state);
            if (tid == 0) {
#pragma unroll
                for (int j = 0; j < NRED; ++j) s->red[j] = acc[j];
                tr(a, tag + 2);
            }
            if (a.ea.comm != nullptr && ar_count > 0) {
                __syncthreads();
                p2p_allreduce(s, ar_count, a.ea.comm);
                if (tid == 0) tr(a, tag + 3);
            }
            if (tid == 0) {
                if (epi != EPI_NONE) run_epilogue(epi, s, a.ea);
                tr(a, tag + 4);
            }
            __syncthreads();
            // write the state back (the comm_error word only when this CTA set it:
            // other CTAs may be raising it in global memory right now)
            if (tid < kStateWords && tid != kCommErrWord)
                reinterpret_cast<double *>(a.state)[tid] = sh.state[tid];
            if (tid == 0 && s->comm_error) a.state->comm_error = 1;
            __syncthreads();
        }
        if (tid == 0) {
            a.bar[0] = 0u;
            st_release(&a.bar[kGenWord], sh.gen + 1u);
        }
    } else if (tid == 0) {
        const long long t0 = clock64();
        sh.timed_out = 0;
        // relaxed polls (an acquire per poll would invalidate this SM's L1 under the
        // CTAs that are still working), one acquire fence once the generation moved
        while (ld_relaxed_u32(&a.bar[kGenWord]) == sh.gen) {
            if (clock64() - t0 > kSpinCycles) {
                sh.timed_out = 1;
                a.state->comm_error = 1;
                break;
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    if (tid == 0) {
        sh.gen += 1u;
        // the scalars of the next phase, once per CTA (L2; the state is hot there)
        const SolveState *g = a.state;
        int done, piz, cerr;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(&g->done) : "memory");
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(piz) : "l"(&g->flag_p_is_z) : "memory");
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(cerr) : "l"(&g->comm_error) : "memory");
        sh.sc.coef_p = ld_l2(&g->coef_p);
        sh.sc.coef_x = ld_l2(&g->coef_x);
        sh.sc.beta = ld_l2(&g->beta);
        sh.sc.done = done | cerr;
        sh.sc.p_is_z = piz;
    }
    __syncthreads();
}

template <int PK>
__global__ void __launch_bounds__(kT, kCtasPerSM) k_pcg_fused(const PcgK a)
{
    extern __shared__ double prod[];
    __shared__ Shared sh;
    const int tid = threadIdx.x;
    const int64_t gid = blockIdx.x * (int64_t)kT + tid;
    const int64_t gstride = (int64_t)gridDim.x * kT;
    const int64_t n2 = a.n >> 1;
    if (tid == 0) {
        sh.gen = ld_acquire(&a.bar[kGenWord]);
        sh.timed_out = 0;
        const SolveState *g = a.state;
        sh.sc.coef_p = ld_l2(&g->coef_p);
        sh.sc.coef_x = ld_l2(&g->coef_x);
        sh.sc.beta = ld_l2(&g->beta);
        sh.sc.done = g->done;
        sh.sc.p_is_z = g->flag_p_is_z;
    }
    __syncthreads();
    double *r_cur = a.r0, *r_nxt = a.r1, *p_old = a.p0, *p_new = a.p1;
    double none[1] = {0.0};
    for (int it = 0; it < a.max_iters; ++it) {
        if (sh.sc.done) break;   // grid-uniform: the state only changes inside barriers
        // ================= P: p' = z + coef_p p ====================================
        {
            if (gid == 0) tr(a, 10);
            const bool p_is_z = sh.sc.p_is_z != 0;
            const double t = sh.sc.coef_p;
            const double *zz = PK == 1 ? a.z : r_cur;
            for (int64_t i = gid; i < n2; i += gstride) {
                const double2 z = ld_coh2(zz + 2 * i);
                double2 p = z;
                if (!p_is_z) {
                    const double2 po = ld_coh2(p_old + 2 * i);
                    p.x = __dadd_rn(z.x, __dmul_rn(t, po.x));
                    p.y = __dadd_rn(z.y, __dmul_rn(t, po.y));
                }
                *reinterpret_cast<double2 *>(p_new + 2 * i) = p;
            }
            if ((a.n & 1) && gid == 0) {
                const int64_t i = a.n - 1;
                const double z = ld_coh(zz + i);
                p_new[i] = p_is_z ? z : __dadd_rn(z, __dmul_rn(t, ld_coh(p_old + i)));
            }
            if (a.ghost) {
                // ghost entries: the neighbours' boundary z (slot 2 of my window), same update
                const CommDev *c = a.ea.comm;
                const unsigned long long *zg =
                    reinterpret_cast<const unsigned long long *>(c->my_recv + 2 * (size_t)c->my_recv_stride);
                const unsigned long long stamp = stamp_of(ld_ar_seq(c));
                const long long t0 = clock64();
                for (int64_t k = gid; k < a.n_ghost; k += gstride) {
                    double z = 0.0;
                    if (!pull_stamped(zg + 2 * k, stamp, t0, z)) a.state->comm_error = 1;
                    p_new[a.n + k] = p_is_z ? z : __dadd_rn(z, __dmul_rn(t, ld_coh(p_old + a.n + k)));
                }
            }
        }
        // ================= S: q = A p', <p',q> =====================================
        double red[2] = {0.0, 0.0};
        {
            label rb = blockIdx.x;
            const label t_last = a.n_row_blocks, t_step = gridDim.x;
            label s = 0, e = 0;
            if (rb < t_last) {
                s = __ldg(&a.row_ptrs[rb * kRows]);
                e = __ldg(&a.row_ptrs[min((rb + 1) * kRows, a.n)]);
            }
            label c[kBatchF];
            double v[kBatchF];
            // the first tile's matrix entries do not depend on p': in flight across the barrier
#pragma unroll
            for (int u = 0; u < kBatchF; ++u) {
                const label q = tid + u * kT;
                c[u] = q < e - s ? ld_mat(&a.cols[s + q], a.mat_policy) : -1;
            }
#pragma unroll
            for (int u = 0; u < kBatchF; ++u) {
                const label q = tid + u * kT;
                v[u] = q < e - s ? ld_mat(&a.vals[s + q], a.mat_policy) : 0.0;
            }
            grid_sync<0>(none, a, sh, EPI_NONE, 0, 10);
            if (sh.timed_out) return;
            if (gid == 0) tr(a, 20);
            for (; rb < t_last; rb += t_step) {
                const label r0 = rb * kRows;
                const label nr = min((label)kRows, a.n - r0);
                const label len = e - s;
                const label rbn = rb + t_step;
                label s2 = 0, e2 = 0;
                if (rbn < t_last) {
                    s2 = __ldg(&a.row_ptrs[rbn * kRows]);
                    e2 = __ldg(&a.row_ptrs[min((rbn + 1) * kRows, a.n)]);
                }
                label rs = 0, re = 0;
                if (tid < nr) {
                    rs = __ldg(&a.row_ptrs[r0 + tid]);
                    re = __ldg(&a.row_ptrs[r0 + tid + 1]);
                }
                {
                    double xv[kBatchF];
#pragma unroll
                    for (int u = 0; u < kBatchF; ++u) xv[u] = c[u] >= 0 ? ld_coh(p_new + c[u]) : 0.0;
#pragma unroll
                    for (int u = 0; u < kBatchF; ++u) {
                        const label q = tid + u * kT;
                        if (q < len) prod[q] = __dmul_rn(v[u], xv[u]);
                    }
                }
                for (label base = kBatchF * kT; base < len; base += kBatchF * kT) {
#pragma unroll
                    for (int u = 0; u < kBatchF; ++u) {
                        const label q = base + tid + u * kT;
                        if (q < len)
                            prod[q] = __dmul_rn(ld_mat(&a.vals[s + q], a.mat_policy),
                                                ld_coh(p_new + ld_mat(&a.cols[s + q], a.mat_policy)));
                    }
                }
#pragma unroll
                for (int u = 0; u < kBatchF; ++u) {
                    const label q = tid + u * kT;
                    c[u] = q < e2 - s2 ? ld_mat(&a.cols[s2 + q], a.mat_policy) : -1;
                }
#pragma unroll
                for (int u = 0; u < kBatchF; ++u) {
                    const label q = tid + u * kT;
                    v[u] = q < e2 - s2 ? ld_mat(&a.vals[s2 + q], a.mat_policy) : 0.0;
                }
                __syncthreads();
                if (tid < nr) {
                    const label row = r0 + tid;
                    double sum = 0.0;
                    for (label q = rs - s; q < re - s; ++q) sum = __dadd_rn(sum, prod[q]);
                    a.q[row] = sum;
                    red[0] = __dadd_rn(red[0], __dmul_rn(ld_coh(p_new + row), sum));
                }
                __syncthreads();
                s = s2;
                e = e2;
            }
        }
        {
            double r1[1] = {red[0]};
            grid_sync<1>(r1, a, sh, EPI_CG_BETA, 1, 20);
            if (sh.timed_out) return;
        }
        // ================= X: x, r', z, <r',z>, |r'|_1 ==============================
        {
            if (gid == 0) tr(a, 30);
            const bool upd = sh.sc.beta != 0.0;
            const double t = sh.sc.coef_x;
            red[0] = red[1] = 0.0;
            if (a.ghost) {
                // new boundary z -> slot 2 of the neighbours' windows (reads the old r: no race)
                const unsigned long long stamp = stamp_of(ld_ar_seq(a.ea.comm) + 1);
                for (int64_t k = gid; k < a.n_ghost; k += gstride) {
                    const label cell = __ldg(&a.send_idx[k]);
                    unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.push_dst[k]);
                    double r = ld_coh(r_cur + cell);
                    if (upd) r = __dsub_rn(r, __dmul_rn(t, ld_coh(a.q + cell)));
                    push_stamped(dst, PK == 1 ? __dmul_rn(r, __ldg(&a.inv_diag[cell])) : r, stamp);
                }
            }
            for (int64_t i = gid; i < n2; i += gstride) {
                double2 r = ld_coh2(r_cur + 2 * i);
                double2 z = make_double2(0.0, 0.0);
                if (upd) {
                    double2 x = ld_coh2(a.x + 2 * i);
                    const double2 p = ld_coh2(p_new + 2 * i);
                    const double2 q = ld_coh2(a.q + 2 * i);
                    x.x = __dadd_rn(x.x, __dmul_rn(t, p.x));
                    x.y = __dadd_rn(x.y, __dmul_rn(t, p.y));
                    r.x = __dsub_rn(r.x, __dmul_rn(t, q.x));
                    r.y = __dsub_rn(r.y, __dmul_rn(t, q.y));
                    *reinterpret_cast<double2 *>(a.x + 2 * i) = x;
                }
                *reinterpret_cast<double2 *>(r_nxt + 2 * i) = r;
                red[1] = __dadd_rn(__dadd_rn(red[1], fabs(r.x)), fabs(r.y));
                if (PK == 1) {
                    const double2 d = __ldg(reinterpret_cast<const double2 *>(a.inv_diag) + i);
                    z.x = __dmul_rn(r.x, d.x);
                    z.y = __dmul_rn(r.y, d.y);
                    red[0] = __dadd_rn(__dadd_rn(red[0], __dmul_rn(r.x, z.x)), __dmul_rn(r.y, z.y));
                    *reinterpret_cast<double2 *>(a.z + 2 * i) = z;
                } else {
                    red[0] = __dadd_rn(__dadd_rn(red[0], __dmul_rn(r.x, r.x)), __dmul_rn(r.y, r.y));
                }
            }
            if ((a.n & 1) && gid == 0) {
                const int64_t i = a.n - 1;
                double r = ld_coh(r_cur + i);
                if (upd) {
                    a.x[i] = __dadd_rn(ld_coh(a.x + i), __dmul_rn(t, ld_coh(p_new + i)));
                    r = __dsub_rn(r, __dmul_rn(t, ld_coh(a.q + i)));
                }
                r_nxt[i] = r;
                red[1] = __dadd_rn(red[1], fabs(r));
                if (PK == 1) {
                    const double z = __dmul_rn(r, __ldg(&a.inv_diag[i]));
                    a.z[i] = z;
                    red[0] = __dadd_rn(red[0], __dmul_rn(r, z));
                } else {
                    red[0] = __dadd_rn(red[0], __dmul_rn(r, r));
                }
            }
            grid_sync<2>(red, a, sh, EPI_CG_RHO_CHECK, 2, 30);
            if (sh.timed_out) return;
        }
        double *tmp = r_cur;
        r_cur = r_nxt;
        r_nxt = tmp;
        tmp = p_old;
        p_old = p_new;
        p_new = tmp;
    }
}

}  // namespace

unsigned long long spmv_l2_policy(Context *ctx);

// the fused loop covers: CG, no or scalar-Jacobi preconditioner, rows short
// enough for the stream tiles; one rank, or several ranks in ghost-p mode
bool pcg_fused_ok(const Context *ctx)
{
    if (!ctx->fused_pcg || ctx->profile_stride > 0) return false;
    // auto: measured faster than three kernels per iteration at 32 k rows (12.9 vs 13.9 us),
    // slower at 1 M (42 vs 38 us: a grid barrier over 740 CTAs costs more than a kernel boundary)
    if (ctx->fused_pcg == 2 && ctx->n > 262144) return false;
    if (ctx->precond_kind != OGL_PRECOND_NONE && (ctx->precond_kind != OGL_PRECOND_BJ || ctx->max_block_size != 1)) return false;
    if (ctx->spmv_variant != 0 && ctx->spmv_variant != 6) return false;
    if (spmv_variant_in_use(ctx) != 6) return false;
    if (ctx->n < 2) return false;
    if (ctx->n_ranks > 1 && !(ctx->ghost_p != 0 && fused_halo_ok(ctx))) return false;
    return true;
}

static int fused_grid(Context *ctx, const void *kernel, size_t smem, int *grid)
{
    int per_sm = 0;
    OGL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    OGL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kT, smem));
    if (per_sm < 1) return fail(ctx, OGL_ERR_UNSUPPORTED, "fused PCG kernel does not fit on an SM");
    if (per_sm > kCtasPerSM) per_sm = kCtasPerSM;
    int sms = kNumSM;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    int64_t g = (int64_t)sms * per_sm;
    const int64_t nblk = ((int64_t)ctx->n + kRows - 1) / kRows;
    if (g > nblk) g = nblk;
    if (ctx->stream_ctas > 0 && g > ctx->stream_ctas) g = ctx->stream_ctas;
    *grid = (int)(g < 1 ? 1 : g);
    return OGL_OK;
}

// r0/z hold the prologue's residual and preconditioned residual; p0/p1/r1/q are scratch
// with room for the ghost entries.  Runs until the device-side criterion fires.
int pcg_fused_run(Context *ctx, double *r0, double *r1, double *z, double *p0, double *p1,
                  double *q, int64_t max_iter)
{
    const bool ghost = ctx->n_ranks > 1;
    const bool ghosted = ghost && ctx->have_ghosted;
    const int pk = ctx->precond_kind == OGL_PRECOND_NONE ? 0 : 1;
    const void *kernel = pk == 1 ? (const void *)k_pcg_fused<1> : (const void *)k_pcg_fused<0>;
    const size_t smem = (size_t)(ghosted ? ctx->max_block_nnz_g : ctx->max_block_nnz) * sizeof(double);
    int grid = 1;
    OGL_TRY(fused_grid(ctx, kernel, smem, &grid));
    if (!ctx->d_bar) {
        OGL_TRY(dev_alloc(ctx, &ctx->d_bar, 2 * kGenWord));
        OGL_CUDA(ctx, cudaMemsetAsync(ctx->d_bar, 0, 2 * kGenWord * sizeof(unsigned int), ctx->stream));
    }
    PcgK a;
    std::memset(&a, 0, sizeof(a));
    a.row_ptrs = ghosted ? ctx->d_g_row_ptrs : ctx->d_row_ptrs;
    a.cols = ghosted ? ctx->d_g_cols : ctx->d_cols;
    a.vals = ghosted ? ctx->d_g_vals : ctx->d_vals;
    a.inv_diag = ctx->d_inv_diag;
    a.n = ctx->n;
    a.n_row_blocks = (label)(((int64_t)ctx->n + kRows - 1) / kRows);
    a.mat_policy = spmv_l2_policy(ctx);
    a.x = ctx->d_x;
    a.z = z;
    a.q = q;
    a.r0 = r0;
    a.r1 = r1;
    a.p0 = p0;
    a.p1 = p1;
    a.state = ctx->d_state;
    a.partials = ctx->d_partials;
    a.bar = ctx->d_bar;
    a.ea = make_epi_args(ctx, 0);
    a.max_iters = (int)(max_iter + 2 < (int64_t)1 << 30 ? max_iter + 2 : (int64_t)1 << 30);
    a.ghost = ghost ? 1 : 0;
    a.n_ghost = ghost ? ctx->n_send : 0;
    a.send_idx = ctx->d_send_idxs;
    a.push_dst = ctx->d_push_dst;
    void *args[] = {&a};
    OGL_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3((unsigned)grid), dim3(kT), args, smem, ctx->stream));
    ctx->launches++;
    return OGL_OK;
}

}  // namespace ogl
