// Incomplete factorisations of the LOCAL matrix block and their triangular sweeps: the
// `preconditioner` keywords ILU, IC and IRILU.
//
// Replaces Preconditioner::init_preconditioner_impl("ILU" | "IC" | "IRILU") + wrap_schwarz
// (Preconditioner/Preconditioner.H:66-82, 106-124, 143-176, 177-196), i.e. Ginkgo's
//   factorization::Ilu / factorization::Ic         exact ILU(0) / IC(0) on the pattern of A
//   preconditioner::Ilu<LowerTrs, UpperTrs>        z = U^-1 (L^-1 r)
//   preconditioner::Ic<LowerTrs>                   z = L^-T (L^-1 r)
//   preconditioner::Ilu<Ir, Ir>                    5 + 5 Jacobi-Richardson sweeps (IRILU)
// generated on distributed::Matrix::get_local_matrix(): no communication in the apply.
//
// The factors live over the CSR pattern of A ("LU in place"): strictly lower part = L (unit
// diagonal implied; IC: L with its diagonal), upper part = U with the diagonal (IC: L^T mirrored),
// so that both sweeps read rows.  Row i of either the factorisation or a sweep needs the finished
// rows named by its own lower (upper) columns: the dependency graph of the pattern.  Once per mesh
// the device computes every row's level (longest dependency chain ending in it), and sorts the rows
// by level (CUB radix sort, stable: rows ascending inside a level).
//   * factorisation: one launch per level, one thread per row, the row's update sequence exactly the
//     sequential IKJ one (bit-identical factors);
//   * sweeps, `tri_variant 1` (default): ONE launch per sweep over the level-ordered rows; a thread
//     polls the entries of x its row needs straight from L2 (the output vector starts as an
//     all-ones NaN pattern, an 8-byte store publishes a finished entry -- the value is its own
//     flag), never blocking: every lane of a warp stays in the same retry loop and stores from
//     inside it, so lanes that wait for lanes of their own warp cannot deadlock, and a row only
//     ever waits for rows at earlier positions of a co-resident grid;
//   * sweeps, `tri_variant 0`: one launch per level (no polling).
// Row sums subtract the products left to right in column order, as the sequential sweep does: both
// variants are bit-identical to each other and to the CPU restatement.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

namespace {

constexpr int kTriThreads = 256;
constexpr long long kTriSpinCycles = 20000000000LL;   // ~10 s: fail loudly instead of hanging

__device__ __forceinline__ double ld_poll(const double *p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_publish(double *p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// position of the diagonal of every row; bad = 1 row without diagonal, 2 a column repeated in a row
// (unsummed cyclic couplings), 3 columns not ascending
__global__ void k_tri_diag(label n, const label *__restrict__ rp, const label *__restrict__ cols,
                           label *__restrict__ dp, int *bad /* [3]: verdict, longest lower part, longest upper part */)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    label d = -1;
    for (label e = rp[i]; e < rp[i + 1]; ++e) {
        const label c = cols[e];
        if (e > rp[i]) {
            if (c == cols[e - 1]) atomicMax(bad, 2);
            else if (c < cols[e - 1]) atomicMax(bad, 3);
        }
        if (c == i && d < 0) d = e;
    }
    dp[i] = d < 0 ? rp[i] : d;
    if (d < 0) {
        atomicMax(bad, 1);
        return;
    }
    atomicMax(bad + 1, d - rp[i]);
    atomicMax(bad + 2, rp[i + 1] - d - 1);
}

// one relaxation pass of  level[i] = 1 + max level[c] over the row's lower (upper) columns, in
// place: levels only grow and never pass the longest-chain value, so the passes end at it
template <bool LOWER>
__global__ void k_tri_levels(label n, const label *__restrict__ rp, const label *__restrict__ cols,
                             const label *__restrict__ dp, label *level, int *changed)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const label lo = LOWER ? rp[i] : dp[i] + 1, hi = LOWER ? dp[i] : rp[i + 1];
    label lv = 1;
    for (label e = lo; e < hi; ++e) {
        const label l = ((volatile label *)level)[cols[e]] + 1;
        lv = l > lv ? l : lv;
    }
    if (lv != level[i]) {
        level[i] = lv;
        *changed = 1;
    }
}

__global__ void k_tri_iota(label n, label *v)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// offs[l] = first position of level l + 1 in the sorted keys (levels start at 1, none is empty)
__global__ void k_tri_level_offsets(label n, const label *__restrict__ keys, label *offs)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || keys[i] != keys[i - 1]) offs[keys[i] - 1] = i;
    if (i == n - 1) offs[keys[i]] = n;
}

// ILU(0), rows perm[lo..hi) of one level.  IKJ: for every lower entry (i,c) ascending
//   l = a_ic / u_cc;  a_ij -= l * u_cj  for the j > c that row i holds.
__global__ void __launch_bounds__(128) k_ilu0_rows(label lo, label hi, const label *__restrict__ perm,
                                                   const label *__restrict__ rp, const label *__restrict__ cols,
                                                   const label *__restrict__ dp, double *F)
{
    const label t = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hi) return;
    const label i = perm[t];
    const label end = rp[i + 1], di = dp[i];
    for (label k = rp[i]; k < di; ++k) {
        const label c = cols[k];
        const double l = F[k] / F[dp[c]];
        F[k] = l;
        label p = k + 1;
        const label cend = rp[c + 1];
        for (label j = dp[c] + 1; j < cend; ++j) {
            const label cj = cols[j];
            while (p < end && cols[p] < cj) ++p;
            if (p < end && cols[p] == cj) F[p] = __dsub_rn(F[p], __dmul_rn(l, F[j]));
        }
    }
}

// IC(0), rows of one level: l_ij = (a_ij - sum_{k<j} l_ik l_jk) / l_jj, l_ii = sqrt(a_ii - sum l_ik^2);
// every l_ij is mirrored into position (j,i).  bad = 4: (j,i) is not in the pattern.
__global__ void __launch_bounds__(128) k_ic0_rows(label lo, label hi, const label *__restrict__ perm,
                                                  const label *__restrict__ rp, const label *__restrict__ cols,
                                                  const label *__restrict__ dp, double *F, int *bad)
{
    const label t = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hi) return;
    const label i = perm[t];
    const label r0 = rp[i], di = dp[i];
    for (label k = r0; k <= di; ++k) {
        const label j = cols[k];
        double s = F[k];
        label a = r0, b = rp[j];
        const label endb = dp[j];
        while (a < k && b < endb) {
            const label ca = cols[a], cb = cols[b];
            if (ca == cb) {
                s = __dsub_rn(s, __dmul_rn(F[a], F[b]));
                ++a;
                ++b;
            } else if (ca < cb) {
                ++a;
            } else {
                ++b;
            }
        }
        F[k] = j < i ? s / F[endb] : sqrt(s);
    }
    for (label k = r0; k < di; ++k) {
        const label j = cols[k];
        bool found = false;
        for (label q = dp[j] + 1; q < rp[j + 1]; ++q)
            if (cols[q] == i) {
                F[q] = F[k];
                found = true;
                break;
            }
        if (!found) atomicMax(bad, 4);
    }
}

struct TriK {
    label n, lo, hi;
    const label *perm, *rp, *cols, *dp;
    const double *F, *b;
    const double *dval;  // F[dp[i]]: the factor's diagonal, contiguous
    double *x;
    const double *xin;   // IR sweeps: the previous iterate
    SolveState *state;
    int guard_done;
    unsigned int sleep_ns;   // dependency-driven sweep: pause between two rounds of polling
};

// dval[i] = F[dp[i]] (after the factorisation): the sweeps read the diagonal without going through dp
__global__ void k_tri_diag_values(label n, const label *__restrict__ dp, const double *__restrict__ F,
                                  double *__restrict__ dval)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dval[i] = F[dp[i]];
}

// x = all-ones bit pattern ("not written yet"), skipped like every kernel of the loop once the solve is done
__global__ void __launch_bounds__(kTriThreads) k_tri_unwritten(const TriK a)
{
    if (a.guard_done && a.state->done) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += stride)
        a.x[i] = __longlong_as_double(-1LL);
}

// one level of a sweep: x_i = (b_i - sum F_ic x_c) / F_ii over the row's lower (upper) columns
template <bool LOWER, bool UNIT>
__global__ void __launch_bounds__(kTriThreads) k_trisolve_rows(const TriK a)
{
    if (a.guard_done && a.state->done) return;
    const label t = a.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.hi) return;
    const label i = a.perm[t];
    const label d = a.dp[i];
    const label lo = LOWER ? a.rp[i] : d + 1, hi = LOWER ? d : a.rp[i + 1];
    double s = a.b[i];
    const double diag = UNIT ? 1.0 : a.dval[i];
    for (label e = lo; e < hi; ++e) s = __dsub_rn(s, __dmul_rn(a.F[e], a.x[a.cols[e]]));
    a.x[i] = UNIT ? s : s / diag;
}

// Dependency-driven sweep for short rows (at most KD entries on the swept side of the diagonal, e.g.
// any hexahedral mesh): the row's columns and factor entries are staged in registers BEFORE the
// first poll, every missing operand is polled in the same round (one L2 round trip per round, not
// one per operand), and the products are subtracted in column order once all have arrived -- the
// sequential sweep's arithmetic.  Between two rounds a warp that still waits pauses for sleep_ns so
// that the warps far ahead of the wavefront do not saturate L2 with polls.
template <bool LOWER, bool UNIT, int KD>
__global__ void __launch_bounds__(kTriThreads) k_trisolve_sf_short(const TriK a)
{
    if (a.guard_done && a.state->done) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5 << 5;
    for (int64_t base = warp * 32; base < a.n; base += stride) {
        const int64_t pos = base + lane;
        bool done = pos >= a.n;
        label row = 0;
        int cnt = 0;
        label c[KD];
        double f[KD], v[KD];
        double s = 0.0, diag = 1.0;
        if (!done) {
            row = a.perm[pos];
            const label d = a.dp[row];
            const label e0 = LOWER ? a.rp[row] : d + 1;
            cnt = (LOWER ? d : a.rp[row + 1]) - e0;
            s = a.b[row];
            if (!UNIT) diag = a.dval[row];
#pragma unroll
            for (int k = 0; k < KD; ++k)
                if (k < cnt) {
                    c[k] = a.cols[e0 + k];
                    f[k] = a.F[e0 + k];
                }
        }
        unsigned int missing = done ? 0u : ((1u << cnt) - 1u);
        long long t0 = 0;
        unsigned int spins = 0;
        while (true) {
            if (!done) {
#pragma unroll
                for (int k = 0; k < KD; ++k)
                    if (missing & (1u << k)) {
                        const double w = ld_poll(a.x + c[k]);
                        if (__double_as_longlong(w) != -1LL) {
                            v[k] = w;
                            missing &= ~(1u << k);
                        }
                    }
                if (missing == 0u) {
#pragma unroll
                    for (int k = 0; k < KD; ++k)
                        if (k < cnt) s = __dsub_rn(s, __dmul_rn(f[k], v[k]));
                    st_publish(a.x + row, UNIT ? s : s / diag);
                    done = true;
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
            if (a.sleep_ns) __nanosleep(a.sleep_ns);
            if ((++spins & 255u) == 0) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kTriSpinCycles) {
                    if (!done) {
                        a.state->comm_error = 2;
                        a.state->done = 1;
                    }
                    break;   // spins is warp-uniform: the whole warp leaves together
                }
            }
        }
    }
}

// a whole sweep in one launch over the level-ordered rows (see the file header).  x must have been
// filled with 0xFF bytes.  Warp w of the grid takes positions 32 w .. 32 w + 31, then strides by the
// grid: a row's dependencies always sit at earlier positions, i.e. in a warp that does not wait for
// this one.
template <bool LOWER, bool UNIT>
__global__ void __launch_bounds__(kTriThreads) k_trisolve_sf(const TriK a)
{
    if (a.guard_done && a.state->done) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5 << 5;
    for (int64_t base = warp * 32; base < a.n; base += stride) {
        const int64_t pos = base + lane;
        bool done = pos >= a.n;
        label row = 0, e = 0, end = 0;
        double s = 0.0, diag = 1.0;
        if (!done) {
            row = a.perm[pos];
            const label d = a.dp[row];
            e = LOWER ? a.rp[row] : d + 1;
            end = LOWER ? d : a.rp[row + 1];
            s = a.b[row];
            if (!UNIT) diag = a.dval[row];
        }
        long long t0 = 0;
        unsigned int spins = 0;
        while (true) {
            if (!done) {
                while (e < end) {
                    const double v = ld_poll(a.x + a.cols[e]);
                    if (__double_as_longlong(v) == -1LL) break;   // not written yet
                    s = __dsub_rn(s, __dmul_rn(a.F[e], v));
                    ++e;
                }
                if (e == end) {
                    st_publish(a.x + row, UNIT ? s : s / diag);
                    done = true;
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
            if (a.sleep_ns) __nanosleep(a.sleep_ns);
            if ((++spins & 1023u) == 0) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kTriSpinCycles) {
                    if (!done) {
                        a.state->comm_error = 2;
                        a.state->done = 1;
                    }
                    break;   // spins is warp-uniform: the whole warp leaves together
                }
            }
        }
    }
}

// one Jacobi-Richardson sweep on a triangular factor T:  out = x + D^-1 (b - T x)
// (gko::solver::Ir with a scalar-Jacobi inner solver, relaxation factor 1)
template <bool LOWER>
__global__ void __launch_bounds__(kTriThreads) k_tri_ir(const TriK a)
{
    if (a.guard_done && a.state->done) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const label d = a.dp[i];
        double t = 0.0;
        if (LOWER) {
            for (label e = a.rp[i]; e < d; ++e) t = __dadd_rn(t, __dmul_rn(a.F[e], a.xin[a.cols[e]]));
            const double xi = a.xin[i];
            t = __dadd_rn(t, __dmul_rn(1.0, xi));
            a.x[i] = __dadd_rn(xi, __dsub_rn(a.b[i], t));
        } else {
            const label hi = a.rp[i + 1];
            for (label e = d; e < hi; ++e) t = __dadd_rn(t, __dmul_rn(a.F[e], a.xin[a.cols[e]]));
            a.x[i] = __dadd_rn(a.xin[i], __dmul_rn(__dsub_rn(a.b[i], t), 1.0 / a.dval[i]));
        }
    }
}

int grid_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// levels of one sweep direction -> perm (rows in level order) + host level offsets
template <bool LOWER>
int analyse_direction(Context *ctx, label *d_level, label *d_keys, label *d_iota, label **perm,
                      std::vector<label> &offsets, int *d_flag)
{
    const label n = ctx->n;
    cudaStream_t st = ctx->stream;
    OGL_CUDA(ctx, cudaMemsetAsync(d_level, 0, sizeof(label) * n, st));
    const int grid = grid_for(n, 256);
    for (int64_t passes = 0;; passes += 16) {
        if (passes > (int64_t)n + 32) return fail(ctx, OGL_ERR_INVALID, "ILU/IC level analysis did not converge");
        OGL_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        for (int p = 0; p < 16; ++p)
            k_tri_levels<LOWER><<<grid, 256, 0, st>>>(n, ctx->d_row_ptrs, ctx->d_cols, ctx->tri.diag_pos, d_level, d_flag);
        ctx->launches += 16;
        int changed = 0;
        OGL_CUDA(ctx, cudaMemcpyAsync(&changed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        OGL_CUDA(ctx, cudaStreamSynchronize(st));
        if (!changed) break;
    }
    OGL_TRY(dev_alloc(ctx, perm, (size_t)n));
    k_tri_iota<<<grid, 256, 0, st>>>(n, d_iota);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_level, d_keys, d_iota, *perm, (int)n, 0, 32, st);
    void *d_tmp = nullptr;
    OGL_CUDA(ctx, cudaMalloc(&d_tmp, bytes + 16));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_tmp, bytes, d_level, d_keys, d_iota, *perm, (int)n, 0, 32, st);
    label n_levels = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n_levels, d_keys + (n - 1), sizeof(label), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_tmp);
    OGL_CUDA(ctx, e);
    if (n_levels < 1 || n_levels > n) return fail(ctx, OGL_ERR_INVALID, "ILU/IC level analysis: bad level count");
    // level offsets, built in the (now free) level array
    k_tri_level_offsets<<<grid, 256, 0, st>>>(n, d_keys, d_level);
    ctx->launches += 2;
    offsets.assign((size_t)n_levels + 1, 0);
    OGL_CUDA(ctx, cudaMemcpyAsync(offsets.data(), d_level, sizeof(label) * ((size_t)n_levels + 1),
                                  cudaMemcpyDeviceToHost, st));
    OGL_CUDA(ctx, cudaStreamSynchronize(st));
    for (label l = 0; l < n_levels; ++l)
        if (offsets[l] >= offsets[l + 1]) return fail(ctx, OGL_ERR_INVALID, "ILU/IC level analysis: empty level");
    return OGL_OK;
}

int analyse(Context *ctx)
{
    Context::TriFactor &t = ctx->tri;
    const label n = ctx->n;
    cudaStream_t st = ctx->stream;
    t.structure_ready = false;
    OGL_TRY(dev_alloc(ctx, &t.diag_pos, (size_t)n));
    int *d_flag = nullptr;
    label *d_level = nullptr, *d_keys = nullptr, *d_iota = nullptr;
    OGL_TRY(dev_alloc(ctx, &d_flag, 3));
    auto cleanup = [&]() {
        cudaFree(d_flag);
        if (d_level) cudaFree(d_level);
        if (d_keys) cudaFree(d_keys);
        if (d_iota) cudaFree(d_iota);
    };
    cudaMemsetAsync(d_flag, 0, 3 * sizeof(int), st);
    k_tri_diag<<<grid_for(n, 256), 256, 0, st>>>(n, ctx->d_row_ptrs, ctx->d_cols, t.diag_pos, d_flag);
    ctx->launches++;
    int verdict[3] = {0, 0, 0};
    cudaMemcpyAsync(verdict, d_flag, 3 * sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    const int bad = verdict[0];
    t.max_lower = verdict[1];
    t.max_upper = verdict[2];
    if (e != cudaSuccess || bad) {
        cleanup();
        if (e != cudaSuccess) return fail(ctx, OGL_ERR_CUDA, std::string("ILU/IC analysis: ") + cudaGetErrorString(e));
        return fail(ctx, bad == 1 ? OGL_ERR_INVALID : OGL_ERR_UNSUPPORTED,
                    bad == 1 ? "ILU/IC: a row without diagonal entry"
                             : "ILU/IC: a row holds the same column twice (unsummed cyclic couplings) or unsorted columns");
    }
    int rc = dev_alloc(ctx, &d_level, (size_t)n + 1);
    if (rc == OGL_OK) rc = dev_alloc(ctx, &d_keys, (size_t)n);
    if (rc == OGL_OK) rc = dev_alloc(ctx, &d_iota, (size_t)n);
    if (rc == OGL_OK) rc = analyse_direction<true>(ctx, d_level, d_keys, d_iota, &t.perm_l, t.lvl_l, d_flag);
    if (rc == OGL_OK) rc = analyse_direction<false>(ctx, d_level, d_keys, d_iota, &t.perm_u, t.lvl_u, d_flag);
    cleanup();
    OGL_TRY(rc);
    // co-resident grid of the dependency-driven sweeps: every CTA must hold an SM slot at once
    int sms = kNumSM, per_sm = 0, lowest = 1 << 30;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
#define TRI_OCC(KERNEL)                                                                  \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KERNEL, kTriThreads, 0);      \
    lowest = per_sm < lowest ? per_sm : lowest
    TRI_OCC((k_trisolve_sf<true, true>));
    TRI_OCC((k_trisolve_sf<true, false>));
    TRI_OCC((k_trisolve_sf<false, false>));
    TRI_OCC((k_trisolve_sf_short<true, true, 8>));
    TRI_OCC((k_trisolve_sf_short<true, false, 8>));
    TRI_OCC((k_trisolve_sf_short<false, false, 8>));
#undef TRI_OCC
    if (lowest < 1) return fail(ctx, OGL_ERR_CUDA, "ILU/IC: the sweep kernel does not fit an SM");
    t.sf_per_sm = lowest;
    t.sf_grid = sms * lowest;
    t.structure_ready = true;
    return OGL_OK;
}

TriK sweep_args(Context *ctx, const double *b, double *x, bool guard)
{
    TriK a{};
    a.n = ctx->n;
    a.rp = ctx->d_row_ptrs;
    a.cols = ctx->d_cols;
    a.dp = ctx->tri.diag_pos;
    a.F = ctx->tri.vals;
    a.dval = ctx->tri.dval;
    a.sleep_ns = (unsigned int)ctx->tri_sleep_ns;
    a.b = b;
    a.x = x;
    a.xin = nullptr;
    a.state = ctx->d_state;
    a.guard_done = guard ? 1 : 0;
    return a;
}

template <bool LOWER, bool UNIT>
int sweep(Context *ctx, const double *b, double *x, bool guard)
{
    const Context::TriFactor &t = ctx->tri;
    cudaStream_t st = ctx->stream;
    TriK a = sweep_args(ctx, b, x, guard);
    a.perm = LOWER ? t.perm_l : t.perm_u;
    if (ctx->tri_variant == 0) {
        const std::vector<label> &lv = LOWER ? t.lvl_l : t.lvl_u;
        for (size_t l = 0; l + 1 < lv.size(); ++l) {
            a.lo = lv[l];
            a.hi = lv[l + 1];
            k_trisolve_rows<LOWER, UNIT><<<grid_for(a.hi - a.lo, kTriThreads), kTriThreads, 0, st>>>(a);
        }
        ctx->launches += (int64_t)lv.size() - 1;
    } else {
        int fill_grid = grid_for(ctx->n, kTriThreads);
        if (fill_grid > kNumSM * 8) fill_grid = kNumSM * 8;
        k_tri_unwritten<<<fill_grid, kTriThreads, 0, st>>>(a);
        // co-resident grid; `tri_ctas` CTAs per SM at most (fewer waiting warps = fewer polls in flight)
        int per_sm = t.sf_per_sm;
        if (ctx->tri_ctas > 0 && ctx->tri_ctas < per_sm) per_sm = (int)ctx->tri_ctas;
        int grid = grid_for(ctx->n, kTriThreads);
        if (grid > t.sf_grid / t.sf_per_sm * per_sm) grid = t.sf_grid / t.sf_per_sm * per_sm;
        const int longest = LOWER ? t.max_lower : t.max_upper;
        if (longest <= 4) k_trisolve_sf_short<LOWER, UNIT, 4><<<grid, kTriThreads, 0, st>>>(a);
        else if (longest <= 8) k_trisolve_sf_short<LOWER, UNIT, 8><<<grid, kTriThreads, 0, st>>>(a);
        else k_trisolve_sf<LOWER, UNIT><<<grid, kTriThreads, 0, st>>>(a);
        ctx->launches += 2;
    }
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

template <bool LOWER>
int ir_sweep(Context *ctx, const double *b, const double *xin, double *xout, bool guard)
{
    TriK a = sweep_args(ctx, b, xout, guard);
    a.xin = xin;
    int grid = grid_for(ctx->n, kTriThreads);
    if (grid > kNumSM * 8) grid = kNumSM * 8;
    k_tri_ir<LOWER><<<grid, kTriThreads, 0, ctx->stream>>>(a);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace

int tri_ensure_structure(Context *ctx)
{
    if (ctx->n == 0) return OGL_OK;
    // the sweeps' intermediates: allocated here, never inside a graph capture
    double *w;
    OGL_TRY(get_work(ctx, 12, &w));
    if (ctx->precond_kind == OGL_PRECOND_IRILU) OGL_TRY(get_work(ctx, 13, &w));
    if (ctx->tri.structure_ready) return OGL_OK;
    if (ctx->tri.vals && ctx->tri.nnz != ctx->nnz)
        return fail(ctx, OGL_ERR_INVALID, "cached ILU/IC factors do not match the regenerated sparsity pattern");
    return analyse(ctx);
}

int tri_setup(Context *ctx, int kind)
{
    if (ctx->n == 0) return OGL_OK;
    Context::TriFactor &t = ctx->tri;
    cudaStream_t st = ctx->stream;
    if (!t.structure_ready) OGL_TRY(analyse(ctx));
    if (!t.vals || t.nnz != ctx->nnz) {
        OGL_TRY(dev_alloc(ctx, &t.vals, (size_t)ctx->nnz));
        t.nnz = ctx->nnz;
    }
    if (!t.dval || t.dval_n != ctx->n) {
        OGL_TRY(dev_alloc(ctx, &t.dval, (size_t)ctx->n));
        t.dval_n = ctx->n;
    }
    OGL_CUDA(ctx, cudaMemcpyAsync(t.vals, ctx->d_vals, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToDevice, st));
    int *d_bad = nullptr;
    const bool ic = kind == OGL_PRECOND_IC;
    if (ic) {
        OGL_TRY(dev_alloc(ctx, &d_bad, 1));
        cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    }
    for (size_t l = 0; l + 1 < t.lvl_l.size(); ++l) {
        const label lo = t.lvl_l[l], hi = t.lvl_l[l + 1];
        const int grid = grid_for(hi - lo, 128);
        if (ic)
            k_ic0_rows<<<grid, 128, 0, st>>>(lo, hi, t.perm_l, ctx->d_row_ptrs, ctx->d_cols, t.diag_pos, t.vals, d_bad);
        else
            k_ilu0_rows<<<grid, 128, 0, st>>>(lo, hi, t.perm_l, ctx->d_row_ptrs, ctx->d_cols, t.diag_pos, t.vals);
    }
    k_tri_diag_values<<<grid_for(ctx->n, 256), 256, 0, st>>>(ctx->n, t.diag_pos, t.vals, t.dval);
    ctx->launches += (int64_t)t.lvl_l.size();
    cudaError_t e = cudaGetLastError();
    int bad = 0;
    if (ic) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_bad);
    }
    OGL_CUDA(ctx, e);
    if (bad) return fail(ctx, OGL_ERR_INVALID, "IC: the sparsity pattern is not structurally symmetric");
    return OGL_OK;
}

int tri_apply(Context *ctx, const double *r, double *z, bool guard)
{
    if (ctx->n == 0) return OGL_OK;
    if (!ctx->tri.structure_ready || !ctx->tri.vals || !ctx->tri.dval)
        return fail(ctx, OGL_ERR_INVALID, "ILU/IC apply before the factorisation");
    double *t;
    OGL_TRY(get_work(ctx, 12, &t));
    if (ctx->precond_kind == OGL_PRECOND_IRILU) {
        // Ilu<Ir, Ir>::apply: the intermediate starts as b, x as the intermediate; 5 sweeps each,
        // iterates ping-pong between two buffers
        double *u;
        OGL_TRY(get_work(ctx, 13, &u));
        OGL_TRY(ir_sweep<true>(ctx, r, r, t, guard));
        OGL_TRY(ir_sweep<true>(ctx, r, t, u, guard));
        OGL_TRY(ir_sweep<true>(ctx, r, u, t, guard));
        OGL_TRY(ir_sweep<true>(ctx, r, t, u, guard));
        OGL_TRY(ir_sweep<true>(ctx, r, u, t, guard));
        OGL_TRY(ir_sweep<false>(ctx, t, t, z, guard));
        OGL_TRY(ir_sweep<false>(ctx, t, z, u, guard));
        OGL_TRY(ir_sweep<false>(ctx, t, u, z, guard));
        OGL_TRY(ir_sweep<false>(ctx, t, z, u, guard));
        return ir_sweep<false>(ctx, t, u, z, guard);
    }
    if (ctx->precond_kind == OGL_PRECOND_ILU) OGL_TRY((sweep<true, true>(ctx, r, t, guard)));
    else OGL_TRY((sweep<true, false>(ctx, r, t, guard)));
    return sweep<false, false>(ctx, t, z, guard);
}

void tri_release(Context *ctx)
{
    Context::TriFactor &t = ctx->tri;
    void *ptrs[] = {t.diag_pos, t.perm_l, t.perm_u, t.vals, t.dval};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    t = Context::TriFactor{};
}

}  // namespace ogl
