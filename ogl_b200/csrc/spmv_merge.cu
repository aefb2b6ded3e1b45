// Load-balanced ("merge-path") FP64 CSR SpMV for matrices with very uneven row lengths
// (`spmv_variant 8`, picked by the row-length histogram when a few rows are far longer than the mean).
//
// BASELINE.json north_star: "FP64 CSR SpMV (warp-per-row or vectorised merge-path chosen by row-length
// histogram, with 128-bit coalesced loads of values and column indices ...)"; replaces the
// csr::spmv / load-balanced strategies of Ginkgo's Csr for the shapes the row-based
// kernels of spmv.cu handle badly: there a CTA (or warp, or thread) owns ROWS, so one row of 10^6
// entries serialises on one warp while the rest of the GPU idles.
//
// Here a CTA owns a fixed slice of kMpTile = 2048 ENTRIES of the (column, value) stream, whatever
// rows they belong to (the merge-path decomposition with the diagonal search done once per sparsity
// pattern: chunk_row[c] = first row starting at or behind entry c * kMpTile):
//   1. the slice is read with 128-bit coalesced loads (int4 columns -> shared memory, double2 values
//      -> registers), x is gathered, the 2048 products are parked in shared memory;
//   2. rows that START inside the slice are summed from shared memory -- one thread per row, left to
//      right, when the slice holds many rows (bit-identical to the row-ordered kernels for every row
//      that fits in one slice), one warp per row when it holds few -- and stored;
//   3. the leading part of the slice that belongs to a row started in an earlier slice is summed by a
//      warp into carry[c];
//   4. a second launch adds each split row's carries to its y in slice order (deterministic), a third
//      one takes the fused reductions.
// A row split over k slices is summed as (first part) + carry_1 + ... + carry_k: the association
// differs from the sequential row sum, so this variant is tested to 1e-13, like the warp-per-row
// kernel.  Local matrix only (the distributed paths keep the halo-fused CSR kernel).
#include "common.cuh"
#include "reduce.cuh"
#include "spmv.cuh"

namespace ogl {

namespace {

constexpr int kMpTile = 2048;
constexpr int kMpThreads = 256;

// chunk_row[c] = first row r with row_ptrs[r] >= c * kMpTile (c < n_chunks); chunk_row[n_chunks] = n
__global__ void k_mp_chunk_rows(label n, int64_t nnz, const label *__restrict__ rp, label *chunk_row, int n_chunks)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    if (c == n_chunks) {
        chunk_row[c] = n;
        return;
    }
    const int64_t target = (int64_t)c * kMpTile;
    label lo = 0, hi = n;   // rp[n] = nnz >= target: the answer is in [0, n]
    while (lo < hi) {
        const label mid = lo + (hi - lo) / 2;
        if ((int64_t)rp[mid] >= target) hi = mid;
        else lo = mid + 1;
    }
    (void)nnz;
    chunk_row[c] = lo;
}

template <bool ADV>
__global__ void __launch_bounds__(kMpThreads) k_spmv_merge(const SpmvK a, const label *__restrict__ chunk_row,
                                                           int64_t nnz, label *carry_row, double *carry_val)
{
    if (a.guard_done && a.state->done) return;
    __shared__ __align__(16) label cols_s[kMpTile];
    __shared__ __align__(16) double prod_s[kMpTile];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int c = blockIdx.x;
    const int64_t c0 = (int64_t)c * kMpTile;
    const int64_t c1 = c0 + kMpTile < nnz ? c0 + kMpTile : nnz;
    const int len = (int)(c1 - c0);
    if (len == kMpTile) {
        const int4 *c4 = reinterpret_cast<const int4 *>(a.cols + c0);
        const double2 *v2 = reinterpret_cast<const double2 *>(a.vals + c0);
        int4 ci[2];
        double2 vv[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) ci[j] = __ldcs(&c4[t + kMpThreads * j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) vv[j] = __ldcs(&v2[t + kMpThreads * j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) reinterpret_cast<int4 *>(cols_s)[t + kMpThreads * j] = ci[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = 2 * (t + kMpThreads * j);
            const double x0 = __ldg(&a.x[cols_s[idx]]), x1 = __ldg(&a.x[cols_s[idx + 1]]);
            prod_s[idx] = prod_of(vv[j].x, x0, a.alpha, ADV);
            prod_s[idx + 1] = prod_of(vv[j].y, x1, a.alpha, ADV);
        }
    } else {
        for (int i = t; i < len; i += kMpThreads)
            prod_s[i] = prod_of(__ldcs(&a.vals[c0 + i]), __ldg(&a.x[__ldcs(&a.cols[c0 + i])]), a.alpha, ADV);
    }
    __syncthreads();
    const label rfo = chunk_row[c], reo = chunk_row[c + 1];
    const int64_t first_start = a.row_ptrs[rfo];   // rfo <= n, row_ptrs[n] = nnz
    if (reo - rfo > 32) {
        // many rows in the slice: one thread per row, products added left to right
        for (label r = rfo + t; r < reo; r += kMpThreads) {
            const int64_t s = a.row_ptrs[r];
            int64_t e = a.row_ptrs[r + 1];
            e = e < c1 ? e : c1;
            double sum = ADV ? __dmul_rn(a.beta, a.y_in[r]) : 0.0;
            for (int64_t q = s; q < e; ++q) sum = __dadd_rn(sum, prod_s[q - c0]);
            a.y[r] = sum;
        }
    } else {
        // few (long) rows: one warp per row
        for (label r = rfo + warp; r < reo; r += kMpThreads / 32) {
            const int64_t s = a.row_ptrs[r];
            int64_t e = a.row_ptrs[r + 1];
            e = e < c1 ? e : c1;
            double sum = 0.0;
            for (int64_t q = s + lane; q < e; q += 32) sum = __dadd_rn(sum, prod_s[q - c0]);
            sum = warp_sum(sum);
            if (lane == 0) {
                if (ADV) sum = __dadd_rn(__dmul_rn(a.beta, a.y_in[r]), sum);
                a.y[r] = sum;
            }
        }
    }
    // the head of the slice continues a row that started in an earlier slice
    if (warp == kMpThreads / 32 - 1) {
        const bool has_cont = first_start > c0;
        const int64_t cont_end = first_start < c1 ? first_start : c1;
        double sum = 0.0;
        if (has_cont) {
            for (int64_t q = c0 + lane; q < cont_end; q += 32) sum = __dadd_rn(sum, prod_s[q - c0]);
            sum = warp_sum(sum);
        }
        if (lane == 0) {
            carry_row[c] = has_cont ? rfo - 1 : -1;
            carry_val[c] = sum;
        }
    }
}

// y[row] += its carries, in slice order; one thread per run of slices continuing the same row
__global__ void k_spmv_merge_fixup(int n_chunks, const label *__restrict__ carry_row,
                                   const double *__restrict__ carry_val, double *y, const SolveState *state,
                                   int guard_done)
{
    if (guard_done && state->done) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const label r = carry_row[c];
    if (r < 0 || (c > 0 && carry_row[c - 1] == r)) return;
    double acc = y[r];
    for (int cc = c; cc < n_chunks && carry_row[cc] == r; ++cc) acc = __dadd_rn(acc, carry_val[cc]);
    y[r] = acc;
}

// the reductions the iteration needs next: red[0] = <dot_with, y>, red[1] = <y, y>
template <int NRED>
__global__ void __launch_bounds__(256) k_spmv_merge_dot(const SpmvK a)
{
    if (a.guard_done && a.state->done) return;
    double red[NRED];
#pragma unroll
    for (int j = 0; j < NRED; ++j) red[j] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double y = a.y[i];
        red[0] = __dadd_rn(red[0], __dmul_rn(a.dot_with[i], y));
        if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(y, y));
    }
    grid_reduce<NRED>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
}

}  // namespace

// once per sparsity pattern (spmv_setup, never inside a graph capture)
int spmv_merge_setup(Context *ctx)
{
    const int64_t n_chunks = (ctx->nnz + kMpTile - 1) / kMpTile;
    ctx->mp_chunks = (int)n_chunks;
    if (n_chunks == 0) return OGL_OK;
    OGL_TRY(dev_alloc(ctx, &ctx->d_mp_chunk_row, (size_t)n_chunks + 1));
    OGL_TRY(dev_alloc(ctx, &ctx->d_mp_carry_row, (size_t)n_chunks));
    OGL_TRY(dev_alloc(ctx, &ctx->d_mp_carry_val, (size_t)n_chunks));
    k_mp_chunk_rows<<<(int)((n_chunks + 1 + 255) / 256), 256, 0, ctx->stream>>>(ctx->n, ctx->nnz, ctx->d_row_ptrs,
                                                                                 ctx->d_mp_chunk_row, (int)n_chunks);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

int spmv_merge(Context *ctx, const SpmvK &k, const SpmvArgs &sa)
{
    const int n_chunks = ctx->mp_chunks;
    if (n_chunks == 0 || !ctx->d_mp_chunk_row)
        return fail(ctx, OGL_ERR_INVALID, "merge-path SpMV without its slice table");
    cudaStream_t st = ctx->stream;
    if (sa.advanced)
        k_spmv_merge<true><<<n_chunks, kMpThreads, 0, st>>>(k, ctx->d_mp_chunk_row, ctx->nnz, ctx->d_mp_carry_row,
                                                            ctx->d_mp_carry_val);
    else
        k_spmv_merge<false><<<n_chunks, kMpThreads, 0, st>>>(k, ctx->d_mp_chunk_row, ctx->nnz, ctx->d_mp_carry_row,
                                                             ctx->d_mp_carry_val);
    k_spmv_merge_fixup<<<(n_chunks + 255) / 256, 256, 0, st>>>(n_chunks, ctx->d_mp_carry_row, ctx->d_mp_carry_val, k.y,
                                                               k.state, k.guard_done);
    ctx->launches += 2;
    if (sa.nred > 0) {
        int64_t grid = ((int64_t)ctx->n + 255) / 256;
        if (grid > ctx->blas1_blocks) grid = ctx->blas1_blocks;
        if (grid < 1) grid = 1;
        if (sa.nred == 1) k_spmv_merge_dot<1><<<(int)grid, 256, 0, st>>>(k);
        else k_spmv_merge_dot<2><<<(int)grid, 256, 0, st>>>(k);
        ctx->launches++;
    }
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace ogl
