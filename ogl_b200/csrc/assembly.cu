// Device-side LDU -> row-major COO/CSR assembly.
//
// Replaces, bit for bit on the integer structures:
//   init_local_sparsity            HostMatrix/HostMatrixFreeFunctions.C:105-201
//   cyclic interface merge          HostMatrix/HostMatrix.C:504-586
//   init_non_local_sparsity_pattern HostMatrix/HostMatrix.C:438-466
//   update_local_matrix_data        HostMatrix/HostMatrix.C:592-705
//   update_non_local_matrix_data    HostMatrix/HostMatrix.C:708-732
//
// The reference sorts two arrays of (row, col, face) tuples with std::sort on
// the host and merges them row by row.  Here the same ordering is produced by a
// counting sort on the row (histogram -> scan -> scatter) followed by an
// in-row sort on (col, staging slot): inside a row the lower entries have
// col < row, the diagonal col == row, the upper entries col > row, and a cyclic
// interface entry that collides with an existing (row, col) has the larger
// staging slot, so it lands behind it exactly like HostMatrix.C:543-575.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ogl {

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work, int threads = kThreads)
{
    int64_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

// row histogram; counts start at 1 (the diagonal).  Also validates the
// addressing: 0 <= lower < upper < n (OpenFOAM owner < neighbour).
__global__ void k_count_rows(label n, label nf, const label *__restrict__ lower,
                             const label *__restrict__ upper, label n_if,
                             const label *__restrict__ if_rows,
                             const label *__restrict__ if_cols, label *counts,
                             int *bad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nf) {
        const label l = lower[i], u = upper[i];
        if (l < 0 || u >= n || l >= u) {
            atomicExch(bad, 1);
            return;
        }
        atomicAdd(&counts[l], 1);
        atomicAdd(&counts[u], 1);
    } else if (i < (int64_t)nf + n_if) {
        const label k = (label)(i - nf);
        const label r = if_rows[k], c = if_cols[k];
        if (r < 0 || r >= n || c < 0 || c >= n) {
            atomicExch(bad, 2);
            return;
        }
        atomicAdd(&counts[r], 1);
    }
}

// longest row (its entry count is known before the scatter).  The per-row insertion sort below is
// quadratic in the row length: rows beyond kLongRow entries (a cell coupled to thousands of others) are
// left to a radix sort of their own (sort_long_rows)
constexpr label kLongRow = 4096;
constexpr int kMaxLongRows = 4096;
__global__ void k_max_count(label n, const label *__restrict__ counts, label *mx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int v = i < n ? counts[i] : 0;
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(mx, (label)v);
}

__global__ void k_fill_ones(label n, label *a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = 1;
}

// scatter every entry into its row segment (order inside the row is fixed by
// k_sort_rows afterwards, so the atomics do not leak nondeterminism)
__global__ void k_scatter(label n, label nf, int symmetric,
                          const label *__restrict__ lower,
                          const label *__restrict__ upper, label n_if,
                          const label *__restrict__ if_rows,
                          const label *__restrict__ if_cols,
                          const label *__restrict__ row_ptrs, label *cursor,
                          label *cols, label *map)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const label diag_base = symmetric ? nf : 2 * nf;
    if (i < nf) {
        const label f = (label)i;
        const label l = lower[f], u = upper[f];
        // upper-triangle entry: row = lowerAddr, col = upperAddr, slot f
        label pos = row_ptrs[l] + atomicAdd(&cursor[l], 1);
        cols[pos] = u;
        map[pos] = f;
        // lower-triangle entry: row = upperAddr, col = lowerAddr,
        // slot f (symmetric: shares the upper coefficient) or F + f
        pos = row_ptrs[u] + atomicAdd(&cursor[u], 1);
        cols[pos] = l;
        map[pos] = symmetric ? f : nf + f;
    } else if (i < (int64_t)nf + n) {
        const label r = (label)(i - nf);
        const label pos = row_ptrs[r] + atomicAdd(&cursor[r], 1);
        cols[pos] = r;
        map[pos] = diag_base + r;
    } else if (i < (int64_t)nf + n + n_if) {
        const label k = (label)(i - nf - n);
        const label r = if_rows[k];
        const label pos = row_ptrs[r] + atomicAdd(&cursor[r], 1);
        cols[pos] = if_cols[k];
        map[pos] = diag_base + n + k;   // HostMatrix.C:574
    }
}

// one thread per row: insertion sort on (col, slot), write the COO row index,
// track the longest row
__global__ void k_sort_rows(label n, const label *__restrict__ row_ptrs, label *rows,
                            label *cols, label *map, label *max_len)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const label s = row_ptrs[r], e = row_ptrs[r + 1];
    for (label i = s + 1; i < e; ++i) {
        const label c = cols[i], m = map[i];
        label j = i - 1;
        while (j >= s && (cols[j] > c || (cols[j] == c && map[j] > m))) {
            cols[j + 1] = cols[j];
            map[j + 1] = map[j];
            --j;
        }
        cols[j + 1] = c;
        map[j + 1] = m;
    }
    for (label i = s; i < e; ++i) rows[i] = (label)r;
    atomicMax(max_len, e - s);
}

// k_sort_rows for a pattern with long rows: those are only recorded as (row, first, last) triples
__global__ void k_sort_rows_capped(label n, const label *__restrict__ row_ptrs, label *rows, label *cols,
                                   label *map, label *max_len, label *long_rows, int *n_long)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const label s = row_ptrs[r], e = row_ptrs[r + 1];
    atomicMax(max_len, e - s);
    if (e - s > kLongRow) {
        const int k = atomicAdd(n_long, 1);
        if (k < kMaxLongRows) {
            long_rows[3 * k] = (label)r;
            long_rows[3 * k + 1] = s;
            long_rows[3 * k + 2] = e;
        }
        return;
    }
    for (label i = s + 1; i < e; ++i) {
        const label c = cols[i], m = map[i];
        label j = i - 1;
        while (j >= s && (cols[j] > c || (cols[j] == c && map[j] > m))) {
            cols[j + 1] = cols[j];
            map[j + 1] = map[j];
            --j;
        }
        cols[j + 1] = c;
        map[j + 1] = m;
    }
    for (label i = s; i < e; ++i) rows[i] = (label)r;
}

// one long row: (col, slot) -> 64-bit keys whose ascending order is the insertion sort's order
__global__ void k_pack_row_keys(label len, const label *__restrict__ cols, const label *__restrict__ map,
                                unsigned long long *keys)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < len) keys[i] = ((unsigned long long)(unsigned int)cols[i] << 32) | (unsigned int)map[i];
}

__global__ void k_unpack_row_keys(label len, const unsigned long long *__restrict__ keys, label *cols, label *map,
                                  label *rows, label row)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= len) return;
    cols[i] = (label)(keys[i] >> 32);
    map[i] = (label)(keys[i] & 0xffffffffull);
    rows[i] = row;
}

__global__ void k_iota(label n, label *a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (label)i;
}

__global__ void k_validate_rows(label n_rows, label n, const label *__restrict__ rows,
                                int *bad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && (rows[i] < 0 || rows[i] >= n_rows)) atomicExch(bad, 3);
}

// heads of the runs of equal rows in the sorted non-local pattern
__global__ void k_mark_heads(label n, const label *__restrict__ rows, label *head)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || rows[i] != rows[i - 1]) ? 1 : 0;
}

__global__ void k_compact_heads(label n, const label *__restrict__ rows,
                                const label *__restrict__ head,
                                const label *__restrict__ head_scan, label *row_ids,
                                label *row_ptrs, label n_unique)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && head[i]) {
        row_ids[head_scan[i]] = rows[i];
        row_ptrs[head_scan[i]] = (label)i;
    }
    if (i == 0) row_ptrs[n_unique] = n;
}

// coefficient gather (HostMatrix.C:685-703 row_gather + CsrMatrixWrapper.H:123-135
// value copy, fused): vals[k] = scaling * staging[map[k]], the local interface
// segment negated (HostMatrix.C:204).
__global__ void k_gather_values(int64_t nnz, const label *__restrict__ map,
                                const double *__restrict__ staging, label iface_base,
                                double scaling, double *__restrict__ vals)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const label m = __ldcs(&map[k]);
        double v = __ldg(&staging[m]);
        if (m >= iface_base) v = v * -1.0;
        __stcs(&vals[k], scaling == 1.0 ? v : scaling * v);
    }
}

// ghosted CSR: structure
__global__ void k_ghost_counts(label n_groups, const label *__restrict__ row_ids,
                               const label *__restrict__ nl_row_ptrs, label *__restrict__ cnt)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < n_groups) cnt[row_ids[g]] = nl_row_ptrs[g + 1] - nl_row_ptrs[g];
}

__global__ void k_ghost_row_ptrs(label n, const label *__restrict__ row_ptrs,
                                 const label *__restrict__ before, label *__restrict__ g_row_ptrs)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r <= n) g_row_ptrs[r] = row_ptrs[r] + before[r];
}

__global__ void k_ghost_local(int64_t nnz, const label *__restrict__ rows,
                              const label *__restrict__ cols, const label *__restrict__ map,
                              const label *__restrict__ before, label *__restrict__ g_cols,
                              label *__restrict__ g_map)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const int64_t dst = k + before[rows[k]];
        g_cols[dst] = cols[k];
        g_map[dst] = map[k];
    }
}

__global__ void k_ghost_halo(label n, label n_groups, const label *__restrict__ row_ids,
                             const label *__restrict__ nl_row_ptrs,
                             const label *__restrict__ nl_cols, const label *__restrict__ nl_map,
                             const label *__restrict__ g_row_ptrs, label *__restrict__ g_cols,
                             label *__restrict__ g_map)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const label row = row_ids[g];
    const label q0 = nl_row_ptrs[g], q1 = nl_row_ptrs[g + 1];
    label dst = g_row_ptrs[row + 1] - (q1 - q0);   // behind the row's local entries
    for (label q = q0; q < q1; ++q, ++dst) {
        g_cols[dst] = n + nl_cols[q];   // ghost column: index into the receive window
        g_map[dst] = ~nl_map[q];        // negative: index into the non-local staging
    }
}

__global__ void k_row_len_max(label n, const label *__restrict__ row_ptrs, int *out)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) atomicMax(out, row_ptrs[r + 1] - row_ptrs[r]);
}

__global__ void k_block_span_max(label n, const label *__restrict__ row_ptrs, int rows_per_block,
                                 int *out)
{
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r0 = b * rows_per_block;
    if (r0 < n) {
        const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
        atomicMax(out, row_ptrs[r1] - row_ptrs[r0]);
    }
}

// ghosted CSR: coefficients, same arithmetic as k_gather_values / k_gather_nonlocal
__global__ void k_gather_ghosted(int64_t nnz, const label *__restrict__ map,
                                 const double *__restrict__ staging, label iface_base,
                                 const double *__restrict__ nl_staging, double scaling,
                                 double *__restrict__ vals)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const label m = __ldcs(&map[k]);
        double v;
        if (m >= 0) {
            v = __ldg(&staging[m]);
            if (m >= iface_base) v = v * -1.0;
        } else {
            v = __ldg(&nl_staging[~m]) * -1.0;
        }
        __stcs(&vals[k], scaling == 1.0 ? v : scaling * v);
    }
}

__global__ void k_gather_nonlocal(label n_halo, const label *__restrict__ map,
                                  const double *__restrict__ bou, double scaling,
                                  double *__restrict__ vals)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_halo) {
        const double v = bou[map[k]] * -1.0;   // HostMatrix.C:204 then :723-726
        vals[k] = scaling == 1.0 ? v : scaling * v;
    }
}

}  // namespace

// Rows of a pattern that has rows longer than kLongRow: the short ones by the per-thread insertion sort,
// every long one by a CUB radix sort of its (col, slot) keys -- the same ascending (col, slot) order.
static int sort_long_rows(Context *ctx, label n, label longest, label *d_max)
{
    cudaStream_t st = ctx->stream;
    label *d_long = nullptr;
    int *d_nlong = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys_s = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_long), cudaFree(d_nlong), cudaFree(d_keys), cudaFree(d_keys_s), cudaFree(d_tmp);
    };
    int rc = dev_alloc(ctx, &d_long, (size_t)3 * kMaxLongRows);
    if (rc == OGL_OK) rc = dev_alloc(ctx, &d_nlong, 1);
    if (rc == OGL_OK) rc = dev_alloc(ctx, &d_keys, (size_t)longest);
    if (rc == OGL_OK) rc = dev_alloc(ctx, &d_keys_s, (size_t)longest);
    if (rc != OGL_OK) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_nlong, 0, sizeof(int), st);
    k_sort_rows_capped<<<grid_for(n), kThreads, 0, st>>>(n, ctx->d_row_ptrs, ctx->d_rows, ctx->d_cols, ctx->d_map,
                                                          d_max, d_long, d_nlong);
    int n_long = 0;
    cudaMemcpyAsync(&n_long, d_nlong, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && n_long > kMaxLongRows) {
        cleanup();
        return fail(ctx, OGL_ERR_UNSUPPORTED, "more than " + std::to_string(kMaxLongRows) + " rows longer than " +
                                                  std::to_string(kLongRow) + " entries");
    }
    std::vector<label> long_rows((size_t)3 * (n_long > 0 ? n_long : 1));
    if (e == cudaSuccess && n_long > 0)
        e = cudaMemcpy(long_rows.data(), d_long, sizeof(label) * 3 * (size_t)n_long, cudaMemcpyDeviceToHost);
    size_t bytes = 0;
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_keys, d_keys_s, (int)longest, 0, 64, st);
    if (e == cudaSuccess) e = cudaMalloc(&d_tmp, bytes + 16);
    for (int k = 0; e == cudaSuccess && k < n_long; ++k) {
        const label r = long_rows[3 * k], s = long_rows[3 * k + 1], len = long_rows[3 * k + 2] - s;
        k_pack_row_keys<<<grid_for(len), kThreads, 0, st>>>(len, ctx->d_cols + s, ctx->d_map + s, d_keys);
        e = cub::DeviceRadixSort::SortKeys(d_tmp, bytes, d_keys, d_keys_s, (int)len, 0, 64, st);
        k_unpack_row_keys<<<grid_for(len), kThreads, 0, st>>>(len, d_keys_s, ctx->d_cols + s, ctx->d_map + s,
                                                              ctx->d_rows + s, r);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(ctx, OGL_ERR_CUDA, std::string("sort_long_rows: ") + cudaGetErrorString(e));
    return OGL_OK;
}

int pattern_from_ldu(Context *ctx, label n, label nf, bool sym, const label *lower,
                     const label *upper, label n_if, const label *if_rows,
                     const label *if_cols)
{
    if (n < 0 || nf < 0 || n_if < 0) return fail(ctx, OGL_ERR_INVALID, "negative size");
    if ((nf > 0 && (!lower || !upper)) || (n_if > 0 && (!if_rows || !if_cols)))
        return fail(ctx, OGL_ERR_INVALID, "null addressing pointer");
    const int64_t nnz = (int64_t)n + 2 * (int64_t)nf + n_if;
    if (nnz > INT32_MAX)
        return fail(ctx, OGL_ERR_UNSUPPORTED, "local nnz exceeds label (int32) range");
    cudaStream_t st = ctx->stream;

    label *d_lower = nullptr, *d_upper = nullptr, *d_ifr = nullptr, *d_ifc = nullptr;
    label *d_cursor = nullptr, *d_counts = nullptr, *d_max = nullptr;
    int *d_bad = nullptr;
    void *d_tmp = nullptr;
    label longest_row = 0;
    auto cleanup = [&]() {
        cudaFree(d_lower), cudaFree(d_upper), cudaFree(d_ifr), cudaFree(d_ifc);
        cudaFree(d_cursor), cudaFree(d_counts), cudaFree(d_max), cudaFree(d_bad);
        cudaFree(d_tmp);
    };
#define TRY_CLEAN(expr)                      \
    do {                                     \
        int rc__ = (expr);                   \
        if (rc__ != OGL_OK) {                \
            cleanup();                       \
            return rc__;                     \
        }                                    \
    } while (0)

    TRY_CLEAN(dev_alloc(ctx, &d_lower, nf));
    TRY_CLEAN(dev_alloc(ctx, &d_upper, nf));
    TRY_CLEAN(dev_alloc(ctx, &d_ifr, n_if));
    TRY_CLEAN(dev_alloc(ctx, &d_ifc, n_if));
    TRY_CLEAN(dev_alloc(ctx, &d_counts, (size_t)n + 1));
    TRY_CLEAN(dev_alloc(ctx, &d_cursor, n));
    TRY_CLEAN(dev_alloc(ctx, &d_max, 1));
    TRY_CLEAN(dev_alloc(ctx, &d_bad, 1));
    TRY_CLEAN(upload(ctx, d_lower, lower, sizeof(label) * nf));
    TRY_CLEAN(upload(ctx, d_upper, upper, sizeof(label) * nf));
    TRY_CLEAN(upload(ctx, d_ifr, if_rows, sizeof(label) * n_if));
    TRY_CLEAN(upload(ctx, d_ifc, if_cols, sizeof(label) * n_if));

    TRY_CLEAN(dev_alloc(ctx, &ctx->d_rows, nnz));
    TRY_CLEAN(dev_alloc(ctx, &ctx->d_cols, nnz));
    TRY_CLEAN(dev_alloc(ctx, &ctx->d_map, nnz));
    TRY_CLEAN(dev_alloc(ctx, &ctx->d_row_ptrs, (size_t)n + 1));

    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    cudaMemsetAsync(d_max, 0, sizeof(label), st);
    cudaMemsetAsync(d_cursor, 0, sizeof(label) * (size_t)(n > 0 ? n : 1), st);
    cudaMemsetAsync(d_counts, 0, sizeof(label) * ((size_t)n + 1), st);
    k_fill_ones<<<grid_for(n), kThreads, 0, st>>>(n, d_counts);
    k_count_rows<<<grid_for((int64_t)nf + n_if), kThreads, 0, st>>>(
        n, nf, d_lower, d_upper, n_if, d_ifr, d_ifc, d_counts, d_bad);
    k_max_count<<<grid_for(n), kThreads, 0, st>>>(n, d_counts, d_max);
    {
        // stop before the scatter if the addressing is malformed (the counts
        // would not cover the entries the scatter writes)
        int bad0 = 0;
        label longest = 0;
        cudaMemcpyAsync(&bad0, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&longest, d_max, sizeof(label), cudaMemcpyDeviceToHost, st);
        cudaError_t e0 = cudaStreamSynchronize(st);
        longest_row = longest;
        if (e0 != cudaSuccess || bad0 != 0) {
            cleanup();
            if (e0 != cudaSuccess)
                return fail(ctx, OGL_ERR_CUDA,
                            std::string("pattern_from_ldu: ") + cudaGetErrorString(e0));
            return fail(ctx, OGL_ERR_INVALID,
                        bad0 == 1 ? "lduAddressing violates 0 <= lowerAddr < upperAddr < nRows"
                                  : "local interface index out of range");
        }
    }
    // exclusive scan of the n counts (+ trailing 0) -> row_ptrs[0..n]
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, ctx->d_row_ptrs, n + 1, st);
    if (cudaMalloc(&d_tmp, tmp_bytes + 16) != cudaSuccess) {
        cleanup();
        return fail(ctx, OGL_ERR_CUDA, "cudaMalloc(scan temp)");
    }
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, ctx->d_row_ptrs, n + 1, st);
    k_scatter<<<grid_for((int64_t)nf + n + n_if), kThreads, 0, st>>>(
        n, nf, sym ? 1 : 0, d_lower, d_upper, n_if, d_ifr, d_ifc, ctx->d_row_ptrs, d_cursor,
        ctx->d_cols, ctx->d_map);
    if (longest_row <= kLongRow) {
        k_sort_rows<<<grid_for(n), kThreads, 0, st>>>(n, ctx->d_row_ptrs, ctx->d_rows,
                                                       ctx->d_cols, ctx->d_map, d_max);
    } else {
        const int rc_long = sort_long_rows(ctx, n, longest_row, d_max);
        if (rc_long != OGL_OK) {
            cleanup();
            return rc_long;
        }
    }
    int bad = 0;
    label max_len = 0;
    cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&max_len, d_max, sizeof(label), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
#undef TRY_CLEAN
    if (e != cudaSuccess)
        return fail(ctx, OGL_ERR_CUDA, std::string("pattern_from_ldu: ") + cudaGetErrorString(e));
    (void)bad;

    const label n_before = ctx->n;
    ctx->have_pattern_before = ctx->have_pattern;
    ctx->n = n;
    ctx->n_faces = nf;
    ctx->n_local_iface = n_if;
    ctx->symmetric = sym;
    ctx->nnz = nnz;
    ctx->max_row_len = max_len;
    ctx->have_pattern = true;
    ell_invalidate(ctx, /*structure=*/true);
    ctx->tri.structure_ready = false;   // ILU / IC dependency levels belong to the old pattern (trifactor.cu)
    ctx->have_values = false;
    // `regenerate true` rebuilds the pattern of the same mesh every solve; the reference's
    // PersistentVector b / x survive that (lduLduBase.H:217-237), so do these
    const bool keep_vectors = n_before == n && ctx->have_pattern_before && ctx->d_b && ctx->d_x;
    if (!keep_vectors) ctx->have_b = ctx->have_x = ctx->have_precond = false;
    // a new local pattern invalidates the halo description built on the old one
    ctx->have_nonlocal = false;
    ctx->have_ghosted = false;
    ctx->have_partition = false;
    ctx->n_halo = 0;
    ctx->n_nl_rows = 0;
    ctx->n_targets = 0;
    ctx->n_send = 0;
    ctx->global_n = n;
    if (!keep_vectors) {
        ctx->n_blocks = 0;
        ctx->bj_pattern_mbs = 0;
        if (ctx->d_inv_diag) {
            cudaFree(ctx->d_inv_diag);   // sized for the previous pattern
            ctx->d_inv_diag = nullptr;
        }
    }
    if (ctx->graph_exec) {
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
    }
    // value buffers sized for this pattern
    ctx->staging_len = (size_t)nf * (sym ? 1 : 2) + n + n_if;
    OGL_TRY(dev_alloc(ctx, &ctx->d_staging, ctx->staging_len));
    OGL_TRY(dev_alloc(ctx, &ctx->d_vals, nnz));
    if (!keep_vectors) {
        OGL_TRY(dev_alloc(ctx, &ctx->d_b, n));
        OGL_TRY(dev_alloc(ctx, &ctx->d_x, n));
    }
    for (auto &w : ctx->work) {
        if (w) cudaFree(w);
        w = nullptr;
    }
    ctx->work_len = 0;
    return spmv_setup(ctx);
}

// Ghosted CSR for the halo-fused SpMV: row r = its local entries (column order)
// followed by its non-local entries (interface order), the latter with column
// n + receive-window index.  One left-to-right row sum over it performs exactly
// `y = A_local x` followed by `y += A_nonlocal recv` of the reference's
// distributed apply, without any per-tile halo bookkeeping in the kernel.
static int build_ghosted(Context *ctx)
{
    cudaStream_t st = ctx->stream;
    const label n = ctx->n;
    const int64_t nnz_g = ctx->nnz + ctx->n_halo;
    if (nnz_g >= (int64_t)1 << 31) return fail(ctx, OGL_ERR_UNSUPPORTED, "ghosted nnz exceeds label range");
    label *d_cnt = nullptr, *d_before = nullptr;
    int *d_max = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() { cudaFree(d_cnt), cudaFree(d_before), cudaFree(d_max), cudaFree(d_tmp); };
    int rc;
    if ((rc = dev_alloc(ctx, &d_cnt, (size_t)n + 1)) || (rc = dev_alloc(ctx, &d_before, (size_t)n + 1)) ||
        (rc = dev_alloc(ctx, &d_max, 1)) || (rc = dev_alloc(ctx, &ctx->d_g_row_ptrs, (size_t)n + 1)) ||
        (rc = dev_alloc(ctx, &ctx->d_g_cols, (size_t)nnz_g)) ||
        (rc = dev_alloc(ctx, &ctx->d_g_map, (size_t)nnz_g)) ||
        (rc = dev_alloc(ctx, &ctx->d_g_vals, (size_t)nnz_g))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_cnt, 0, sizeof(label) * ((size_t)n + 1), st);
    cudaMemsetAsync(d_max, 0, sizeof(int), st);
    k_ghost_counts<<<grid_for(ctx->n_nl_rows), kThreads, 0, st>>>(ctx->n_nl_rows, ctx->d_nl_row_ids,
                                                                 ctx->d_nl_row_ptrs, d_cnt);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_cnt, d_before, n + 1, st);
    if (cudaMalloc(&d_tmp, scan_bytes + 16) != cudaSuccess) {
        cleanup();
        return fail(ctx, OGL_ERR_CUDA, "cudaMalloc(scan temp)");
    }
    cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, d_cnt, d_before, n + 1, st);
    k_ghost_row_ptrs<<<grid_for((int64_t)n + 1), kThreads, 0, st>>>(n, ctx->d_row_ptrs, d_before,
                                                                   ctx->d_g_row_ptrs);
    int g = grid_for(ctx->nnz);
    if (g > kNumSM * 16) g = kNumSM * 16;
    k_ghost_local<<<g, kThreads, 0, st>>>(ctx->nnz, ctx->d_rows, ctx->d_cols, ctx->d_map, d_before,
                                          ctx->d_g_cols, ctx->d_g_map);
    k_ghost_halo<<<grid_for(ctx->n_nl_rows), kThreads, 0, st>>>(
        n, ctx->n_nl_rows, ctx->d_nl_row_ids, ctx->d_nl_row_ptrs, ctx->d_nl_cols, ctx->d_nl_map,
        ctx->d_g_row_ptrs, ctx->d_g_cols, ctx->d_g_map);
    const int64_t nblk = ((int64_t)n + 255) / 256;
    k_block_span_max<<<(int)((nblk + 255) / 256), 256, 0, st>>>(n, ctx->d_g_row_ptrs, 256, d_max);
    int mx = 0, mx_row = 0;
    cudaMemcpyAsync(&mx, d_max, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaMemsetAsync(d_max, 0, sizeof(int), st);
    k_row_len_max<<<grid_for((int64_t)n), kThreads, 0, st>>>(n, ctx->d_g_row_ptrs, d_max);
    cudaMemcpyAsync(&mx_row, d_max, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess)
        return fail(ctx, OGL_ERR_CUDA, std::string("build_ghosted: ") + cudaGetErrorString(e));
    ctx->max_block_nnz_g = mx;
    ctx->max_row_len_g = mx_row;
    ctx->gell.structure_ready = ctx->gell.ready = false;
    ctx->have_ghosted = true;
    return OGL_OK;
}

int nonlocal_pattern(Context *ctx, label n_halo, const label *face_cells)
{
    if (!ctx->have_pattern)
        return fail(ctx, OGL_ERR_INVALID, "ogl_nonlocal_pattern before ogl_pattern_from_ldu");
    if (n_halo < 0 || (n_halo > 0 && !face_cells))
        return fail(ctx, OGL_ERR_INVALID, "bad halo arguments");
    cudaStream_t st = ctx->stream;
    ctx->n_halo = n_halo;
    ctx->n_nl_rows = 0;
    ctx->have_ghosted = false;
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_rows, n_halo));
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_cols, n_halo));
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_map, n_halo));
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_vals, n_halo));
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_staging, n_halo));
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_row_ids, n_halo));
    OGL_TRY(dev_alloc(ctx, &ctx->d_nl_row_ptrs, (size_t)n_halo + 1));
    ctx->have_nonlocal = true;
    if (n_halo == 0) return OGL_OK;

    label *d_keys = nullptr, *d_iota = nullptr, *d_head = nullptr, *d_scan = nullptr;
    int *d_bad = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_keys), cudaFree(d_iota), cudaFree(d_head), cudaFree(d_scan);
        cudaFree(d_bad), cudaFree(d_tmp);
    };
    int rc = OGL_OK;
    if ((rc = dev_alloc(ctx, &d_keys, n_halo)) || (rc = dev_alloc(ctx, &d_iota, n_halo)) ||
        (rc = dev_alloc(ctx, &d_head, n_halo)) || (rc = dev_alloc(ctx, &d_scan, n_halo)) ||
        (rc = dev_alloc(ctx, &d_bad, 1)) ||
        (rc = upload(ctx, d_keys, face_cells, sizeof(label) * n_halo))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    k_validate_rows<<<grid_for(n_halo), kThreads, 0, st>>>(ctx->n, n_halo, d_keys, d_bad);
    k_iota<<<grid_for(n_halo), kThreads, 0, st>>>(n_halo, d_iota);
    // stable LSD radix sort by row: ties keep the running interface index order
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, ctx->d_nl_rows, d_iota,
                                    ctx->d_nl_cols, n_halo, 0, 32, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_head, d_scan, n_halo, st);
    const size_t tmp_bytes = (sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 16;
    if (cudaMalloc(&d_tmp, tmp_bytes) != cudaSuccess) {
        cleanup();
        return fail(ctx, OGL_ERR_CUDA, "cudaMalloc(sort temp)");
    }
    cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, d_keys, ctx->d_nl_rows, d_iota,
                                    ctx->d_nl_cols, n_halo, 0, 32, st);
    // cols == mapping == running interface index (HostMatrix.C:459-465)
    cudaMemcpyAsync(ctx->d_nl_map, ctx->d_nl_cols, sizeof(label) * n_halo,
                    cudaMemcpyDeviceToDevice, st);
    // group by row for the non-local SpMV
    k_mark_heads<<<grid_for(n_halo), kThreads, 0, st>>>(n_halo, ctx->d_nl_rows, d_head);
    cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, d_head, d_scan, n_halo, st);
    label last_head = 0, last_scan = 0;
    int bad = 0;
    cudaMemcpyAsync(&last_head, d_head + (n_halo - 1), sizeof(label), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last_scan, d_scan + (n_halo - 1), sizeof(label), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cleanup();
        return fail(ctx, OGL_ERR_CUDA, std::string("nonlocal_pattern: ") + cudaGetErrorString(e));
    }
    if (bad) {
        cleanup();
        ctx->have_nonlocal = false;
        return fail(ctx, OGL_ERR_INVALID, "processor faceCells index out of range");
    }
    ctx->n_nl_rows = last_scan + last_head;
    k_compact_heads<<<grid_for(n_halo), kThreads, 0, st>>>(
        n_halo, ctx->d_nl_rows, d_head, d_scan, ctx->d_nl_row_ids, ctx->d_nl_row_ptrs,
        ctx->n_nl_rows);
    e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess)
        return fail(ctx, OGL_ERR_CUDA, std::string("nonlocal_pattern: ") + cudaGetErrorString(e));
    return build_ghosted(ctx);
}

int values_update(Context *ctx, const double *diag, const double *upper,
                  const double *lower, const double *if_bou, const double *nl_bou,
                  double scaling)
{
    if (!ctx->have_pattern)
        return fail(ctx, OGL_ERR_INVALID, "ogl_values_update before ogl_pattern_from_ldu");
    const label n = ctx->n, nf = ctx->n_faces, n_if = ctx->n_local_iface;
    if ((n > 0 && !diag) || (nf > 0 && !upper) || (nf > 0 && !ctx->symmetric && !lower) ||
        (n_if > 0 && !if_bou) || (ctx->n_halo > 0 && !nl_bou))
        return fail(ctx, OGL_ERR_INVALID, "null coefficient pointer");
    cudaStream_t st = ctx->stream;
    // staging layout of HostMatrix.C:644-681: [upper | lower (asym) | diag | iface]
    size_t off = 0;
    OGL_TRY(upload(ctx, ctx->d_staging + off, upper, sizeof(double) * nf));
    off += nf;
    if (!ctx->symmetric) {
        OGL_TRY(upload(ctx, ctx->d_staging + off, lower, sizeof(double) * nf));
        off += nf;
    }
    OGL_TRY(upload(ctx, ctx->d_staging + off, diag, sizeof(double) * n));
    off += n;
    const label iface_base = (label)off;
    OGL_TRY(upload(ctx, ctx->d_staging + off, if_bou, sizeof(double) * n_if));
    int g = grid_for(ctx->nnz);
    if (g > kNumSM * 16) g = kNumSM * 16;
    k_gather_values<<<g, kThreads, 0, st>>>(ctx->nnz, ctx->d_map, ctx->d_staging, iface_base,
                                            scaling, ctx->d_vals);
    ctx->launches++;
    if (ctx->n_halo > 0) {
        if (!ctx->have_nonlocal)
            return fail(ctx, OGL_ERR_INVALID, "non-local coefficients without ogl_nonlocal_pattern");
        OGL_TRY(upload(ctx, ctx->d_nl_staging, nl_bou, sizeof(double) * ctx->n_halo));
        k_gather_nonlocal<<<grid_for(ctx->n_halo), kThreads, 0, st>>>(
            ctx->n_halo, ctx->d_nl_map, ctx->d_nl_staging, scaling, ctx->d_nl_vals);
        ctx->launches++;
        if (ctx->have_ghosted) {
            const int64_t nnz_g = ctx->nnz + ctx->n_halo;
            k_gather_ghosted<<<g, kThreads, 0, st>>>(nnz_g, ctx->d_g_map, ctx->d_staging, iface_base,
                                                     ctx->d_nl_staging, scaling, ctx->d_g_vals);
            ctx->launches++;
        }
    }
    OGL_CUDA(ctx, cudaGetLastError());
    ctx->have_values = true;
    ell_invalidate(ctx, /*structure=*/false);   // the ELL copies (if in use) take the new values on demand
    // the preconditioner is NOT invalidated here: the host layer regenerates it every solve
    // unless `caching N` keeps the (intentionally stale) one for N solves, Preconditioner.H:384-422
    return OGL_OK;
}

}  // namespace ogl
