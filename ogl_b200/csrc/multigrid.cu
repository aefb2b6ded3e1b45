// Algebraic multigrid preconditioner on the LOCAL matrix block: the `preconditioner Multigrid`
// keyword.
//
// Replaces Preconditioner::init_preconditioner_impl("Multigrid") + wrap_schwarz
// (Preconditioner/Preconditioner.H:66-82, 261-341), i.e. Ginkgo's
//   multigrid::Pgm (deterministic)   size-2 aggregation by mutually strongest neighbours
//   solver::Multigrid                one V cycle per apply from a zero guess, at most maxLevels (9)
//                                    coarsenings while a level has more than minCoarseRows (10) rows
//   pre/post smoother                Ir(2 sweeps, relaxation 0.9, scalar Jacobi)
//   coarsest solver                  coarseSolverIters (4) unpreconditioned CG iterations
// generated on distributed::Matrix::get_local_matrix(): no communication in the apply.
//
// Everything is built ON THE DEVICE, every time the preconditioner is (re)generated (the reference
// regenerates it every solve unless `caching` says otherwise):
//   * aggregation: one thread per row per matching round; a round reads the aggregates as they were
//     when it started (snapshot), so the result does not depend on thread scheduling;
//   * Galerkin product A_c = P^T A P with the piecewise-constant P: every entry gets the 64-bit key
//     (coarse row << 32 | coarse column), a STABLE radix sort (CUB) groups equal keys, one thread
//     per coarse entry adds its group in the fine matrix's storage order -- integer structure and
//     FP64 sums are reproducible bit for bit;
//   * restriction as a CSR over the aggregates' members (ascending), prolongation as a gather.
// The V cycle is a fixed sequence of launches (captured into the solver's chunk graphs); the level-0
// SpMVs run on the tuned kernels of spmv.cu / ell.cu, coarse levels on a thread-per-row CSR kernel;
// the coarsest CG keeps its scalars on the device (two-stage deterministic dot products).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

namespace {

constexpr int kT = 256;
constexpr int kDotBlocks = 256;
constexpr double kRelax = 0.9;   // Preconditioner.H:271 with_relaxation_factor(0.9)

int grid_for(int64_t n, int threads = kT) { return (int)((n + threads - 1) / threads); }
int capped_grid(int64_t n) { const int g = grid_for(n); return g > kNumSM * 8 ? kNumSM * 8 : (g < 1 ? 1 : g); }

// ---- setup kernels --------------------------------------------------------------------------------

// w_e = (|a_rc| + |a_cr|) / 2 for e = (r,c)  [0.5 |A| + 0.5 |A|^T];  diag_r = w_rr
__global__ void k_mg_weights(label n, const label *__restrict__ rp, const label *__restrict__ cols,
                             const double *__restrict__ vals, double *__restrict__ w, double *__restrict__ diag)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool have_diag = false;
    double d = 0.0;
    for (label e = rp[i]; e < rp[i + 1]; ++e) {
        const label c = cols[e];
        double t = 0.0;
        for (label q = rp[c]; q < rp[c + 1]; ++q)
            if (cols[q] == i) {
                t = fabs(vals[q]);
                break;
            }
        const double we = __dadd_rn(__dmul_rn(0.5, fabs(vals[e])), __dmul_rn(0.5, t));
        w[e] = we;
        if (c == i && !have_diag) {
            d = we;
            have_diag = true;
        }
    }
    diag[i] = d;
}

__device__ __forceinline__ double mg_weight(const double *w, const double *diag, label row, label col, label e)
{
    return w[e] / fmax(fabs(diag[row]), fabs(diag[col]));
}

// pgm::find_strongest_neighbor on a snapshot of the aggregates
__global__ void k_mg_find_strongest(label n, const label *__restrict__ rp, const label *__restrict__ cols,
                                    const double *__restrict__ w, const double *__restrict__ diag,
                                    const label *__restrict__ snapshot, label *agg, label *strongest)
{
    const label row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n || snapshot[row] != -1) return;
    double max_unagg = 0.0, max_agg = 0.0;
    label s_unagg = -1, s_agg = -1;
    for (label e = rp[row]; e < rp[row + 1]; ++e) {
        const label c = cols[e];
        if (c == row) continue;
        const double wt = mg_weight(w, diag, row, c, e);
        const bool c_unagg = snapshot[c] == -1;
        if (c_unagg && (wt > max_unagg || (wt == max_unagg && c > s_unagg))) {
            max_unagg = wt;
            s_unagg = c;
        } else if (!c_unagg && (wt > max_agg || (wt == max_agg && c > s_agg))) {
            max_agg = wt;
            s_agg = c;
        }
    }
    if (s_unagg == -1 && s_agg != -1) agg[row] = snapshot[s_agg];
    else if (s_unagg != -1) strongest[row] = s_unagg;
    else strongest[row] = row;
}

// pgm::match_edge: mutually strongest pairs; the smaller index names the aggregate (pairs are disjoint)
__global__ void k_mg_match(label n, label *agg, const label *__restrict__ strongest)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || agg[i] != -1) return;
    const label nb = strongest[i];
    if (nb != -1 && strongest[nb] == i && i <= nb) {
        agg[i] = i;
        agg[nb] = i;
    }
}

__global__ void k_mg_count_unagg(label n, const label *__restrict__ agg, int *count)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    const int un = (i < n && agg[i] == -1) ? 1 : 0;
    const int total = __syncthreads_count(un);
    if (threadIdx.x == 0 && total) atomicAdd(count, total);
}

// pgm::assign_to_exist_agg (deterministic: reads the snapshot)
__global__ void k_mg_assign(label n, const label *__restrict__ rp, const label *__restrict__ cols,
                            const double *__restrict__ w, const double *__restrict__ diag,
                            const label *__restrict__ snapshot, label *agg)
{
    const label row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n || snapshot[row] != -1) return;
    double max_agg = 0.0;
    label s_agg = -1;
    for (label e = rp[row]; e < rp[row + 1]; ++e) {
        const label c = cols[e];
        if (c == row) continue;
        const double wt = mg_weight(w, diag, row, c, e);
        if (snapshot[c] != -1 && (wt > max_agg || (wt == max_agg && c > s_agg))) {
            max_agg = wt;
            s_agg = c;
        }
    }
    agg[row] = s_agg != -1 ? snapshot[s_agg] : row;
}

__global__ void k_mg_fill_label(int64_t n, label *v, label value)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = value;
}

__global__ void k_mg_mark(label n, const label *__restrict__ agg, label *map)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) map[agg[i]] = 1;
}

__global__ void k_mg_renumber(label n, label *agg, const label *__restrict__ map_scan)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) agg[i] = map_scan[agg[i]];
}

__global__ void k_mg_keys(int64_t nnz, const label *__restrict__ rows, const label *__restrict__ cols,
                          const label *__restrict__ agg, unsigned long long *keys, label *idx)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    keys[e] = ((unsigned long long)(unsigned int)agg[rows[e]] << 32) | (unsigned int)agg[cols[e]];
    idx[e] = (label)e;
}

template <typename K>
__global__ void k_mg_heads(int64_t n, const K *__restrict__ keys, label *head)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) head[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

// starts[u] = first sorted position of unique key u; starts[n_unique] = n
__global__ void k_mg_starts(int64_t n, const label *__restrict__ head, const label *__restrict__ upos,
                            label *starts, label n_unique)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (head[k]) starts[upos[k]] = (label)k;
    if (k == n - 1) starts[n_unique] = (label)n;
}

// one thread per coarse entry: add its group in sorted (= fine storage) order
__global__ void k_mg_coarse_entries(label nnz_c, const label *__restrict__ starts,
                                    const unsigned long long *__restrict__ keys, const label *__restrict__ idx,
                                    const double *__restrict__ vals, label *rows_c, label *cols_c, double *vals_c)
{
    const label u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nnz_c) return;
    const label lo = starts[u], hi = starts[u + 1];
    double s = vals[idx[lo]];
    for (label k = lo + 1; k < hi; ++k) s = __dadd_rn(s, vals[idx[k]]);
    const unsigned long long key = keys[lo];
    rows_c[u] = (label)(key >> 32);
    cols_c[u] = (label)(key & 0xffffffffull);
    vals_c[u] = s;
}

// row pointers from sorted row indices in which every row 0..n_rows-1 occurs
__global__ void k_mg_row_ptrs(int64_t n_entries, const label *__restrict__ rows, label *rp, label n_rows)
{
    const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n_entries) return;
    if (u == 0 || rows[u] != rows[u - 1]) rp[rows[u]] = (label)u;
    if (u == n_entries - 1) rp[n_rows] = (label)n_entries;
}

__global__ void k_mg_inv_diag(label n, const label *__restrict__ rp, const label *__restrict__ cols,
                              const double *__restrict__ vals, double *inv_diag)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    for (label e = rp[i]; e < rp[i + 1]; ++e)
        if (cols[e] == i) {
            d = vals[e];
            break;
        }
    inv_diag[i] = 1.0 / d;
}

__global__ void k_mg_iota(label n, label *v)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// ---- cycle kernels (all skipped once the solve is done, like every kernel of the loop) ------------

struct MgK {
    label n;
    const label *rp, *cols;
    const double *vals, *inv_diag;
    const double *b;
    const double *in;     // x of an SpMV / coarse correction / second operand
    double *x, *out;
    const label *agg, *r_ptr, *r_idx;
    double *scal, *partials;
    SolveState *state;
    int guard_done;
};

#define MG_GUARD(a) \
    if ((a).guard_done && (a).state->done) return
#define MG_STRIDE(i, n) \
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__device__ __forceinline__ double mg_row_sum(const MgK &a, int64_t i, const double *x)
{
    double s = 0.0;
    for (label e = a.rp[i]; e < a.rp[i + 1]; ++e) s = __dadd_rn(s, __dmul_rn(a.vals[e], x[a.cols[e]]));
    return s;
}

// out = A in
__global__ void __launch_bounds__(kT) k_mg_spmv(const MgK a)
{
    MG_GUARD(a);
    MG_STRIDE(i, a.n) a.out[i] = mg_row_sum(a, i, a.in);
}

// first smoothing sweep from a zero guess: x = (0.9 b) / d
__global__ void __launch_bounds__(kT) k_mg_jacobi_first(const MgK a)
{
    MG_GUARD(a);
    MG_STRIDE(i, a.n) a.x[i] = __dmul_rn(__dmul_rn(kRelax, a.b[i]), a.inv_diag[i]);
}

// x += (0.9 (b - in)) / d   with in = A x
__global__ void __launch_bounds__(kT) k_mg_jacobi_update(const MgK a)
{
    MG_GUARD(a);
    MG_STRIDE(i, a.n)
    {
        const double res = __dsub_rn(a.b[i], a.in[i]);
        a.x[i] = __dadd_rn(a.x[i], __dmul_rn(__dmul_rn(kRelax, res), a.inv_diag[i]));
    }
}

// out = b - in
__global__ void __launch_bounds__(kT) k_mg_residual(const MgK a)
{
    MG_GUARD(a);
    MG_STRIDE(i, a.n) a.out[i] = __dsub_rn(a.b[i], a.in[i]);
}

// restriction: out_I = sum of in over the members of aggregate I (ascending); also zeroes x_I (the coarse guess)
__global__ void __launch_bounds__(kT) k_mg_restrict(const MgK a)
{
    MG_GUARD(a);
    MG_STRIDE(I, a.n)
    {
        double s = 0.0;
        for (label k = a.r_ptr[I]; k < a.r_ptr[I + 1]; ++k) s = __dadd_rn(s, a.in[a.r_idx[k]]);
        a.out[I] = s;
        a.x[I] = 0.0;
    }
}

// prolongation: x_i += in_{agg(i)}
__global__ void __launch_bounds__(kT) k_mg_prolong(const MgK a)
{
    MG_GUARD(a);
    MG_STRIDE(i, a.n) a.x[i] = __dadd_rn(a.x[i], a.in[a.agg[i]]);
}

__device__ __forceinline__ void mg_block_sum(double v, double *partials)
{
    __shared__ double sh[kT / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < kT / 32 ? sh[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) partials[blockIdx.x] = t;
    }
}

// coarsest CG, start: x = 0, r = b (out), p = 0 (in is unused), partial <r,r>
__global__ void __launch_bounds__(kT) k_mg_cg_init(const MgK a, double *p)
{
    MG_GUARD(a);
    double acc = 0.0;
    MG_STRIDE(i, a.n)
    {
        const double r = a.b[i];
        a.x[i] = 0.0;
        a.out[i] = r;
        p[i] = 0.0;
        acc += __dmul_rn(r, r);
    }
    mg_block_sum(acc, a.partials);
}

// mode 0: scal[0] = sum, scal[1] = 1 (rho, prev_rho at the start)
// mode 1: scal[2] = sum (beta)
// mode 2: scal[1] = scal[0], scal[0] = sum (prev_rho = rho, rho = <r,r>)
__global__ void __launch_bounds__(kT) k_mg_dot_final(const MgK a, int n_partials, int mode)
{
    MG_GUARD(a);
    __shared__ double sh[kT / 32];
    double v = 0.0;
    for (int k = threadIdx.x; k < n_partials; k += blockDim.x) v += a.partials[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int k = 0; k < kT / 32; ++k) total += sh[k];
        if (mode == 0) {
            a.scal[0] = total;
            a.scal[1] = 1.0;
        } else if (mode == 1) {
            a.scal[2] = total;
        } else {
            a.scal[1] = a.scal[0];
            a.scal[0] = total;
        }
    }
}

// cg::step_1: p = r + (rho / prev_rho) p   (p = r when prev_rho == 0);  in = r, x = p
__global__ void __launch_bounds__(kT) k_mg_cg_p(const MgK a)
{
    MG_GUARD(a);
    const double rho = a.scal[0], prev = a.scal[1];
    const bool restart = prev == 0.0;
    const double t = restart ? 0.0 : rho / prev;
    MG_STRIDE(i, a.n) a.x[i] = restart ? a.in[i] : __dadd_rn(a.in[i], __dmul_rn(t, a.x[i]));
}

// partial <in, b'> with b' = a.b (two read-only vectors)
__global__ void __launch_bounds__(kT) k_mg_dot(const MgK a)
{
    MG_GUARD(a);
    double acc = 0.0;
    MG_STRIDE(i, a.n) acc += __dmul_rn(a.in[i], a.b[i]);
    mg_block_sum(acc, a.partials);
}

// cg::step_2: x += (rho / beta) p, r -= (rho / beta) q (skipped when beta == 0); partial <r,r>
//   a.x = x, a.out = r, a.in = p, a.b = q
__global__ void __launch_bounds__(kT) k_mg_cg_xr(const MgK a)
{
    MG_GUARD(a);
    const double rho = a.scal[0], beta = a.scal[2];
    const bool skip = beta == 0.0;
    const double al = skip ? 0.0 : rho / beta;
    double acc = 0.0;
    MG_STRIDE(i, a.n)
    {
        double r = a.out[i];
        if (!skip) {
            a.x[i] = __dadd_rn(a.x[i], __dmul_rn(al, a.in[i]));
            r = __dsub_rn(r, __dmul_rn(al, a.b[i]));
            a.out[i] = r;
        }
        acc += __dmul_rn(r, r);
    }
    mg_block_sum(acc, a.partials);
}

template <typename T>
int mg_alloc(Context *ctx, T **p, size_t count)
{
    if (count == 0) count = 1;
    OGL_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T) + 64));
    return OGL_OK;
}

void free_level(Context::MgLevel &L, bool owns_matrix)
{
    void *own[] = {L.inv_diag, L.agg, L.r_ptr, L.r_idx, L.b, L.x, L.r, L.p, L.q};
    for (void *p : own)
        if (p) cudaFree(p);
    if (owns_matrix) {
        void *m[] = {L.rp, L.cols, L.rows, L.vals};
        for (void *p : m)
            if (p) cudaFree(p);
    }
    L = Context::MgLevel{};
}

// aggregation + Galerkin product of level `l`; appends level l + 1.  Returns through `shrunk` whether
// the coarse matrix is smaller than the fine one.
int coarsen_level(Context *ctx, size_t l, bool *shrunk)
{
    Context::Multigrid &M = ctx->mg;
    cudaStream_t st = ctx->stream;
    const label n = M.levels[l].n;
    const int64_t nnz = M.levels[l].nnz;
    const label *rp = M.levels[l].rp, *cols = M.levels[l].cols, *rows = M.levels[l].rows;
    const double *vals = M.levels[l].vals;
    *shrunk = false;

    double *w = nullptr, *diag = nullptr;
    label *agg = nullptr, *snapshot = nullptr, *strongest = nullptr, *map = nullptr, *map_scan = nullptr;
    label *idx = nullptr, *idx_s = nullptr, *head = nullptr, *upos = nullptr, *starts = nullptr;
    label *agg_s = nullptr, *iota = nullptr;
    unsigned long long *keys = nullptr, *keys_s = nullptr;
    int *d_count = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        void *ptrs[] = {w, diag, snapshot, strongest, map, map_scan, idx, idx_s, head, upos, starts, agg_s, iota,
                        keys, keys_s, d_count, d_tmp};
        for (void *p : ptrs)
            if (p) cudaFree(p);
    };
#define MG_TRY(expr)              \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != OGL_OK) {     \
            cleanup();            \
            if (agg) cudaFree(agg); \
            return rc__;          \
        }                         \
    } while (0)
#define MG_CUDA(expr)                                                                            \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            cleanup();                                                                           \
            if (agg) cudaFree(agg);                                                              \
            return fail(ctx, OGL_ERR_CUDA, std::string("multigrid setup: ") + cudaGetErrorString(e__)); \
        }                                                                                        \
    } while (0)

    MG_TRY(mg_alloc(ctx, &w, (size_t)nnz));
    MG_TRY(mg_alloc(ctx, &diag, (size_t)n));
    MG_TRY(mg_alloc(ctx, &agg, (size_t)n));
    MG_TRY(mg_alloc(ctx, &snapshot, (size_t)n));
    MG_TRY(mg_alloc(ctx, &strongest, (size_t)n));
    MG_TRY(mg_alloc(ctx, &d_count, 1));
    const int g_rows = grid_for(n);
    k_mg_weights<<<g_rows, kT, 0, st>>>(n, rp, cols, vals, w, diag);
    k_mg_fill_label<<<g_rows, kT, 0, st>>>(n, agg, -1);
    k_mg_fill_label<<<g_rows, kT, 0, st>>>(n, strongest, -1);
    ctx->launches += 3;
    int num_unagg = 0, num_unagg_prev = 0;
    for (int it = 0; it < 15; ++it) {   // Pgm max_iterations
        MG_CUDA(cudaMemcpyAsync(snapshot, agg, sizeof(label) * n, cudaMemcpyDeviceToDevice, st));
        k_mg_find_strongest<<<g_rows, kT, 0, st>>>(n, rp, cols, w, diag, snapshot, agg, strongest);
        k_mg_match<<<g_rows, kT, 0, st>>>(n, agg, strongest);
        MG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
        k_mg_count_unagg<<<g_rows, kT, 0, st>>>(n, agg, d_count);
        ctx->launches += 3;
        MG_CUDA(cudaMemcpyAsync(&num_unagg, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        MG_CUDA(cudaStreamSynchronize(st));
        // Pgm max_unassigned_ratio 0.05
        if (num_unagg == 0 || num_unagg == num_unagg_prev || num_unagg < 0.05 * n) break;
        num_unagg_prev = num_unagg;
    }
    if (num_unagg != 0) {
        MG_CUDA(cudaMemcpyAsync(snapshot, agg, sizeof(label) * n, cudaMemcpyDeviceToDevice, st));
        k_mg_assign<<<g_rows, kT, 0, st>>>(n, rp, cols, w, diag, snapshot, agg);
        ctx->launches++;
    }
    // renumber: flags of the aggregates' root rows -> exclusive scan -> new names
    MG_TRY(mg_alloc(ctx, &map, (size_t)n + 1));
    MG_TRY(mg_alloc(ctx, &map_scan, (size_t)n + 1));
    MG_CUDA(cudaMemsetAsync(map, 0, sizeof(label) * ((size_t)n + 1), st));
    k_mg_mark<<<g_rows, kT, 0, st>>>(n, agg, map);
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, map, map_scan, n + 1, st);
    MG_CUDA(cudaMalloc(&d_tmp, bytes + 16));
    MG_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, bytes, map, map_scan, n + 1, st));
    k_mg_renumber<<<g_rows, kT, 0, st>>>(n, agg, map_scan);
    ctx->launches += 2;
    label n_c = 0;
    MG_CUDA(cudaMemcpyAsync(&n_c, map_scan + n, sizeof(label), cudaMemcpyDeviceToHost, st));
    MG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_tmp);
    d_tmp = nullptr;
    if (n_c < 1 || n_c > n) MG_CUDA(cudaErrorUnknown);
    if (n_c == n) {   // Multigrid::generate stops at a coarsening that does not shrink the matrix
        cleanup();
        cudaFree(agg);
        return OGL_OK;
    }
    // Galerkin product: keys, stable sort, group sums
    MG_TRY(mg_alloc(ctx, &keys, (size_t)nnz));
    MG_TRY(mg_alloc(ctx, &keys_s, (size_t)nnz));
    MG_TRY(mg_alloc(ctx, &idx, (size_t)nnz));
    MG_TRY(mg_alloc(ctx, &idx_s, (size_t)nnz));
    MG_TRY(mg_alloc(ctx, &head, (size_t)nnz));
    MG_TRY(mg_alloc(ctx, &upos, (size_t)nnz));
    const int g_nnz = grid_for(nnz);
    k_mg_keys<<<g_nnz, kT, 0, st>>>(nnz, rows, cols, agg, keys, idx);
    bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys_s, idx, idx_s, (int)nnz, 0, 64, st);
    MG_CUDA(cudaMalloc(&d_tmp, bytes + 16));
    MG_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, keys, keys_s, idx, idx_s, (int)nnz, 0, 64, st));
    cudaFree(d_tmp);
    d_tmp = nullptr;
    k_mg_heads<unsigned long long><<<g_nnz, kT, 0, st>>>(nnz, keys_s, head);
    bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, head, upos, (int)nnz, st);
    MG_CUDA(cudaMalloc(&d_tmp, bytes + 16));
    MG_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, bytes, head, upos, (int)nnz, st));
    cudaFree(d_tmp);
    d_tmp = nullptr;
    label last_pos = 0, last_head = 0;
    MG_CUDA(cudaMemcpyAsync(&last_pos, upos + (nnz - 1), sizeof(label), cudaMemcpyDeviceToHost, st));
    MG_CUDA(cudaMemcpyAsync(&last_head, head + (nnz - 1), sizeof(label), cudaMemcpyDeviceToHost, st));
    MG_CUDA(cudaStreamSynchronize(st));
    const label nnz_c = last_pos + last_head;
    ctx->launches += 3;
    MG_TRY(mg_alloc(ctx, &starts, (size_t)nnz_c + 1));
    k_mg_starts<<<g_nnz, kT, 0, st>>>(nnz, head, upos, starts, nnz_c);
    Context::MgLevel C;
    C.n = n_c;
    C.nnz = nnz_c;
    int rc = mg_alloc(ctx, &C.rp, (size_t)n_c + 1);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.cols, (size_t)nnz_c);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.rows, (size_t)nnz_c);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.vals, (size_t)nnz_c);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.inv_diag, (size_t)n_c);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.b, (size_t)n_c);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.x, (size_t)n_c);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &C.r, (size_t)n_c);
    if (rc != OGL_OK) {
        free_level(C, true);
        MG_TRY(rc);
    }
    k_mg_coarse_entries<<<grid_for(nnz_c), kT, 0, st>>>(nnz_c, starts, keys_s, idx_s, vals, C.rows, C.cols, C.vals);
    k_mg_row_ptrs<<<grid_for(nnz_c), kT, 0, st>>>(nnz_c, C.rows, C.rp, n_c);
    k_mg_inv_diag<<<grid_for(n_c), kT, 0, st>>>(n_c, C.rp, C.cols, C.vals, C.inv_diag);
    ctx->launches += 4;
    // restriction CSR: the members of every aggregate, ascending (stable sort of the rows by aggregate)
    Context::MgLevel &F = M.levels[l];
    rc = mg_alloc(ctx, &F.r_ptr, (size_t)n_c + 1);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &F.r_idx, (size_t)n);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &agg_s, (size_t)n);
    if (rc == OGL_OK) rc = mg_alloc(ctx, &iota, (size_t)n);
    if (rc != OGL_OK) {
        free_level(C, true);
        MG_TRY(rc);
    }
    k_mg_iota<<<g_rows, kT, 0, st>>>(n, iota);
    bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, agg, agg_s, iota, F.r_idx, (int)n, 0, 32, st);
    cudaError_t ce = cudaMalloc(&d_tmp, bytes + 16);
    if (ce == cudaSuccess)
        ce = cub::DeviceRadixSort::SortPairs(d_tmp, bytes, agg, agg_s, iota, F.r_idx, (int)n, 0, 32, st);
    if (ce == cudaSuccess) {
        k_mg_row_ptrs<<<g_rows, kT, 0, st>>>(n, agg_s, F.r_ptr, n_c);
        ce = cudaStreamSynchronize(st);
    }
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        free_level(C, true);
        MG_CUDA(ce);
    }
    ctx->launches += 3;
    F.agg = agg;
    F.n_coarse = n_c;
    agg = nullptr;
    cleanup();
    M.levels.push_back(C);
    *shrunk = true;
    return OGL_OK;
#undef MG_TRY
#undef MG_CUDA
}

MgK level_args(Context *ctx, const Context::MgLevel &L, bool guard)
{
    MgK a{};
    a.n = L.n;
    a.rp = L.rp;
    a.cols = L.cols;
    a.vals = L.vals;
    a.inv_diag = L.inv_diag;
    a.agg = L.agg;
    a.r_ptr = L.r_ptr;
    a.r_idx = L.r_idx;
    a.scal = ctx->mg.scal;
    a.partials = ctx->mg.partials;
    a.state = ctx->d_state;
    a.guard_done = guard ? 1 : 0;
    return a;
}

// out = A_l in: level 0 on the tuned kernels of the solver (same row order, same bits)
int level_spmv(Context *ctx, size_t l, const double *in, double *out, bool guard)
{
    if (l == 0) {
        SpmvArgs s;
        s.x = in;
        s.y = out;
        s.guard_done = guard;
        return spmv_local(ctx, s);
    }
    MgK a = level_args(ctx, ctx->mg.levels[l], guard);
    a.in = in;
    a.out = out;
    k_mg_spmv<<<capped_grid(a.n), kT, 0, ctx->stream>>>(a);
    ctx->launches++;
    return OGL_OK;
}

// Ir(2, 0.9, Jacobi): two sweeps; from a zero guess the first one needs no SpMV
int smooth(Context *ctx, size_t l, const double *b, double *x, double *tmp, bool x_is_zero, bool guard)
{
    MgK a = level_args(ctx, ctx->mg.levels[l], guard);
    a.b = b;
    a.x = x;
    a.in = tmp;
    const int grid = capped_grid(a.n);
    for (int it = 0; it < 2; ++it) {
        if (it == 0 && x_is_zero) {
            k_mg_jacobi_first<<<grid, kT, 0, ctx->stream>>>(a);
            ctx->launches++;
            continue;
        }
        OGL_TRY(level_spmv(ctx, l, x, tmp, guard));
        k_mg_jacobi_update<<<grid, kT, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    return OGL_OK;
}

// coarseSolverIters CG iterations on level l from a zero guess: x <- ~ A_l^-1 b
int coarse_cg(Context *ctx, size_t l, const double *b, double *x, bool guard)
{
    Context::MgLevel &L = ctx->mg.levels[l];
    cudaStream_t st = ctx->stream;
    MgK a = level_args(ctx, L, guard);
    int grid = capped_grid(a.n);
    if (grid > kDotBlocks) grid = kDotBlocks;
    double *r = L.r, *p = L.p, *q = L.q;
    a.b = b;
    a.x = x;
    a.out = r;
    k_mg_cg_init<<<grid, kT, 0, st>>>(a, p);
    k_mg_dot_final<<<1, kT, 0, st>>>(a, grid, 0);
    ctx->launches += 2;
    for (int it = 0; it < (int)ctx->mg_coarse_iters; ++it) {
        MgK pa = a;
        pa.in = r;
        pa.x = p;
        k_mg_cg_p<<<grid, kT, 0, st>>>(pa);
        ctx->launches++;
        OGL_TRY(level_spmv(ctx, l, p, q, guard));
        MgK da = a;
        da.in = p;
        da.b = q;
        k_mg_dot<<<grid, kT, 0, st>>>(da);
        k_mg_dot_final<<<1, kT, 0, st>>>(a, grid, 1);
        MgK ua = a;
        ua.x = x;
        ua.out = r;
        ua.in = p;
        ua.b = q;
        k_mg_cg_xr<<<grid, kT, 0, st>>>(ua);
        k_mg_dot_final<<<1, kT, 0, st>>>(a, grid, 2);
        ctx->launches += 4;
    }
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace

void mg_release(Context *ctx)
{
    Context::Multigrid &M = ctx->mg;
    for (size_t l = 0; l < M.levels.size(); ++l) free_level(M.levels[l], l > 0);
    M.levels.clear();
    if (M.scal) cudaFree(M.scal);
    if (M.partials) cudaFree(M.partials);
    M.scal = M.partials = nullptr;
    M.ready = false;
}

int mg_setup(Context *ctx)
{
    if (ctx->n == 0) return OGL_OK;
    if (ctx->nnz > 0x7fffffffLL) return fail(ctx, OGL_ERR_UNSUPPORTED, "Multigrid: more than 2^31 entries per rank");
    invalidate_graph(ctx);   // the chunk graph holds the old hierarchy's addresses
    mg_release(ctx);
    Context::Multigrid &M = ctx->mg;
    cudaStream_t st = ctx->stream;
    OGL_TRY(mg_alloc(ctx, &M.scal, 8));
    OGL_TRY(mg_alloc(ctx, &M.partials, (size_t)kDotBlocks + 8));
    Context::MgLevel L0;
    L0.n = ctx->n;
    L0.nnz = ctx->nnz;
    L0.rp = ctx->d_row_ptrs;
    L0.cols = ctx->d_cols;
    L0.rows = ctx->d_rows;
    L0.vals = ctx->d_vals;
    OGL_TRY(mg_alloc(ctx, &L0.inv_diag, (size_t)ctx->n));
    k_mg_inv_diag<<<grid_for(ctx->n), kT, 0, st>>>(ctx->n, L0.rp, L0.cols, L0.vals, L0.inv_diag);
    ctx->launches++;
    M.levels.push_back(L0);
    // Multigrid::generate: coarsen while level < max_levels and rows > min_coarse_rows
    int level = 0;
    while (level < (int)ctx->mg_max_levels && M.levels.back().n > (label)ctx->mg_min_coarse_rows) {
        bool shrunk = false;
        const int rc = coarsen_level(ctx, M.levels.size() - 1, &shrunk);
        if (rc != OGL_OK) {
            mg_release(ctx);
            return rc;
        }
        if (!shrunk) break;
        ++level;
    }
    // the coarsest level's CG vectors (p, q; r is every level's residual vector).  Level 0 uses the
    // caller's vectors as b / x and a solver work vector as r.
    Context::MgLevel &Lc = M.levels.back();
    OGL_TRY(mg_alloc(ctx, &Lc.p, (size_t)Lc.n));
    OGL_TRY(mg_alloc(ctx, &Lc.q, (size_t)Lc.n));
    if (M.levels.size() == 1) OGL_TRY(mg_alloc(ctx, &Lc.r, (size_t)Lc.n));
    OGL_CUDA(ctx, cudaStreamSynchronize(st));
    OGL_CUDA(ctx, cudaGetLastError());
    M.ready = true;
    return OGL_OK;
}

// before a solve (never inside a graph capture): work vector of level 0, and whatever the level-0
// SpMV kernel prepares lazily (ELL copy of the current coefficients)
int mg_ensure(Context *ctx)
{
    if (ctx->n == 0) return OGL_OK;
    if (!ctx->mg.ready) return fail(ctx, OGL_ERR_INVALID, "Multigrid apply before its generation");
    if (ctx->mg.levels[0].nnz != ctx->nnz || ctx->mg.levels[0].n != ctx->n)
        return fail(ctx, OGL_ERR_INVALID, "cached Multigrid hierarchy does not match the regenerated sparsity pattern");
    // a regenerated pattern of the same mesh reallocates the level-0 arrays
    Context::MgLevel &L0 = ctx->mg.levels[0];
    if (L0.rp != ctx->d_row_ptrs || L0.cols != ctx->d_cols || L0.vals != ctx->d_vals || L0.rows != ctx->d_rows) {
        invalidate_graph(ctx);
        L0.rp = ctx->d_row_ptrs;
        L0.cols = ctx->d_cols;
        L0.rows = ctx->d_rows;
        L0.vals = ctx->d_vals;
    }
    double *t0, *t1;
    OGL_TRY(get_work(ctx, 12, &t0));
    OGL_TRY(get_work(ctx, 13, &t1));
    OGL_CUDA(ctx, cudaMemsetAsync(t1, 0, sizeof(double) * ctx->n, ctx->stream));
    SpmvArgs s;
    s.x = t1;
    s.y = t0;
    return spmv_local(ctx, s);
}

// z = one V cycle on A z = r from z = 0
int mg_apply(Context *ctx, const double *r, double *z, bool guard)
{
    if (ctx->n == 0) return OGL_OK;
    Context::Multigrid &M = ctx->mg;
    if (!M.ready) return fail(ctx, OGL_ERR_INVALID, "Multigrid apply before its generation");
    cudaStream_t st = ctx->stream;
    const size_t nl = M.levels.size();
    double *t0;
    OGL_TRY(get_work(ctx, 12, &t0));
    if (nl == 1) {
        // no coarsening happened: the coarsest solver is the whole preconditioner
        return coarse_cg(ctx, 0, r, z, guard);
    }
    // downward leg
    for (size_t l = 0; l + 1 < nl; ++l) {
        Context::MgLevel &L = M.levels[l];
        const double *b = l == 0 ? r : L.b;
        double *x = l == 0 ? z : L.x;
        double *tmp = l == 0 ? t0 : L.r;
        OGL_TRY(smooth(ctx, l, b, x, tmp, true, guard));
        OGL_TRY(level_spmv(ctx, l, x, tmp, guard));
        MgK a = level_args(ctx, L, guard);
        a.b = b;
        a.in = tmp;
        a.out = tmp;
        k_mg_residual<<<capped_grid(a.n), kT, 0, st>>>(a);
        Context::MgLevel &C = M.levels[l + 1];
        MgK ra = level_args(ctx, L, guard);
        ra.n = C.n;
        ra.in = tmp;
        ra.out = C.b;
        ra.x = C.x;
        k_mg_restrict<<<capped_grid(C.n), kT, 0, st>>>(ra);
        ctx->launches += 2;
    }
    // coarsest solve
    OGL_TRY(coarse_cg(ctx, nl - 1, M.levels[nl - 1].b, M.levels[nl - 1].x, guard));
    // upward leg
    for (size_t l = nl - 1; l-- > 0;) {
        Context::MgLevel &L = M.levels[l];
        const double *b = l == 0 ? r : L.b;
        double *x = l == 0 ? z : L.x;
        double *tmp = l == 0 ? t0 : L.r;
        MgK a = level_args(ctx, L, guard);
        a.x = x;
        a.in = M.levels[l + 1].x;
        k_mg_prolong<<<capped_grid(a.n), kT, 0, st>>>(a);
        ctx->launches++;
        OGL_TRY(smooth(ctx, l, b, x, tmp, false, guard));
    }
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

int mg_level_info(Context *ctx, int level, label *n, label *nnz, label *n_coarse)
{
    const Context::Multigrid &M = ctx->mg;
    if (!M.ready || level < 0 || level >= (int)M.levels.size())
        return fail(ctx, OGL_ERR_INVALID, "no such Multigrid level");
    *n = M.levels[level].n;
    *nnz = (label)M.levels[level].nnz;
    *n_coarse = M.levels[level].agg ? M.levels[level].n_coarse : 0;
    return OGL_OK;
}

int mg_level_download(Context *ctx, int level, label *rp, label *cols, double *vals, label *agg)
{
    const Context::Multigrid &M = ctx->mg;
    if (!M.ready || level < 0 || level >= (int)M.levels.size())
        return fail(ctx, OGL_ERR_INVALID, "no such Multigrid level");
    const Context::MgLevel &L = M.levels[level];
    if (rp) OGL_TRY(download(ctx, rp, L.rp, sizeof(label) * ((size_t)L.n + 1)));
    if (cols) OGL_TRY(download(ctx, cols, L.cols, sizeof(label) * (size_t)L.nnz));
    if (vals) OGL_TRY(download(ctx, vals, L.vals, sizeof(double) * (size_t)L.nnz));
    if (agg && L.agg) OGL_TRY(download(ctx, agg, L.agg, sizeof(label) * (size_t)L.n));
    return OGL_OK;
}

}  // namespace ogl
