// ELL-family SpMV kernels for sm_100a (spmv_variant 7, `matrixFormat Ell`; SURVEY.md section 8
// a20 and f3, reference: DevicePersistent/CsrMatrixWrapper/CsrMatrixWrapper.H:140-160 builds the
// local and the non-local block in the chosen format).
//
// Layout: slot j of row r at [j * pitch + r], so a warp reads 32 consecutive values per slot and
// gathers x from (for a mesh-ordered matrix) 32 consecutive addresses; no row pointers, no shared
// memory staging, no barrier.  Slots are added left to right in the row's CSR order and padding is
// skipped, products are rounded before the add: bit-identical to the CSR kernels and to the
// sequential reference-executor row sum.
//
// The path is HBM-bound, so the kernels go after the bytes:
//   * pattern-coded columns.  On a mesh-ordered matrix almost every row has one of a handful of
//     (column - row) tuples (7-point stencil: interior, 6 faces, 12 edges, 8 corners = 27).  A
//     device-side hash table finds the distinct tuples once per sparsity pattern; rows whose tuple
//     is among the first 255 carry a ONE-BYTE code, the kernel rebuilds the columns from a table in
//     shared memory: 8 B values + 1/width B per entry instead of 12 B.  Other rows (ghost columns of
//     a decomposed case, irregular rows) keep their 4-byte columns behind the escape code 255.
//     Nothing about the arithmetic changes.  Unstructured patterns (> 25% escapes) stay plain ELL.
//   * CG step_1 fused into the SpMV (k_spmv_ell_cgp): the kernel gathers z and p, forms
//     p' = z + (rho/rho_prev) p per operand on the fly -- the same two operations on the same two
//     numbers as the p-update kernel, hence the same bits --, the row owner stores p'[row].  On
//     several ranks the ghost operands come straight from the stamped words the neighbours'
//     x/r-update kernels pushed into this rank's window, and the (single) row that references a
//     ghost column maintains the ghost entry of p.  A PCG iteration is two launches.
#include <climits>
#include <cstring>

#include "spmv.cuh"

namespace ogl {

int spmv_variant_in_use(const Context *ctx);

namespace {

constexpr label kPadDelta = INT_MIN;   // table entry of a padding slot
constexpr int kPatSlots = 1024;        // open-addressing hash table of tuples
constexpr int kPatMaxW = 8;            // widest row the tuple table holds
constexpr int kEscape = 255;
constexpr int kEllBatch = 8;
constexpr int kEllThreads = 256;
constexpr int kEllCtasPerSM = 4;

struct EllK {
    const label *cols;
    const double *vals;
    const unsigned char *code;
    const label *ptab;
    int64_t pitch;
    int width, n_patterns;
};

// ---- construction ------------------------------------------------------------------------

__global__ void k_ell_structure(label n, const label *__restrict__ row_ptrs, const label *__restrict__ cols,
                                int width, int64_t pitch, label *__restrict__ ell_cols)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    const label rs = row_ptrs[row], len = row_ptrs[row + 1] - rs;
    for (int j = 0; j < width; ++j) ell_cols[j * pitch + row] = j < len ? cols[rs + j] : -1;
}

__global__ void k_ell_values(label n, const label *__restrict__ row_ptrs, const double *__restrict__ vals,
                             int width, int64_t pitch, double *__restrict__ ell_vals)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    const label rs = row_ptrs[row], len = row_ptrs[row + 1] - rs;
    for (int j = 0; j < width; ++j) ell_vals[j * pitch + row] = j < len ? vals[rs + j] : 0.0;
}

// (column - row) tuple of a row; false when the row references a ghost column (>= n)
__device__ __forceinline__ bool row_tuple(label n, int64_t row, const label *__restrict__ ell_cols, int width,
                                          int64_t pitch, label (&d)[kPatMaxW], unsigned long long &h)
{
    bool local = true;
    h = 0xcbf29ce484222325ull;
#pragma unroll
    for (int u = 0; u < kPatMaxW; ++u) {
        label c = -1;
        if (u < width) c = ell_cols[u * pitch + row];
        d[u] = c < 0 ? kPadDelta : c - (label)row;
        if (c >= n) local = false;
        h = (h ^ (unsigned long long)(unsigned int)d[u]) * 0x100000001b3ull;
        h ^= h >> 29;
    }
    if (h == 0) h = 1;   // 0 marks an empty slot
    return local;
}

// pass 1: every distinct tuple takes a slot (the CAS winner stores the tuple)
__global__ void k_pat_insert(label n, const label *__restrict__ ell_cols, int width, int64_t pitch,
                             unsigned long long *keys, label *tuples, int *overflow)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    label d[kPatMaxW];
    unsigned long long h;
    if (!row_tuple(n, row, ell_cols, width, pitch, d, h)) return;
    unsigned int slot = (unsigned int)(h % kPatSlots);
    for (int probe = 0; probe < kPatSlots; ++probe, slot = (slot + 1) % kPatSlots) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&keys[slot]);
        if (cur == 0) cur = atomicCAS(&keys[slot], 0ull, h);
        if (cur == 0) {   // mine
#pragma unroll
            for (int u = 0; u < kPatMaxW; ++u) tuples[slot * kPatMaxW + u] = d[u];
            return;
        }
        if (cur == h) return;
    }
    *overflow = 1;
}

// pass 2 (one thread): codes 0..254 for the occupied slots in slot order, compact table
__global__ void k_pat_ids(const unsigned long long *keys, const label *tuples, int width, int *ids,
                          label *ptab, int *n_patterns)
{
    int count = 0;
    for (int s = 0; s < kPatSlots; ++s) {
        if (keys[s] == 0) {
            ids[s] = kEscape;
            continue;
        }
        if (count < kEscape) {
            ids[s] = count;
            for (int u = 0; u < width; ++u) ptab[count * width + u] = tuples[s * kPatMaxW + u];
            ++count;
        } else {
            ids[s] = kEscape;
        }
    }
    *n_patterns = count;
}

// pass 3: the row's code; the tuple is compared in full (two tuples may share a hash)
__global__ void k_pat_assign(label n, const label *__restrict__ ell_cols, int width, int64_t pitch,
                             const unsigned long long *__restrict__ keys, const label *__restrict__ tuples,
                             const int *__restrict__ ids, unsigned char *__restrict__ code,
                             unsigned long long *n_escape)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int cd = kEscape;
    if (row < n) {
        label d[kPatMaxW];
        unsigned long long h;
        if (row_tuple(n, row, ell_cols, width, pitch, d, h)) {
            unsigned int slot = (unsigned int)(h % kPatSlots);
            for (int probe = 0; probe < kPatSlots; ++probe, slot = (slot + 1) % kPatSlots) {
                const unsigned long long cur = keys[slot];
                if (cur == 0) break;
                if (cur == h) {
                    bool same = true;
#pragma unroll
                    for (int u = 0; u < kPatMaxW; ++u) same = same && tuples[slot * kPatMaxW + u] == d[u];
                    if (same) cd = ids[slot];
                    break;
                }
            }
        }
        code[row] = (unsigned char)cd;
    }
    const unsigned int esc = __ballot_sync(0xffffffffu, row < n && cd == kEscape);
    if ((threadIdx.x & 31) == 0 && esc) atomicAdd(n_escape, (unsigned long long)__popc(esc));
}

// ---- kernels -------------------------------------------------------------------------------

__device__ __forceinline__ unsigned int ld_code(const unsigned char *p)
{
    unsigned int r;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// columns of a row: from the pattern table (PC, code != escape) or from the column array
template <int W, bool PC>
__device__ __forceinline__ void ell_columns(label (&c)[W], int64_t row, unsigned int cd, const EllK &m,
                                            const label *tab, unsigned long long pol)
{
    if (PC && cd != (unsigned int)kEscape) {
#pragma unroll
        for (int u = 0; u < W; ++u) {
            const label d = tab[cd * W + u];
            c[u] = d == kPadDelta ? -1 : (label)row + d;
        }
    } else {
#pragma unroll
        for (int u = 0; u < W; ++u) c[u] = ld_mat(&m.cols[u * m.pitch + row], pol);
    }
}

// W > 0: compile-time row width -- all loads of a row are issued before the first dependent
// instruction (with a run-time slot loop ptxas interleaves DMUL/DADD with the loads of the fused
// instantiation: two extra memory round trips per row for an in-order warp).  W == 0: any width.
template <bool ADV, int NRED, int W, bool PC>
__global__ void __launch_bounds__(kEllThreads, kEllCtasPerSM)
k_spmv_ell(const SpmvK a, const EllK m)
{
    __shared__ label tab[PC ? 256 * (W > 0 ? W : 1) : 1];
    if (a.guard_done && a.state->done) return;
    if (PC) {
        for (int i = threadIdx.x; i < m.n_patterns * W; i += kEllThreads) tab[i] = m.ptab[i];
        __syncthreads();
    }
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < a.n; row += stride) {
        // everything the row needs is requested up front: a warp issues in order, so a load
        // placed behind the row sum would add its whole latency to every trip
        unsigned int cd = kEscape;
        if (PC) cd = ld_code(m.code + row);
        double dw = 0.0;
        if (NRED >= 1) asm volatile("ld.global.f64 %0, [%1];" : "=d"(dw) : "l"(a.dot_with + row));
        double sum = ADV ? __dmul_rn(a.beta, a.y_in[row]) : 0.0;
        if (W > 0) {
            constexpr int WW = W > 0 ? W : 1;
            label c[WW];
            double v[WW], xv[WW];
#pragma unroll
            for (int u = 0; u < WW; ++u) v[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
            ell_columns<WW, PC>(c, row, cd, m, tab, a.mat_policy);
#pragma unroll
            for (int u = 0; u < WW; ++u) xv[u] = c[u] >= 0 ? __ldg(&a.x[c[u]]) : 0.0;
#pragma unroll
            for (int u = 0; u < WW; ++u)
                if (c[u] >= 0) sum = __dadd_rn(sum, prod_of(v[u], xv[u], a.alpha, ADV));
        } else {
            for (int j0 = 0; j0 < m.width; j0 += kEllBatch) {
                label c[kEllBatch];
                double v[kEllBatch], xv[kEllBatch];
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u)
                    c[u] = j0 + u < m.width ? ld_mat(&m.cols[(j0 + u) * m.pitch + row], a.mat_policy) : -1;
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u)
                    v[u] = j0 + u < m.width ? ld_mat(&m.vals[(j0 + u) * m.pitch + row], a.mat_policy) : 0.0;
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u) xv[u] = c[u] >= 0 ? __ldg(&a.x[c[u]]) : 0.0;
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u)
                    if (c[u] >= 0) sum = __dadd_rn(sum, prod_of(v[u], xv[u], a.alpha, ADV));
            }
        }
        a.y[row] = sum;
        if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
        if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// CG step_1 fused into the SpMV:  x = z (or r), y_in = p (previous), y = q, p_new = the other p
// buffer.  GHOST (several ranks, peer-memory path): columns >= n are ghost operands -- z from the
// stamped words in slot 2 of this rank's window (pushed by the neighbours' k_cg_xr, stamped with
// the number of the all-reduce that kernel ended with), p from / to the ghost part of the p
// buffers, which the one row referencing the column keeps up to date.
template <int W, bool PC, bool GHOST>
__global__ void __launch_bounds__(kEllThreads, kEllCtasPerSM)
k_spmv_ell_cgp(const SpmvK a, const EllK m, double *__restrict__ p_new)
{
    __shared__ label tab[PC ? 256 * W : 1];
    if (a.guard_done && a.state->done) return;
    if (PC) {
        for (int i = threadIdx.x; i < m.n_patterns * W; i += kEllThreads) tab[i] = m.ptab[i];
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(a.ea, 0);
    const bool p_is_z = a.state->flag_p_is_z != 0;
    const double t = a.state->coef_p;
    const unsigned long long *zg = nullptr;
    unsigned long long stamp = 0;
    long long t0 = 0;
    if (GHOST) {
        const CommDev *cm = a.ea.comm;
        zg = reinterpret_cast<const unsigned long long *>(cm->my_recv + 2 * (size_t)cm->my_recv_stride);
        stamp = stamp_of(ld_ar_seq(cm));
        t0 = clock64();
    }
    double red[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < a.n; row += stride) {
        unsigned int cd = kEscape;
        if (PC) cd = ld_code(m.code + row);
        label c[W];
        double v[W], zc[W], pc[W];
#pragma unroll
        for (int u = 0; u < W; ++u) v[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
        ell_columns<W, PC>(c, row, cd, m, tab, a.mat_policy);
        // ghost columns (GHOST: c >= n) sit BEHIND the local ones in a row of the ghosted matrix:
        // the unrolled part handles the local operands, the rare rows with ghost entries append
        // theirs below in slot order -- the same left-to-right sum
        bool has_ghost = false;
#pragma unroll
        for (int u = 0; u < W; ++u) {
            if (GHOST && c[u] >= a.n) {
                has_ghost = true;
                c[u] = -1;
            }
            zc[u] = c[u] >= 0 ? __ldg(&a.x[c[u]]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < W; ++u) pc[u] = (c[u] >= 0 && !p_is_z) ? a.y_in[c[u]] : 0.0;
        double sum = 0.0, mine = 0.0;
#pragma unroll
        for (int u = 0; u < W; ++u) {
            if (c[u] >= 0) {
                const double pv = p_is_z ? zc[u] : __dadd_rn(zc[u], __dmul_rn(t, pc[u]));
                if (c[u] == (label)row) mine = pv;                 // the diagonal slot: my own p'
                sum = __dadd_rn(sum, __dmul_rn(v[u], pv));
            }
        }
        if (GHOST && has_ghost) {
#pragma unroll 1
            for (int u = 0; u < W; ++u) {
                const label cg = m.cols[u * m.pitch + row];
                if (cg < a.n) continue;
                double zgv = 0.0;
                if (!pull_stamped(zg + 2 * (size_t)(cg - a.n), stamp, t0, zgv)) a.state->comm_error = 1;
                const double pv = p_is_z ? zgv : __dadd_rn(zgv, __dmul_rn(t, a.y_in[cg]));
                p_new[cg] = pv;   // ghost entry of p: referenced by this row only
                sum = __dadd_rn(sum, __dmul_rn(m.vals[u * m.pitch + row], pv));
            }
        }
        p_new[row] = mine;
        a.y[row] = sum;
        red[0] = __dadd_rn(red[0], __dmul_rn(mine, sum));
    }
    grid_reduce<1>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
}

}  // namespace

// --------------------------------------------------------------------------------------------

void ell_invalidate(Context *ctx, bool structure)
{
    for (Context::EllMatrix *e : {&ctx->ell, &ctx->gell}) {
        e->ready = false;
        if (structure) {
            e->structure_ready = false;
            e->coded = false;
            e->n_patterns = 0;
            e->n_escape = 0;
        }
    }
}

// find the distinct (column - row) tuples and give every row its code
static int ell_detect_patterns(Context *ctx, Context::EllMatrix &e)
{
    e.coded = false;
    e.n_patterns = 0;
    e.n_escape = 0;
    if (ctx->ell_coded == 0 || e.width > kPatMaxW || ctx->n == 0) return OGL_OK;
    cudaStream_t st = ctx->stream;
    unsigned long long *d_keys = nullptr, *d_esc = nullptr;
    label *d_tuples = nullptr;
    int *d_ids = nullptr, *d_misc = nullptr;
    auto cleanup = [&]() { cudaFree(d_keys), cudaFree(d_esc), cudaFree(d_tuples), cudaFree(d_ids), cudaFree(d_misc); };
    int rc;
    if ((rc = dev_alloc(ctx, &d_keys, kPatSlots)) || (rc = dev_alloc(ctx, &d_esc, 1)) ||
        (rc = dev_alloc(ctx, &d_tuples, kPatSlots * kPatMaxW)) || (rc = dev_alloc(ctx, &d_ids, kPatSlots)) ||
        (rc = dev_alloc(ctx, &d_misc, 2))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_keys, 0, sizeof(unsigned long long) * kPatSlots, st);
    cudaMemsetAsync(d_esc, 0, sizeof(unsigned long long), st);
    cudaMemsetAsync(d_misc, 0, 2 * sizeof(int), st);
    const int grid = (int)(((int64_t)ctx->n + 255) / 256);
    k_pat_insert<<<grid, 256, 0, st>>>(ctx->n, e.cols, e.width, e.pitch, d_keys, d_tuples, d_misc);
    k_pat_ids<<<1, 1, 0, st>>>(d_keys, d_tuples, e.width, d_ids, e.ptab, d_misc + 1);
    k_pat_assign<<<grid, 256, 0, st>>>(ctx->n, e.cols, e.width, e.pitch, d_keys, d_tuples, d_ids, e.code, d_esc);
    ctx->launches += 3;
    int misc[2] = {0, 0};
    unsigned long long esc = 0;
    cudaMemcpyAsync(misc, d_misc, sizeof(misc), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&esc, d_esc, sizeof(esc), cudaMemcpyDeviceToHost, st);
    cudaError_t err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = cudaGetLastError();
    cleanup();
    if (err != cudaSuccess) return fail(ctx, OGL_ERR_CUDA, std::string("ell_detect_patterns: ") + cudaGetErrorString(err));
    e.n_patterns = misc[1];
    e.n_escape = (int64_t)esc;
    // a full table (misc[0]) only means more escapes; the codes that were assigned are exact
    const bool worth = e.n_patterns > 0 && (ctx->ell_coded == 2 || 4 * e.n_escape <= (int64_t)ctx->n);
    e.coded = worth;
    return OGL_OK;
}

// (re)build the ELL copy: structure once per pattern, values once per coefficient update; never
// inside a graph capture
static int ell_prepare(Context *ctx, bool ghosted)
{
    Context::EllMatrix &e = ghosted ? ctx->gell : ctx->ell;
    if (e.structure_ready && e.ready) return OGL_OK;
    if (ctx->capturing) return fail(ctx, OGL_ERR_INVALID, "ELL matrix not built before the graph capture");
    const int width = (int)(ghosted ? ctx->max_row_len_g : ctx->max_row_len);
    const int64_t nnz = ghosted ? ctx->nnz + ctx->n_halo : ctx->nnz;
    if (width < 1 || width > 64 || (int64_t)width * ctx->n > 3 * nnz)
        return fail(ctx, OGL_ERR_UNSUPPORTED, "rows too long or too irregular for the ELL format");
    const int64_t pitch = ((int64_t)ctx->n + 31) / 32 * 32;
    const label *row_ptrs = ghosted ? ctx->d_g_row_ptrs : ctx->d_row_ptrs;
    const int grid = (int)(((int64_t)ctx->n + 255) / 256);
    if (!e.structure_ready) {
        if (e.width != width || e.pitch != pitch || !e.cols) {
            OGL_TRY(dev_alloc(ctx, &e.cols, (size_t)(width * pitch)));
            OGL_TRY(dev_alloc(ctx, &e.vals, (size_t)(width * pitch)));
            OGL_TRY(dev_alloc(ctx, &e.code, (size_t)pitch));
            OGL_TRY(dev_alloc(ctx, &e.ptab, (size_t)256 * (width < kPatMaxW ? kPatMaxW : width)));
            e.width = width;
            e.pitch = pitch;
            invalidate_graph(ctx);   // a captured chunk holds the old addresses
        }
        k_ell_structure<<<grid, 256, 0, ctx->stream>>>(ctx->n, row_ptrs, ghosted ? ctx->d_g_cols : ctx->d_cols,
                                                      width, pitch, e.cols);
        ctx->launches++;
        const bool was_coded = e.coded;
        OGL_TRY(ell_detect_patterns(ctx, e));
        if (was_coded != e.coded) invalidate_graph(ctx);
        e.structure_ready = true;
        e.ready = false;
    }
    if (!e.ready) {
        k_ell_values<<<grid, 256, 0, ctx->stream>>>(ctx->n, row_ptrs, ghosted ? ctx->d_g_vals : ctx->d_vals, width,
                                                   pitch, e.vals);
        ctx->launches++;
        OGL_CUDA(ctx, cudaGetLastError());
        e.ready = true;
    }
    return OGL_OK;
}

static EllK ell_args(const Context::EllMatrix &e)
{
    EllK m;
    m.cols = e.cols;
    m.vals = e.vals;
    m.code = e.code;
    m.ptab = e.ptab;
    m.pitch = e.pitch;
    m.width = e.width;
    m.n_patterns = e.n_patterns;
    return m;
}

static int ell_grid(const Context *ctx)
{
    const int64_t need = ((int64_t)ctx->n + kEllThreads - 1) / kEllThreads;
    const int64_t cap = ctx->stream_ctas > 0 ? ctx->stream_ctas : (int64_t)kNumSM * kEllCtasPerSM;   // persistent
    const int64_t g = need < cap ? need : cap;
    return g < 1 ? 1 : (int)g;
}

int spmv_ell(Context *ctx, SpmvK &k, const SpmvArgs &sa, bool ghosted)
{
    OGL_TRY(ell_prepare(ctx, ghosted));
    const Context::EllMatrix &e = ghosted ? ctx->gell : ctx->ell;
    if (sa.ghost_x) k.ea = make_epi_args(ctx, sa.nred), k.ea.trace_tag = 20;   // all-reduce inside the launch
    const EllK m = ell_args(e);
    const int grid = ell_grid(ctx);
    cudaStream_t st = ctx->stream;
    const int nred = sa.nred;
    const bool pc = e.coded && (e.width == 7 || e.width == 5);
#define ELL_GO(A, R, W, P) k_spmv_ell<A, R, W, P><<<grid, kEllThreads, 0, st>>>(k, m)
#define ELL_W(A, R)                                           \
    do {                                                      \
        if (e.width == 7) {                                   \
            if (pc) ELL_GO(A, R, 7, true);                    \
            else ELL_GO(A, R, 7, false);                      \
        } else if (e.width == 5) {                            \
            if (pc) ELL_GO(A, R, 5, true);                    \
            else ELL_GO(A, R, 5, false);                      \
        } else if (e.width == 8) {                            \
            ELL_GO(A, R, 8, false);                           \
        } else {                                              \
            ELL_GO(A, R, 0, false);                           \
        }                                                     \
    } while (0)
    if (sa.advanced) {
        if (nred == 0) ELL_W(true, 0);
        else if (nred == 1) ELL_W(true, 1);
        else ELL_W(true, 2);
    } else {
        if (nred == 0) ELL_W(false, 0);
        else if (nred == 1) ELL_W(false, 1);
        else ELL_W(false, 2);
    }
#undef ELL_W
#undef ELL_GO
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

// q = A p', p' = z + coef_p p, <p',q>, CG_BETA epilogue -- one launch (see k_spmv_ell_cgp).
// Returns OGL_ERR_UNSUPPORTED when the fused form does not apply (caller falls back).
int spmv_ell_cgp(Context *ctx, const double *z, const double *p_old, double *p_new, double *q, bool ghost)
{
    if (spmv_variant_in_use(ctx) != 7) return OGL_ERR_UNSUPPORTED;
    if (!ghost && ctx->n_ranks != 1) return OGL_ERR_UNSUPPORTED;
    // a rank without halo rows has no ghosted matrix: it runs its local one, all-reduce included
    const bool use_g = ghost && ctx->have_ghosted;
    {
        // the fused kernel exists for the stencil widths only: decide before building anything
        const int width = (int)(use_g ? ctx->max_row_len_g : ctx->max_row_len);
        if (width != 7 && width != 5) return OGL_ERR_UNSUPPORTED;
    }
    OGL_TRY(ell_prepare(ctx, use_g));
    const Context::EllMatrix &e = use_g ? ctx->gell : ctx->ell;
    SpmvK k;
    std::memset(&k, 0, sizeof(k));
    k.x = z;
    k.y_in = p_old;
    k.y = q;
    k.n = ctx->n;
    k.mat_policy = spmv_l2_policy(ctx);
    k.partials = ctx->d_partials;
    k.ticket = ctx->d_ticket;
    k.state = ctx->d_state;
    k.epi = EPI_CG_BETA;
    k.inline_epi = 1;
    k.guard_done = 1;
    k.ea = make_epi_args(ctx, ghost ? 1 : 0);
    k.ea.trace_tag = 20;
    const EllK m = ell_args(e);
    const int grid = ell_grid(ctx);
    cudaStream_t st = ctx->stream;
#define CGP_GO(W, P, G) k_spmv_ell_cgp<W, P, G><<<grid, kEllThreads, 0, st>>>(k, m, p_new)
#define CGP_W(W)                                   \
    do {                                           \
        if (e.coded) {                             \
            if (ghost) CGP_GO(W, true, true);      \
            else CGP_GO(W, true, false);           \
        } else {                                   \
            if (ghost) CGP_GO(W, false, true);     \
            else CGP_GO(W, false, false);          \
        }                                          \
    } while (0)
    if (e.width == 7) CGP_W(7);
    else CGP_W(5);
#undef CGP_W
#undef CGP_GO
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace ogl
