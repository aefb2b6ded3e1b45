// (Block-)Jacobi preconditioner on the LOCAL matrix block.
//
// Replaces Preconditioner::init_preconditioner_impl("BJ") + wrap_schwarz
// (Preconditioner/Preconditioner.H:47-64, :91-105), i.e. Ginkgo
// preconditioner::Jacobi generated on distributed::Matrix::get_local_matrix()
// and wrapped in a non-overlapping Schwarz: no communication in the apply.
//   max_block_size == 1 : jacobi::invert_diagonal + scalar_apply (z = r * 1/a_ii)
//   max_block_size  > 1 : jacobi::find_blocks (natural blocks = consecutive
//       rows with identical column pattern, then greedy agglomeration up to
//       max_block_size), Gauss-Jordan inversion with column pivoting,
//       apply = dense block mat-vec.
#include <vector>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

namespace {

__global__ void k_invert_diagonal(label n, const label *__restrict__ row_ptrs,
                                  const label *__restrict__ cols,
                                  const double *__restrict__ vals, double *inv_diag)
{
    const label row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double d = 0.0;
    for (label q = row_ptrs[row]; q < row_ptrs[row + 1]; ++q) {
        if (cols[q] == row) {
            d = vals[q];
            break;   // a duplicate (row,row) cyclic coupling sorts behind the diagonal
        }
    }
    inv_diag[row] = 1.0 / d;
}

// flag[i] = 1 when rows i and i+1 have the same column pattern
__global__ void k_same_pattern(label n, const label *__restrict__ row_ptrs,
                               const label *__restrict__ cols, unsigned char *flag,
                               int *any)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const label a = row_ptrs[i], b = row_ptrs[i + 1], c = row_ptrs[i + 2];
    unsigned char same = (b - a) == (c - b);
    for (label k = 0; same && k < b - a; ++k) same = cols[a + k] == cols[b + k];
    flag[i] = same;
    if (same) atomicExch(any, 1);
}

// one thread per block: gather the dense diagonal block (row-major) and invert
// it in place with Gauss-Jordan + column-max pivoting
__global__ void k_generate_blocks(label n_blocks, const label *__restrict__ block_ptrs,
                                  const int64_t *__restrict__ block_offs,
                                  const label *__restrict__ row_ptrs,
                                  const label *__restrict__ cols,
                                  const double *__restrict__ vals, double *inv)
{
    const label b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const label lo = block_ptrs[b], hi = block_ptrs[b + 1], sz = hi - lo;
    double *m = inv + block_offs[b];
    for (label i = 0; i < sz * sz; ++i) m[i] = 0.0;
    for (label i = lo; i < hi; ++i)
        for (label q = row_ptrs[i]; q < row_ptrs[i + 1]; ++q) {
            const label c = cols[q];
            if (c >= lo && c < hi) m[(i - lo) * sz + (c - lo)] = vals[q];
        }
    label perm[32];
    for (label i = 0; i < sz; ++i) perm[i] = i;
    for (label k = 0; k < sz; ++k) {
        label p = k;
        double best = fabs(m[k * sz + k]);
        for (label i = k + 1; i < sz; ++i) {
            const double v = fabs(m[i * sz + k]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (p != k) {
            for (label j = 0; j < sz; ++j) {
                const double t = m[k * sz + j];
                m[k * sz + j] = m[p * sz + j];
                m[p * sz + j] = t;
            }
            const label t = perm[k];
            perm[k] = perm[p];
            perm[p] = t;
        }
        const double d = m[k * sz + k];
        for (label i = 0; i < sz; ++i) m[i * sz + k] = m[i * sz + k] / -d;
        m[k * sz + k] = 0.0;
        for (label i = 0; i < sz; ++i) {
            const double f = m[i * sz + k];
            for (label j = 0; j < sz; ++j)
                if (j != k) m[i * sz + j] = __dadd_rn(m[i * sz + j], __dmul_rn(f, m[k * sz + j]));
        }
        for (label j = 0; j < sz; ++j) m[k * sz + j] = m[k * sz + j] / d;
        m[k * sz + k] = 1.0 / d;
    }
    // undo the row permutation on the columns: column perm[k] <- column k
    // (cycle-following would save the scratch; blocks are tiny)
    double tmp[32];
    for (label i = 0; i < sz; ++i) {
        for (label k = 0; k < sz; ++k) tmp[k] = m[i * sz + k];
        for (label k = 0; k < sz; ++k) m[i * sz + perm[k]] = tmp[k];
    }
}

__global__ void k_row_block(label n_blocks, const label *__restrict__ block_ptrs,
                            label *row_block)
{
    const label b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    for (label i = block_ptrs[b]; i < block_ptrs[b + 1]; ++i) row_block[i] = b;
}

// ---- ISAI / GISAI (sparsity power 1) ------------------------------------------------------------
// One thread per row: gather the row's column set J (spd: columns <= row), the dense A(J,J) (or its
// transpose), solve with Gaussian elimination + partial pivoting (k <= 8) and scatter the row of the
// approximate inverse over the CSR pattern (spd: also the transposed position).
constexpr int kIsaiMax = 8;

template <bool SPD>
__global__ void __launch_bounds__(128) k_isai_generate(label n, const label *__restrict__ rp,
                                                       const label *__restrict__ cols,
                                                       const double *__restrict__ vals, double *__restrict__ W,
                                                       double *__restrict__ WT, int *bad)
{
    const label i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    label J[kIsaiMax], pos[kIsaiMax];
    int k = 0, p = -1;
    if (rp[i + 1] - rp[i] > kIsaiMax) {
        *bad = 1;
        return;
    }
    for (label e = rp[i]; e < rp[i + 1]; ++e) {
        const label c = cols[e];
        if (SPD && c > i) continue;
        if (c == i) p = k;
        J[k] = c;
        pos[k] = e;
        ++k;
    }
    if (p < 0) {
        *bad = 2;
        return;
    }
    double M[kIsaiMax * kIsaiMax], y[kIsaiMax];
    for (int a = 0; a < k * kIsaiMax; ++a) M[a] = 0.0;
    for (int a = 0; a < k; ++a) {
        y[a] = a == p ? 1.0 : 0.0;
        const label j = J[a];
        for (label e = rp[j]; e < rp[j + 1]; ++e) {
            const label c = cols[e];
            const double v = vals[e];
            for (int b = 0; b < k; ++b)
                if (c == J[b]) {
                    if (SPD) M[a * kIsaiMax + b] = v;   // A(J,J)
                    else M[b * kIsaiMax + a] = v;       // A(J,J)^T
                }
        }
    }
    for (int c = 0; c < k; ++c) {
        int piv = c;
        double best = fabs(M[c * kIsaiMax + c]);
        for (int r = c + 1; r < k; ++r) {
            const double v = fabs(M[r * kIsaiMax + c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (piv != c) {
            for (int cc = 0; cc < k; ++cc) {
                const double t = M[c * kIsaiMax + cc];
                M[c * kIsaiMax + cc] = M[piv * kIsaiMax + cc];
                M[piv * kIsaiMax + cc] = t;
            }
            const double t = y[c];
            y[c] = y[piv];
            y[piv] = t;
        }
        const double d = M[c * kIsaiMax + c];
        for (int r = c + 1; r < k; ++r) {
            const double f = M[r * kIsaiMax + c] / d;
            for (int cc = c; cc < k; ++cc)
                M[r * kIsaiMax + cc] = __dsub_rn(M[r * kIsaiMax + cc], __dmul_rn(f, M[c * kIsaiMax + cc]));
            y[r] = __dsub_rn(y[r], __dmul_rn(f, y[c]));
        }
    }
    for (int r = k - 1; r >= 0; --r) {
        double s = y[r];
        for (int cc = r + 1; cc < k; ++cc) s = __dsub_rn(s, __dmul_rn(M[r * kIsaiMax + cc], y[cc]));
        y[r] = s / M[r * kIsaiMax + r];
    }
    if (SPD) {
        const double scale = 1.0 / sqrt(y[p]);
        for (int a = 0; a < k; ++a) {
            const double w = __dmul_rn(y[a], scale);
            W[pos[a]] = w;
            const label j = J[a];
            for (label e = rp[j]; e < rp[j + 1]; ++e)
                if (cols[e] == i) WT[e] = w;
        }
    } else {
        for (int a = 0; a < k; ++a) W[pos[a]] = y[a];
    }
}

struct ApplyK {
    label n;
    const double *r;
    double *z;
    const double *inv_diag;
    const label *row_block, *block_ptrs;
    const int64_t *block_offs;
    const double *inv_blocks;
    const double *dot_with;   // fused <dot_with, z>
    double *partials;
    unsigned int *ticket;
    SolveState *state;
    int red_base, epi, inline_epi, guard_done;
    EpiArgs ea;
};

// z = M^-1 r, one thread per row.  KIND 1: scalar, 2: block, 0: z was written by the launches
// before (ILU / IC sweeps) and only <dot_with, z> + the epilogue are left.
template <int KIND, int NRED>
__global__ void __launch_bounds__(256) k_jacobi_apply(const ApplyK a)
{
    if (a.guard_done && a.state->done) return;
    double red[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < a.n; row += stride) {
        double z;
        if (KIND == 0) {
            z = a.z[row];
        } else if (KIND == 1) {
            z = __dmul_rn(a.r[row], a.inv_diag[row]);   // jacobi::scalar_apply
        } else {
            const label b = a.row_block[row];
            const label lo = a.block_ptrs[b], sz = a.block_ptrs[b + 1] - lo;
            const double *m = a.inv_blocks + a.block_offs[b] + (int64_t)(row - lo) * sz;
            z = 0.0;
            for (label j = 0; j < sz; ++j) z = __dadd_rn(z, __dmul_rn(m[j], a.r[lo + j]));
        }
        if (KIND != 0) a.z[row] = z;
        if (NRED) red[0] += __dmul_rn(a.dot_with[row], z);
    }
    if (NRED)
        grid_reduce<1>(red, a.partials, a.ticket, a.state, a.red_base, a.epi,
                       a.inline_epi != 0, a.ea);
}

}  // namespace

int precond_setup(Context *ctx, int kind, label mbs)
{
    if (!ctx->have_pattern || !ctx->have_values)
        return fail(ctx, OGL_ERR_INVALID, "ogl_precond_setup before the matrix is assembled");
    if (kind != OGL_PRECOND_NONE && kind != OGL_PRECOND_BJ && kind != OGL_PRECOND_ISAI && kind != OGL_PRECOND_GISAI &&
        !is_tri_precond(kind) && kind != OGL_PRECOND_MULTIGRID)
        return fail(ctx, OGL_ERR_UNSUPPORTED,
                    "preconditioner not supported; valid choices: none, BJ, ISAI, GISAI, ILU, IC, IRILU, Multigrid");
    if (kind == OGL_PRECOND_MULTIGRID) {
        OGL_TRY(mg_setup(ctx));
        ctx->precond_kind = kind;
        ctx->max_block_size = 1;
        ctx->have_precond = true;
        ctx->precond_setups++;
        return OGL_OK;
    }
    if (is_tri_precond(kind)) {
        if (kind == OGL_PRECOND_IC && !ctx->symmetric)
            return fail(ctx, OGL_ERR_INVALID, "IC needs a symmetric matrix; use ILU");
        OGL_TRY(tri_setup(ctx, kind));
        ctx->precond_kind = kind;
        ctx->max_block_size = 1;
        ctx->have_precond = true;
        ctx->precond_setups++;
        return OGL_OK;
    }
    if (kind == OGL_PRECOND_ISAI || kind == OGL_PRECOND_GISAI) {
        if (ctx->max_row_len > kIsaiMax)
            return fail(ctx, OGL_ERR_UNSUPPORTED, "ISAI: rows longer than 8 entries are not supported");
        if (kind == OGL_PRECOND_ISAI && !ctx->symmetric)
            return fail(ctx, OGL_ERR_INVALID, "ISAI (spd) needs a symmetric matrix; use GISAI");
        const bool spd = kind == OGL_PRECOND_ISAI;
        if (!ctx->d_isai_w) OGL_TRY(dev_alloc(ctx, &ctx->d_isai_w, (size_t)ctx->nnz));
        if (spd && !ctx->d_isai_wt) OGL_TRY(dev_alloc(ctx, &ctx->d_isai_wt, (size_t)ctx->nnz));
        int *d_bad = nullptr;
        OGL_TRY(dev_alloc(ctx, &d_bad, 1));
        cudaStream_t s2 = ctx->stream;
        cudaMemsetAsync(d_bad, 0, sizeof(int), s2);
        cudaMemsetAsync(ctx->d_isai_w, 0, sizeof(double) * ctx->nnz, s2);
        if (spd) cudaMemsetAsync(ctx->d_isai_wt, 0, sizeof(double) * ctx->nnz, s2);
        const int grid = (ctx->n + 127) / 128;
        if (ctx->n > 0) {
            if (spd)
                k_isai_generate<true><<<grid, 128, 0, s2>>>(ctx->n, ctx->d_row_ptrs, ctx->d_cols, ctx->d_vals,
                                                            ctx->d_isai_w, ctx->d_isai_wt, d_bad);
            else
                k_isai_generate<false><<<grid, 128, 0, s2>>>(ctx->n, ctx->d_row_ptrs, ctx->d_cols, ctx->d_vals,
                                                             ctx->d_isai_w, ctx->d_isai_wt, d_bad);
            ctx->launches++;
        }
        int bad = 0;
        cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s2);
        cudaError_t e = cudaStreamSynchronize(s2);
        cudaFree(d_bad);
        if (e != cudaSuccess) return fail(ctx, OGL_ERR_CUDA, std::string("ISAI generation: ") + cudaGetErrorString(e));
        if (bad) return fail(ctx, OGL_ERR_INVALID, "ISAI: a row without diagonal entry");
        ctx->precond_kind = kind;
        ctx->max_block_size = 1;
        ctx->have_precond = true;
        ctx->precond_setups++;
        return OGL_OK;
    }
    if (mbs < 1) mbs = 1;
    if (mbs > 32) return fail(ctx, OGL_ERR_INVALID, "maxBlockSize must be in [1, 32]");
    ctx->precond_kind = kind;
    ctx->max_block_size = mbs;
    ctx->have_precond = true;
    ctx->precond_setups++;
    if (kind == OGL_PRECOND_NONE || ctx->n == 0) return OGL_OK;
    cudaStream_t st = ctx->stream;
    const label n = ctx->n;
    if (mbs == 1) {
        if (!ctx->d_inv_diag) OGL_TRY(dev_alloc(ctx, &ctx->d_inv_diag, n));
        k_invert_diagonal<<<(n + 255) / 256, 256, 0, st>>>(n, ctx->d_row_ptrs, ctx->d_cols,
                                                         ctx->d_vals, ctx->d_inv_diag);
        ctx->launches++;
        OGL_CUDA(ctx, cudaGetLastError());
        return OGL_OK;
    }
    // --- block pointers (structure only; cached while the pattern lives) ---
    if (!ctx->d_block_ptrs || ctx->n_blocks == 0 || ctx->bj_pattern_mbs != mbs) {
        unsigned char *d_flag = nullptr;
        int *d_any = nullptr;
        OGL_TRY(dev_alloc(ctx, &d_flag, n));
        OGL_TRY(dev_alloc(ctx, &d_any, 1));
        cudaMemsetAsync(d_flag, 0, n, st);
        cudaMemsetAsync(d_any, 0, sizeof(int), st);
        k_same_pattern<<<(n + 255) / 256, 256, 0, st>>>(n, ctx->d_row_ptrs, ctx->d_cols, d_flag,
                                                      d_any);
        int any = 0;
        cudaMemcpyAsync(&any, d_any, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        std::vector<unsigned char> flag;
        if (any) {
            flag.resize(n);
            cudaMemcpy(flag.data(), d_flag, n, cudaMemcpyDeviceToHost);
        }
        cudaFree(d_flag);
        cudaFree(d_any);
        // Ginkgo find_natural_blocks + agglomerate_supervariables (sequential by nature)
        std::vector<label> nat;
        nat.reserve(n + 1);
        nat.push_back(0);
        label cur = 1;
        for (label i = 0; i + 1 < n; ++i) {
            if (any && flag[i] && cur < mbs) {
                ++cur;
            } else {
                nat.push_back(nat.back() + cur);
                cur = 1;
            }
        }
        nat.push_back(nat.back() + cur);
        std::vector<label> bp;
        bp.reserve(n / mbs + 2);
        bp.push_back(0);
        cur = nat[1] - nat[0];
        for (size_t i = 1; i + 1 < nat.size(); ++i) {
            const label bs = nat[i + 1] - nat[i];
            if (cur + bs <= mbs) {
                cur += bs;
            } else {
                bp.push_back(bp.back() + cur);
                cur = bs;
            }
        }
        bp.push_back(bp.back() + cur);
        const label nb = (label)bp.size() - 1;
        std::vector<int64_t> offs(nb + 1, 0);
        for (label b = 0; b < nb; ++b) {
            const int64_t sz = bp[b + 1] - bp[b];
            offs[b + 1] = offs[b] + sz * sz;
        }
        bool uniform = true;
        for (label b = 0; b + 1 < nb && uniform; ++b) uniform = (bp[b + 1] - bp[b]) == mbs;
        ctx->bj_uniform = uniform && nb > 0 && (bp[nb] - bp[nb - 1]) <= mbs;
        ctx->n_blocks = nb;
        ctx->inv_blocks_len = offs[nb];
        ctx->bj_pattern_mbs = mbs;
        OGL_TRY(dev_alloc(ctx, &ctx->d_block_ptrs, (size_t)nb + 1));
        OGL_TRY(dev_alloc(ctx, &ctx->d_block_offs, (size_t)nb + 1));
        OGL_TRY(dev_alloc(ctx, &ctx->d_row_block, n));
        OGL_TRY(dev_alloc(ctx, &ctx->d_inv_blocks, (size_t)offs[nb]));
        OGL_CUDA(ctx, cudaMemcpy(ctx->d_block_ptrs, bp.data(), sizeof(label) * (nb + 1),
                                 cudaMemcpyHostToDevice));
        OGL_CUDA(ctx, cudaMemcpy(ctx->d_block_offs, offs.data(), sizeof(int64_t) * (nb + 1),
                                 cudaMemcpyHostToDevice));
        k_row_block<<<(nb + 255) / 256, 256, 0, st>>>(nb, ctx->d_block_ptrs, ctx->d_row_block);
        ctx->launches++;
    }
    k_generate_blocks<<<(ctx->n_blocks + 127) / 128, 128, 0, st>>>(
        ctx->n_blocks, ctx->d_block_ptrs, ctx->d_block_offs, ctx->d_row_ptrs, ctx->d_cols,
        ctx->d_vals, ctx->d_inv_blocks);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

int precond_apply(Context *ctx, const double *r, double *z, const double *dot_with,
                  int red_base, bool guard_done, int epi, bool inline_epi, int ar_count)
{
    if (ctx->precond_kind == OGL_PRECOND_NONE || ctx->n == 0) {
        if (r != z)
            OGL_CUDA(ctx, cudaMemcpyAsync(z, r, sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice,
                                          ctx->stream));
        return OGL_OK;
    }
    if (ctx->precond_kind == OGL_PRECOND_ISAI || ctx->precond_kind == OGL_PRECOND_GISAI) {
        // Schwarz: the LOCAL approximate inverse, as sparse mat-vecs over the pattern of A
        const bool spd = ctx->precond_kind == OGL_PRECOND_ISAI;
        const double *in = r;
        if (spd) {
            double *t;
            OGL_TRY(get_work(ctx, 11, &t));
            SpmvArgs s1;
            s1.x = r;
            s1.y = t;
            s1.vals_override = ctx->d_isai_w;
            s1.guard_done = guard_done;
            OGL_TRY(spmv_local(ctx, s1));
            in = t;
        }
        SpmvArgs s2;
        s2.x = in;
        s2.y = z;
        s2.vals_override = spd ? ctx->d_isai_wt : ctx->d_isai_w;
        s2.guard_done = guard_done;
        if (dot_with) {
            if (red_base != 0) return fail(ctx, OGL_ERR_INVALID, "ISAI apply reduces into red[0] only");
            s2.dot_with = dot_with;
            s2.nred = 1;
            s2.epi = epi;
            s2.inline_epi = inline_epi;
            s2.ar_count = ar_count;
        }
        return spmv_local(ctx, s2);
    }
    const bool tri = is_tri_precond(ctx->precond_kind) || ctx->precond_kind == OGL_PRECOND_MULTIGRID;
    if (tri) {
        // Schwarz: the LOCAL factors / hierarchy; two triangular sweeps (or 5 + 5 Richardson sweeps) or
        // one V cycle, then the reduction the caller asked for as one more pass over z
        if (ctx->precond_kind == OGL_PRECOND_MULTIGRID) OGL_TRY(mg_apply(ctx, r, z, guard_done));
        else OGL_TRY(tri_apply(ctx, r, z, guard_done));
        if (!dot_with) return OGL_OK;
    }
    ApplyK a;
    a.n = ctx->n;
    a.r = r;
    a.z = z;
    a.inv_diag = ctx->d_inv_diag;
    a.row_block = ctx->d_row_block;
    a.block_ptrs = ctx->d_block_ptrs;
    a.block_offs = ctx->d_block_offs;
    a.inv_blocks = ctx->d_inv_blocks;
    a.dot_with = dot_with;
    a.partials = ctx->d_partials;
    a.ticket = ctx->d_ticket;
    a.state = ctx->d_state;
    a.red_base = red_base;
    a.epi = epi;
    a.inline_epi = inline_epi ? 1 : 0;
    a.guard_done = guard_done ? 1 : 0;
    a.ea = make_epi_args(ctx, dot_with ? ar_count : 0);
    int grid = (ctx->n + 255) / 256;
    if (grid > ctx->blas1_blocks) grid = (int)ctx->blas1_blocks;
    const bool scalar = ctx->max_block_size == 1;
    if (tri) {
        k_jacobi_apply<0, 1><<<grid, 256, 0, ctx->stream>>>(a);
    } else if (dot_with) {
        if (scalar) k_jacobi_apply<1, 1><<<grid, 256, 0, ctx->stream>>>(a);
        else k_jacobi_apply<2, 1><<<grid, 256, 0, ctx->stream>>>(a);
    } else {
        if (scalar) k_jacobi_apply<1, 0><<<grid, 256, 0, ctx->stream>>>(a);
        else k_jacobi_apply<2, 0><<<grid, 256, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace ogl
