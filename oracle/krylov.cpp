// TEST INFRASTRUCTURE -- parity oracle, see oracle.h.  Not product code.
//
// PARITY UNPINNED: the arithmetic restated here lives in the third-party
// dependency Ginkgo (ginkgo-project/ginkgo @ fc86d48b78cebd2b2c5833a2dcf0fe40f615cf19,
// third_party/ginkgo/CMakeLists.txt:6-24), which is neither vendored under
// /root/reference nor installable here.  What follows restates its published
// `reference`-executor algorithms (core/solver/{cg,bicgstab,gmres}.cpp,
// reference/solver/*_kernels.cpp, reference/preconditioner/jacobi_kernels.cpp,
// core/distributed/{matrix,vector}.cpp) at the call sites OGL uses:
//   lduLduBase/lduLduBase.H:268-276      generate + apply
//   Solver/CG/GKOCG.H:45-61              Cg, BiCGStab/GMRES analogues
//   Preconditioner/Preconditioner.H:47-64,91-105   Schwarz(Jacobi(local))
// and OGL's own stopping criterion, which IS in the tree and is followed line
// by line:
//   StoppingCriterion/StoppingCriterion.C:11-151
//
// Reference-executor order: every reduction is a plain left-to-right sum, a
// row of the SpMV sums its entries in storage order (row-major: lower columns
// ascending, diagonal, upper columns ascending), the non-local block is added
// afterwards (`y += A_nl * recv`), a distributed reduction sums the ranks'
// partial results in rank order.  With params->threads > 1 the same algorithms
// run with OpenMP (the analogue of Ginkgo's omp executor; used only as the CPU
// baseline in bench.py).
#include "oracle.h"
#include "multigrid.hpp"
#include "trifactor.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using vec = std::vector<orc_scalar>;
using dvec = std::vector<vec>;  // one local vector per rank

constexpr orc_scalar SMALL = 1e-15;  // OpenFOAM DP `SMALL`

struct Sys {
    int R;
    const orc_rank_system *rk;
    std::vector<std::vector<orc_label>> row_ptrs;  // CSR view of the sorted COO
    // recv_src[r][t] = index of r inside the target list of rank target_ids[t]
    std::vector<std::vector<orc_label>> peer_slot;
    std::vector<std::vector<orc_label>> send_off;  // per-rank offsets into send_idxs
    int threads;
};

Sys make_sys(int R, const orc_rank_system *rk, int threads)
{
    Sys s{R, rk, {}, {}, {}, threads};
    s.row_ptrs.resize(R);
    s.peer_slot.resize(R);
    s.send_off.resize(R);
    for (int r = 0; r < R; ++r) {
        auto &rp = s.row_ptrs[r];
        rp.assign(static_cast<size_t>(rk[r].n) + 1, 0);
        for (orc_label k = 0; k < rk[r].nnz; ++k) rp[rk[r].rows[k] + 1]++;
        for (orc_label i = 0; i < rk[r].n; ++i) rp[i + 1] += rp[i];
        auto &so = s.send_off[r];
        so.assign(static_cast<size_t>(rk[r].n_targets) + 1, 0);
        for (orc_label t = 0; t < rk[r].n_targets; ++t)
            so[t + 1] = so[t] + rk[r].target_sizes[t];
    }
    for (int r = 0; r < R; ++r) {
        s.peer_slot[r].assign(static_cast<size_t>(rk[r].n_targets), -1);
        for (orc_label t = 0; t < rk[r].n_targets; ++t) {
            const int q = rk[r].target_ids[t];
            for (orc_label u = 0; u < rk[q].n_targets; ++u)
                if (rk[q].target_ids[u] == r) s.peer_slot[r][t] = u;
        }
    }
    return s;
}

dvec zeros_like(const Sys &s)
{
    dvec v(s.R);
    for (int r = 0; r < s.R; ++r) v[r].assign(static_cast<size_t>(s.rk[r].n), 0.0);
    return v;
}

// Halo exchange: the recv buffer of rank r is blocked by ascending neighbour
// rank (Partition.H:66-67 build_from_blocked_recv); block t holds what rank
// target_ids[t] gathers through ITS send list towards r (HostMatrix.C:262-303).
void exchange(const Sys &s, const dvec &x, dvec &recv)
{
    for (int r = 0; r < s.R; ++r) {
        recv[r].resize(static_cast<size_t>(s.rk[r].n_halo));
        orc_label w = 0;
        for (orc_label t = 0; t < s.rk[r].n_targets; ++t) {
            const int q = s.rk[r].target_ids[t];
            const orc_label u = s.peer_slot[r][t];
            const orc_label o = s.send_off[q][u];
            const orc_label cnt = s.rk[q].target_sizes[u];
            for (orc_label k = 0; k < cnt; ++k)
                recv[r][w++] = x[q][s.rk[q].send_idxs[o + k]];
        }
    }
}

// y = alpha*A*x + beta*y (advanced) or y = A*x.  Reference COO/CSR kernels:
// row sum starts from beta*y (advanced) or 0, adds alpha*val*x[col] entry by
// entry; then the non-local block adds alpha*val*recv[col] onto the stored y.
void spmv(const Sys &s, const dvec &x, dvec &y, bool advanced = false,
          orc_scalar alpha = 1.0, orc_scalar beta = 0.0)
{
    dvec recv(s.R);
    exchange(s, x, recv);
    for (int r = 0; r < s.R; ++r) {
        const auto &k = s.rk[r];
        const auto &rp = s.row_ptrs[r];
        const orc_scalar *xv = x[r].data();
        orc_scalar *yv = y[r].data();
#pragma omp parallel for schedule(static) num_threads(s.threads) if (s.threads > 1)
        for (orc_label i = 0; i < k.n; ++i) {
            orc_scalar sum = advanced ? beta * yv[i] : 0.0;
            if (advanced) {
                for (orc_label e = rp[i]; e < rp[i + 1]; ++e)
                    sum += alpha * k.vals[e] * xv[k.cols[e]];
            } else {
                for (orc_label e = rp[i]; e < rp[i + 1]; ++e)
                    sum += k.vals[e] * xv[k.cols[e]];
            }
            yv[i] = sum;
        }
        const orc_scalar a = advanced ? alpha : 1.0;
        for (orc_label e = 0; e < k.n_halo; ++e)
            yv[k.nl_rows[e]] += a * k.nl_vals[e] * recv[r][k.nl_cols[e]];
    }
}

orc_scalar dot(const Sys &s, const dvec &a, const dvec &b)
{
    orc_scalar total = 0.0;
    for (int r = 0; r < s.R; ++r) {
        orc_scalar part = 0.0;
        const orc_label n = s.rk[r].n;
        const orc_scalar *av = a[r].data(), *bv = b[r].data();
        if (s.threads > 1) {
#pragma omp parallel for reduction(+ : part) schedule(static) num_threads(s.threads)
            for (orc_label i = 0; i < n; ++i) part += av[i] * bv[i];
        } else {
            for (orc_label i = 0; i < n; ++i) part += av[i] * bv[i];
        }
        total += part;  // MPI_Allreduce(SUM), rank order
    }
    return total;
}

orc_scalar norm1(const Sys &s, const dvec &a)
{
    orc_scalar total = 0.0;
    for (int r = 0; r < s.R; ++r) {
        orc_scalar part = 0.0;
        const orc_label n = s.rk[r].n;
        const orc_scalar *av = a[r].data();
        if (s.threads > 1) {
#pragma omp parallel for reduction(+ : part) schedule(static) num_threads(s.threads)
            for (orc_label i = 0; i < n; ++i) part += std::fabs(av[i]);
        } else {
            for (orc_label i = 0; i < n; ++i) part += std::fabs(av[i]);
        }
        total += part;
    }
    return total;
}

orc_scalar norm2(const Sys &s, const dvec &a) { return std::sqrt(dot(s, a, a)); }

// distributed::Vector::compute_mean: local mean scaled by local/global size,
// summed over ranks (StoppingCriterion.C:19).
orc_scalar mean(const Sys &s, const dvec &a)
{
    orc_scalar global = 0.0;
    for (int r = 0; r < s.R; ++r) global += s.rk[r].n;
    orc_scalar total = 0.0;
    for (int r = 0; r < s.R; ++r) {
        const orc_label n = s.rk[r].n;
        if (n == 0) continue;
        orc_scalar sum = 0.0;
        for (orc_label i = 0; i < n; ++i) sum += a[r][i];
        const orc_scalar inv = 1.0 / static_cast<orc_scalar>(n);
        total += (sum * inv) * (static_cast<orc_scalar>(n) / global);
    }
    return total;
}

// ---- block Jacobi --------------------------------------------------------

bool same_pattern(const orc_label *rp, const orc_label *cols, orc_label a,
                  orc_label b)
{
    const orc_label la = rp[a + 1] - rp[a], lb = rp[b + 1] - rp[b];
    if (la != lb) return false;
    for (orc_label k = 0; k < la; ++k)
        if (cols[rp[a] + k] != cols[rp[b] + k]) return false;
    return true;
}

orc_label find_blocks(orc_label n, const orc_label *rp, const orc_label *cols,
                      orc_label mbs, orc_label *bp)
{
    if (n == 0) {
        bp[0] = 0;
        return 0;
    }
    // natural blocks: consecutive rows with identical column pattern
    std::vector<orc_label> nat(static_cast<size_t>(n) + 1);
    nat[0] = 0;
    orc_label nn = 1, cur = 1;
    for (orc_label i = 0; i + 1 < n; ++i) {
        if (same_pattern(rp, cols, i, i + 1) && cur < mbs) {
            ++cur;
        } else {
            nat[nn] = nat[nn - 1] + cur;
            ++nn;
            cur = 1;
        }
    }
    nat[nn] = nat[nn - 1] + cur;
    // greedy agglomeration of adjacent natural blocks while the sum fits
    bp[0] = 0;
    orc_label nb = 1;
    cur = nat[1] - nat[0];
    for (orc_label i = 1; i < nn; ++i) {
        const orc_label bs = nat[i + 1] - nat[i];
        if (cur + bs <= mbs) {
            cur += bs;
        } else {
            bp[nb] = bp[nb - 1] + cur;
            ++nb;
            cur = bs;
        }
    }
    bp[nb] = bp[nb - 1] + cur;
    return nb;
}

// Gauss-Jordan with partial (column-max) pivoting on a row-major b x b block,
// in place; the row permutation is undone on the columns of the result.
void invert_block(orc_label b, orc_scalar *m)
{
    std::vector<orc_label> perm(static_cast<size_t>(b));
    for (orc_label i = 0; i < b; ++i) perm[i] = i;
    for (orc_label k = 0; k < b; ++k) {
        orc_label p = k;
        orc_scalar best = std::fabs(m[k * b + k]);
        for (orc_label i = k + 1; i < b; ++i) {
            if (std::fabs(m[i * b + k]) > best) {
                best = std::fabs(m[i * b + k]);
                p = i;
            }
        }
        if (p != k) {
            for (orc_label j = 0; j < b; ++j) std::swap(m[k * b + j], m[p * b + j]);
            std::swap(perm[k], perm[p]);
        }
        const orc_scalar d = m[k * b + k];
        for (orc_label i = 0; i < b; ++i) m[i * b + k] /= -d;
        m[k * b + k] = 0.0;
        for (orc_label i = 0; i < b; ++i)
            for (orc_label j = 0; j < b; ++j)
                if (j != k) m[i * b + j] += m[i * b + k] * m[k * b + j];
        for (orc_label j = 0; j < b; ++j) m[k * b + j] /= d;
        m[k * b + k] = 1.0 / d;
    }
    // inv(P A) = inv(A) P^T  ->  column perm[k] of inv(A) is column k of m
    std::vector<orc_scalar> tmp(m, m + static_cast<size_t>(b) * b);
    for (orc_label i = 0; i < b; ++i)
        for (orc_label k = 0; k < b; ++k) m[i * b + perm[k]] = tmp[i * b + k];
}

// ---- incomplete sparse approximate inverse (ISAI) ---------------------------------------------
// [upstream Ginkgo core/preconditioner/isai.cpp + reference/preconditioner/isai_kernels.cpp;
// Anzt, Huckle, Braeckle, Dongarra, "Incomplete Sparse Approximate Inverses for Parallel
// Preconditioning", 2018], sparsity power 1, on the LOCAL block (Schwarz).
//   general (GISAI): M with the pattern of A, row i: M(i,J) A(J,J) = I(i,J), J = columns of row i
//                    -> one small dense solve A(J,J)^T m = e_pos(i) per row;  z = M r
//   spd (ISAI):      W with the pattern of tril(A), row i: A(J,J) y = e_last, J = columns <= i,
//                    w = y / sqrt(y_last)  (factorised sparse approximate inverse of the Cholesky
//                    factor);  z = W^T (W r)
// Dense solves: Gaussian elimination with partial pivoting.
constexpr int kIsaiMax = 8;

// solve M y = rhs in place (k x k row-major, k <= kIsaiMax); the solution overwrites rhs
void isai_dense_solve(int k, orc_scalar *M, orc_scalar *rhs)
{
    for (int c = 0; c < k; ++c) {
        int piv = c;
        orc_scalar best = std::fabs(M[c * kIsaiMax + c]);
        for (int r = c + 1; r < k; ++r) {
            const orc_scalar v = std::fabs(M[r * kIsaiMax + c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (piv != c) {
            for (int cc = 0; cc < k; ++cc) std::swap(M[c * kIsaiMax + cc], M[piv * kIsaiMax + cc]);
            std::swap(rhs[c], rhs[piv]);
        }
        const orc_scalar d = M[c * kIsaiMax + c];
        for (int r = c + 1; r < k; ++r) {
            const orc_scalar f = M[r * kIsaiMax + c] / d;
            for (int cc = c; cc < k; ++cc) M[r * kIsaiMax + cc] -= f * M[c * kIsaiMax + cc];
            rhs[r] -= f * rhs[c];
        }
    }
    for (int r = k - 1; r >= 0; --r) {
        orc_scalar s = rhs[r];
        for (int cc = r + 1; cc < k; ++cc) s -= M[r * kIsaiMax + cc] * rhs[cc];
        rhs[r] = s / M[r * kIsaiMax + r];
    }
}

// W (and, spd, its transpose WT) as value arrays over the CSR pattern of A; false: unsupported row
bool isai_generate(orc_label n, const orc_label *rp, const orc_label *cols, const orc_scalar *vals, bool spd,
                   orc_scalar *W, orc_scalar *WT)
{
    for (orc_label e = 0; e < rp[n]; ++e) W[e] = 0.0;
    if (spd)
        for (orc_label e = 0; e < rp[n]; ++e) WT[e] = 0.0;
    for (orc_label i = 0; i < n; ++i) {
        orc_label J[kIsaiMax], pos[kIsaiMax];
        int k = 0, p = -1;
        if (rp[i + 1] - rp[i] > kIsaiMax) return false;
        for (orc_label e = rp[i]; e < rp[i + 1]; ++e) {
            const orc_label c = cols[e];
            if (c >= n || (spd && c > i)) continue;
            if (c == i) p = k;
            J[k] = c;
            pos[k] = e;
            ++k;
        }
        if (p < 0) return false;
        orc_scalar M[kIsaiMax * kIsaiMax], y[kIsaiMax];
        for (int a = 0; a < k * kIsaiMax; ++a) M[a] = 0.0;
        for (int a = 0; a < k; ++a) {
            y[a] = a == p ? 1.0 : 0.0;
            const orc_label j = J[a];
            for (orc_label e = rp[j]; e < rp[j + 1]; ++e)
                for (int b = 0; b < k; ++b)
                    if (cols[e] == J[b]) {
                        if (spd) M[a * kIsaiMax + b] = vals[e];   // A(J,J)
                        else M[b * kIsaiMax + a] = vals[e];       // A(J,J)^T
                    }
        }
        isai_dense_solve(k, M, y);
        if (spd) {
            const orc_scalar scale = 1.0 / std::sqrt(y[p]);
            for (int a = 0; a < k; ++a) {
                const orc_scalar w = y[a] * scale;
                W[pos[a]] = w;
                const orc_label j = J[a];
                for (orc_label e = rp[j]; e < rp[j + 1]; ++e)
                    if (cols[e] == i) WT[e] = w;
            }
        } else {
            for (int a = 0; a < k; ++a) W[pos[a]] = y[a];
        }
    }
    return true;
}

struct Jacobi {
    int kind = ORC_PRECOND_NONE;
    orc_label mbs = 1;
    dvec isai_w, isai_wt;                        // ISAI / GISAI values over the pattern of A
    dvec fact;                                   // ILU / IC / IRILU: factors over the pattern of A (trifactor.hpp)
    std::vector<std::vector<orc_label>> dpos;    // ... and the position of every row's diagonal entry
    std::vector<orc_mg::Hierarchy> mg;           // Multigrid: one hierarchy per rank (multigrid.hpp)
    dvec inv_diag;                               // mbs == 1
    std::vector<std::vector<orc_label>> bptr;    // mbs > 1
    std::vector<std::vector<orc_label>> boff;
    dvec inv_blocks;
};

orc_mg::Csr local_csr(const Sys &s, int r)
{
    orc_mg::Csr A;
    A.n = s.rk[r].n;
    A.rp = s.row_ptrs[r];
    A.cols.assign(s.rk[r].cols, s.rk[r].cols + s.rk[r].nnz);
    A.vals.assign(s.rk[r].vals, s.rk[r].vals + s.rk[r].nnz);
    return A;
}

Jacobi make_jacobi(const Sys &s, int kind, orc_label mbs, const orc_solve_params *prm = nullptr)
{
    Jacobi J;
    J.kind = kind;
    J.mbs = mbs < 1 ? 1 : mbs;
    if (kind == ORC_PRECOND_MULTIGRID) {
        // Preconditioner.H:261-341 defaults: maxLevels 9, minCoarseRows 10, coarseSolverIters 4
        const int max_levels = prm && prm->mg_max_levels > 0 ? prm->mg_max_levels : 9;
        const orc_label min_rows = prm && prm->mg_min_coarse_rows > 0 ? prm->mg_min_coarse_rows : 10;
        const int coarse_iters = prm && prm->mg_coarse_iters > 0 ? prm->mg_coarse_iters : 4;
        for (int r = 0; r < s.R; ++r) J.mg.push_back(orc_mg::build(local_csr(s, r), max_levels, min_rows, coarse_iters));
        return J;
    }
    if (kind == ORC_PRECOND_ISAI || kind == ORC_PRECOND_GISAI) {
        J.isai_w.resize(s.R);
        J.isai_wt.resize(s.R);
        for (int r = 0; r < s.R; ++r) {
            const auto &k = s.rk[r];
            J.isai_w[r].assign(static_cast<size_t>(k.nnz), 0.0);
            J.isai_wt[r].assign(static_cast<size_t>(k.nnz), 0.0);
            if (!isai_generate(k.n, s.row_ptrs[r].data(), k.cols, k.vals, kind == ORC_PRECOND_ISAI,
                               J.isai_w[r].data(), J.isai_wt[r].data()))
                J.kind = -1;   // reported by orc_solve
        }
        return J;
    }
    if (kind == ORC_PRECOND_ILU || kind == ORC_PRECOND_IC || kind == ORC_PRECOND_IRILU) {
        // factorisation of the LOCAL block (Schwarz, Preconditioner.H:66-82, 113-123)
        J.fact.resize(s.R);
        J.dpos.resize(s.R);
        for (int r = 0; r < s.R; ++r) {
            const auto &k = s.rk[r];
            const orc_label *rp = s.row_ptrs[r].data();
            J.fact[r].assign(k.vals, k.vals + k.nnz);
            if (!orc_tri::diag_positions(k.n, rp, k.cols, J.dpos[r])) {
                J.kind = -1;
                continue;
            }
            if (kind == ORC_PRECOND_IC) {
                if (!orc_tri::ic0(k.n, rp, k.cols, J.dpos[r].data(), J.fact[r].data())) J.kind = -1;
            } else {
                orc_tri::ilu0(k.n, rp, k.cols, J.dpos[r].data(), J.fact[r].data());
            }
        }
        return J;
    }
    if (kind != ORC_PRECOND_BJ) return J;
    if (J.mbs == 1) {
        // jacobi::invert_diagonal on the LOCAL block (Schwarz, Preconditioner.H:53-62)
        J.inv_diag = zeros_like(s);
        for (int r = 0; r < s.R; ++r) {
            const auto &k = s.rk[r];
            for (orc_label e = 0; e < k.nnz; ++e)
                if (k.rows[e] == k.cols[e]) J.inv_diag[r][k.rows[e]] = 1.0 / k.vals[e];
        }
        return J;
    }
    J.bptr.resize(s.R);
    J.boff.resize(s.R);
    J.inv_blocks.resize(s.R);
    for (int r = 0; r < s.R; ++r) {
        const auto &k = s.rk[r];
        J.bptr[r].assign(static_cast<size_t>(k.n) + 1, 0);
        const orc_label nb =
            find_blocks(k.n, s.row_ptrs[r].data(), k.cols, J.mbs, J.bptr[r].data());
        J.bptr[r].resize(static_cast<size_t>(nb) + 1);
        J.boff[r].assign(static_cast<size_t>(nb) + 1, 0);
        for (orc_label b = 0; b < nb; ++b) {
            const orc_label sz = J.bptr[r][b + 1] - J.bptr[r][b];
            J.boff[r][b + 1] = J.boff[r][b] + sz * sz;
        }
        J.inv_blocks[r].assign(static_cast<size_t>(J.boff[r][nb]), 0.0);
        orc_bj_invert_blocks(k.n, s.row_ptrs[r].data(), k.cols, k.vals, nb,
                             J.bptr[r].data(), J.inv_blocks[r].data());
    }
    return J;
}

// z = M^-1 r.  none: Ginkgo's default preconditioner is the identity (copy).
void precond_apply(const Sys &s, const Jacobi &J, const dvec &r, dvec &z)
{
    for (int q = 0; q < s.R; ++q) {
        const orc_label n = s.rk[q].n;
        if (J.kind == ORC_PRECOND_ISAI || J.kind == ORC_PRECOND_GISAI) {
            // sequential CSR row sums over the LOCAL pattern with the ISAI values
            const orc_label *rp = s.row_ptrs[q].data(), *cols = s.rk[q].cols;
            auto apply = [&](const orc_scalar *v, const orc_scalar *in, orc_scalar *out) {
#pragma omp parallel for schedule(static) num_threads(s.threads) if (s.threads > 1)
                for (orc_label i = 0; i < n; ++i) {
                    orc_scalar acc = 0.0;
                    for (orc_label e = rp[i]; e < rp[i + 1]; ++e) acc += v[e] * in[cols[e]];
                    out[i] = acc;
                }
            };
            if (J.kind == ORC_PRECOND_GISAI) {
                apply(J.isai_w[q].data(), r[q].data(), z[q].data());
            } else {
                vec t(static_cast<size_t>(n));
                apply(J.isai_w[q].data(), r[q].data(), t.data());
                apply(J.isai_wt[q].data(), t.data(), z[q].data());
            }
        } else if (J.kind == ORC_PRECOND_MULTIGRID) {
            if (n > 0) {
                std::fill(z[q].begin(), z[q].end(), 0.0);
                orc_mg::vcycle(J.mg[q], 0, r[q].data(), z[q].data());
            }
        } else if (J.kind == ORC_PRECOND_ILU || J.kind == ORC_PRECOND_IC || J.kind == ORC_PRECOND_IRILU) {
            const orc_label *rp = s.row_ptrs[q].data(), *cols = s.rk[q].cols, *dp = J.dpos[q].data();
            const orc_scalar *F = J.fact[q].data();
            vec t(static_cast<size_t>(n));
            if (J.kind == ORC_PRECOND_IRILU) {
                // Ilu<Ir, Ir>::apply: the intermediate starts as a copy of b, x as a copy of the
                // intermediate (both inner solvers use their initial guess); 5 sweeps each
                vec scratch;
                t = r[q];
                orc_tri::ir_jacobi(n, rp, cols, dp, F, true, 5, r[q].data(), t.data(), scratch);
                z[q] = t;
                orc_tri::ir_jacobi(n, rp, cols, dp, F, false, 5, t.data(), z[q].data(), scratch);
            } else {
                orc_tri::lower_solve(n, rp, cols, dp, F, J.kind == ORC_PRECOND_ILU, r[q].data(), t.data());
                orc_tri::upper_solve(n, rp, cols, dp, F, t.data(), z[q].data());
            }
        } else if (J.kind != ORC_PRECOND_BJ) {
            std::memcpy(z[q].data(), r[q].data(), sizeof(orc_scalar) * n);
        } else if (J.mbs == 1) {
            // jacobi::scalar_apply: x = b * inv_diag
            const orc_scalar *rv = r[q].data(), *dv = J.inv_diag[q].data();
            orc_scalar *zv = z[q].data();
#pragma omp parallel for schedule(static) num_threads(s.threads) if (s.threads > 1)
            for (orc_label i = 0; i < n; ++i) zv[i] = rv[i] * dv[i];
        } else {
            const orc_label nb = static_cast<orc_label>(J.bptr[q].size()) - 1;
            for (orc_label b = 0; b < nb; ++b) {
                const orc_label lo = J.bptr[q][b], sz = J.bptr[q][b + 1] - lo;
                const orc_scalar *m = J.inv_blocks[q].data() + J.boff[q][b];
                for (orc_label i = 0; i < sz; ++i) z[q][lo + i] = 0.0;
                // apply_block: x[row] += block(row, inner) * b[inner], inner outer loop
                for (orc_label in = 0; in < sz; ++in)
                    for (orc_label i = 0; i < sz; ++i)
                        z[q][lo + i] += m[i * sz + in] * r[q][lo + in];
            }
        }
    }
}

// ---- OGL stopping criterion (StoppingCriterion.C:71-151) ------------------

struct Criterion {
    const Sys &s;
    const orc_solve_params &p;
    const dvec &x0;  // x at generate time == the vector the solver updates
    const dvec &b;
    orc_label iter = 0;
    orc_scalar norm_factor = 1.0;
    orc_scalar init_res = 0.0;
    orc_scalar res = 0.0;
    orc_scalar *history;
    orc_label history_cap;
    orc_label n_history = 0;

    // StoppingCriterion.C:32-69
    orc_scalar compute_norm_factor(const dvec &r) const
    {
        const orc_scalar xavg = mean(s, x0);          // :17-19
        dvec xref = zeros_like(s);
        for (auto &v : xref) std::fill(v.begin(), v.end(), xavg);  // :27
        dvec w = zeros_like(s);
        spmv(s, xref, w);                             // :29
        orc_scalar total = 0.0;
        for (int q = 0; q < s.R; ++q) {
            orc_scalar part = 0.0;
            for (orc_label i = 0; i < s.rk[q].n; ++i) {
                const orc_scalar bs = b[q][i] - 1.0 * w[q][i];   // :54
                const orc_scalar part2 = std::fabs(bs);          // :56
                orc_scalar t = bs - 1.0 * r[q][i];               // :58
                t = std::fabs(t);                                // :59
                t = t + 1.0 * part2;                             // :61
                part += std::fabs(t);                            // :63
            }
            total += part;
        }
        return total + SMALL;                                    // :68
    }

    bool check(const dvec &r)
    {
        if (iter > 0 && iter < p.min_iter) {   // :77-81
            ++iter;
            return false;
        }
        if (iter % p.frequency != 0) {         // :84-87
            ++iter;
            return false;
        }
        orc_scalar rn = norm1(s, r);           // :92-97
        bool stop = false;
        if (iter == 0) {                       // :102-111
            norm_factor = compute_norm_factor(r);
            init_res = rn / norm_factor;
        }
        rn /= norm_factor;                     // :113
        if (history && iter < history_cap) {   // :115-117 residual_norms->at(iter)
            history[iter] = rn;
            n_history = iter + 1;
        }
        res = rn;                              // :119
        if (iter >= p.max_iter) stop = true;   // :124-126
        if (rn < p.tolerance) stop = true;     // :128-130
        if (p.rel_tol > 0 && rn < p.rel_tol * init_res) stop = true;  // :132-136
        ++iter;                                // :143
        return stop;
    }
};

void axpy(const Sys &s, orc_scalar a, const dvec &x, dvec &y)
{
    for (int q = 0; q < s.R; ++q) {
        const orc_label n = s.rk[q].n;
        const orc_scalar *xv = x[q].data();
        orc_scalar *yv = y[q].data();
#pragma omp parallel for schedule(static) num_threads(s.threads) if (s.threads > 1)
        for (orc_label i = 0; i < n; ++i) yv[i] += a * xv[i];
    }
}

// ---- solvers ---------------------------------------------------------------

// Ginkgo core/solver/cg.cpp + reference/solver/cg_kernels.cpp
void run_cg(const Sys &s, const Jacobi &J, Criterion &crit, dvec &x, const dvec &b)
{
    dvec r = b, z = zeros_like(s), p = zeros_like(s), q = zeros_like(s);
    orc_scalar rho = 0.0, prev_rho = 1.0, beta = 0.0;
    spmv(s, x, r, true, -1.0, 1.0);  // r = b - A x
    while (true) {
        precond_apply(s, J, r, z);
        rho = dot(s, r, z);
        if (crit.check(r)) break;
        // step_1: p = z + (rho / prev_rho) p
        {
            const bool zero = (prev_rho == 0.0);
            const orc_scalar t = zero ? 0.0 : rho / prev_rho;
            for (int k = 0; k < s.R; ++k) {
                const orc_label n = s.rk[k].n;
                orc_scalar *pv = p[k].data();
                const orc_scalar *zv = z[k].data();
#pragma omp parallel for schedule(static) num_threads(s.threads) if (s.threads > 1)
                for (orc_label i = 0; i < n; ++i)
                    pv[i] = zero ? zv[i] : zv[i] + t * pv[i];
            }
        }
        spmv(s, p, q);
        beta = dot(s, p, q);
        // step_2: x += (rho / beta) p ; r -= (rho / beta) q
        if (beta != 0.0) {
            const orc_scalar t = rho / beta;
            for (int k = 0; k < s.R; ++k) {
                const orc_label n = s.rk[k].n;
                orc_scalar *xv = x[k].data(), *rv = r[k].data();
                const orc_scalar *pv = p[k].data(), *qv = q[k].data();
#pragma omp parallel for schedule(static) num_threads(s.threads) if (s.threads > 1)
                for (orc_label i = 0; i < n; ++i) {
                    xv[i] += t * pv[i];
                    rv[i] -= t * qv[i];
                }
            }
        }
        std::swap(prev_rho, rho);
    }
}

// Ginkgo core/solver/bicgstab.cpp + reference/solver/bicgstab_kernels.cpp
void run_bicgstab(const Sys &s, const Jacobi &J, Criterion &crit, dvec &x,
                  const dvec &b)
{
    dvec r = b, rr, y = zeros_like(s), sv = zeros_like(s), t = zeros_like(s),
         z = zeros_like(s), v = zeros_like(s), p = zeros_like(s);
    orc_scalar prev_rho = 1.0, rho = 1.0, alpha = 1.0, beta = 1.0, gamma = 1.0,
               omega = 1.0;
    spmv(s, x, r, true, -1.0, 1.0);
    rr = r;
    while (true) {
        rho = dot(s, rr, r);
        if (crit.check(r)) break;
        // step_1
        {
            const bool ok = (prev_rho * omega != 0.0);
            const orc_scalar tmp = ok ? rho / prev_rho * alpha / omega : 0.0;
            for (int k = 0; k < s.R; ++k)
                for (orc_label i = 0; i < s.rk[k].n; ++i)
                    p[k][i] = ok ? r[k][i] + tmp * (p[k][i] - omega * v[k][i])
                                 : r[k][i];
        }
        precond_apply(s, J, p, y);
        spmv(s, y, v);
        beta = dot(s, rr, v);
        // step_2
        if (beta != 0.0) {
            alpha = rho / beta;
            for (int k = 0; k < s.R; ++k)
                for (orc_label i = 0; i < s.rk[k].n; ++i)
                    sv[k][i] = r[k][i] - alpha * v[k][i];
        } else {
            alpha = 0.0;
            sv = r;
        }
        if (crit.check(sv)) {
            axpy(s, alpha, y, x);  // finalize: x += alpha y
            break;
        }
        precond_apply(s, J, sv, z);
        spmv(s, z, t);
        gamma = dot(s, sv, t);
        beta = dot(s, t, t);
        // step_3
        omega = (beta != 0.0) ? gamma / beta : 0.0;
        for (int k = 0; k < s.R; ++k)
            for (orc_label i = 0; i < s.rk[k].n; ++i) {
                x[k][i] += alpha * y[k][i] + omega * z[k][i];
                r[k][i] = sv[k][i] - omega * t[k][i];
            }
        std::swap(prev_rho, rho);
    }
}

// Ginkgo core/solver/gmres.cpp (modified Gram-Schmidt) +
// reference/solver/{gmres,common_gmres}_kernels.cpp.  The criterion is handed
// `.residual(residual)`, a vector Ginkgo refreshes only at (re)starts
// (SURVEY.md Appendix B-8): between restarts OGL's L1 criterion sees the
// residual of the last restart.  That behaviour is restated as is.
void run_gmres(const Sys &s, const Jacobi &J, Criterion &crit, dvec &x,
               const dvec &b)
{
    const orc_label m = crit.p.krylov_dim > 0 ? crit.p.krylov_dim : 100;
    dvec residual = b;
    spmv(s, x, residual, true, -1.0, 1.0);
    orc_scalar res_norm = norm2(s, residual);
    std::vector<dvec> V(static_cast<size_t>(m) + 1);
    std::vector<orc_scalar> H(static_cast<size_t>(m + 1) * m, 0.0);  // H(i,j)=H[i*m+j]
    std::vector<orc_scalar> gs(m, 0.0), gc(m, 0.0), g(static_cast<size_t>(m) + 1, 0.0),
        yv(m, 0.0);
    orc_label final_iter = 0;
    auto restart = [&]() {
        g.assign(static_cast<size_t>(m) + 1, 0.0);
        g[0] = res_norm;
        V[0] = zeros_like(s);
        for (int k = 0; k < s.R; ++k)
            for (orc_label i = 0; i < s.rk[k].n; ++i)
                V[0][k][i] = residual[k][i] / res_norm;
        final_iter = 0;
    };
    auto update_x = [&]() {
        // solve_krylov: back substitution on the rotated Hessenberg
        for (orc_label i = final_iter - 1; i >= 0; --i) {
            orc_scalar t = g[i];
            for (orc_label j = i + 1; j < final_iter; ++j) t -= H[i * m + j] * yv[j];
            yv[i] = t / H[i * m + i];
        }
        // multi_axpy: before_precond = V y ; x += M^-1 before_precond
        dvec bp = zeros_like(s), ap = zeros_like(s);
        for (int k = 0; k < s.R; ++k)
            for (orc_label i = 0; i < s.rk[k].n; ++i) {
                orc_scalar acc = 0.0;
                for (orc_label j = 0; j < final_iter; ++j) acc += V[j][k][i] * yv[j];
                bp[k][i] = acc;
            }
        precond_apply(s, J, bp, ap);
        axpy(s, 1.0, ap, x);
    };
    restart();
    orc_label ri = 0;
    dvec pv = zeros_like(s), w = zeros_like(s);
    while (true) {
        if (crit.check(residual)) break;
        if (ri == m) {
            update_x();
            residual = b;
            spmv(s, x, residual, true, -1.0, 1.0);
            res_norm = norm2(s, residual);
            restart();
            ri = 0;
        }
        precond_apply(s, J, V[ri], pv);
        spmv(s, pv, w);
        // modified Gram-Schmidt against V[0..ri]
        for (orc_label k = 0; k <= ri; ++k) {
            const orc_scalar h = dot(s, w, V[k]);
            H[k * m + ri] = h;
            axpy(s, -h, V[k], w);
        }
        const orc_scalar hn = norm2(s, w);
        H[(ri + 1) * m + ri] = hn;
        V[ri + 1] = zeros_like(s);
        for (int k = 0; k < s.R; ++k)
            for (orc_label i = 0; i < s.rk[k].n; ++i) V[ri + 1][k][i] = w[k][i] / hn;
        // hessenberg_qr
        ++final_iter;
        for (orc_label j = 0; j < ri; ++j) {
            const orc_scalar t = gc[j] * H[j * m + ri] + gs[j] * H[(j + 1) * m + ri];
            H[(j + 1) * m + ri] = -gs[j] * H[j * m + ri] + gc[j] * H[(j + 1) * m + ri];
            H[j * m + ri] = t;
        }
        const orc_scalar ha = H[ri * m + ri], hb = H[(ri + 1) * m + ri];
        if (ha == 0.0) {
            gc[ri] = 0.0;
            gs[ri] = 1.0;
        } else {
            const orc_scalar scale = std::fabs(ha) + std::fabs(hb);
            const orc_scalar hyp =
                scale * std::sqrt((ha / scale) * (ha / scale) + (hb / scale) * (hb / scale));
            gc[ri] = ha / hyp;
            gs[ri] = hb / hyp;
        }
        H[ri * m + ri] = gc[ri] * ha + gs[ri] * hb;
        H[(ri + 1) * m + ri] = 0.0;
        g[ri + 1] = -gs[ri] * g[ri];
        g[ri] = gc[ri] * g[ri];
        res_norm = std::fabs(g[ri + 1]);
        ++ri;
    }
    update_x();
}

}  // namespace

extern "C" {

void orc_dist_spmv(int n_ranks, const orc_rank_system *ranks,
                   const orc_scalar *const *xs, orc_scalar *const *ys)
{
    Sys s = make_sys(n_ranks, ranks, 1);
    dvec x(n_ranks), y = zeros_like(s);
    for (int r = 0; r < n_ranks; ++r) x[r].assign(xs[r], xs[r] + ranks[r].n);
    spmv(s, x, y);
    for (int r = 0; r < n_ranks; ++r)
        std::memcpy(ys[r], y[r].data(), sizeof(orc_scalar) * ranks[r].n);
}

int orc_solve(int n_ranks, const orc_rank_system *ranks,
              const orc_solve_params *params, orc_solve_result *result,
              orc_scalar *history, orc_label history_cap)
{
    if (n_ranks < 1 || !ranks || !params || !result) return 1;
    if (params->frequency < 1) return 2;
    const int threads = params->threads > 1 ? params->threads : 1;
    Sys s = make_sys(n_ranks, ranks, threads);
    for (int r = 0; r < n_ranks; ++r)
        for (auto u : s.peer_slot[r])
            if (u < 0) return 3;  // asymmetric neighbour lists
    dvec x(n_ranks), b(n_ranks);
    for (int r = 0; r < n_ranks; ++r) {
        x[r].assign(ranks[r].x, ranks[r].x + ranks[r].n);
        b[r].assign(ranks[r].b, ranks[r].b + ranks[r].n);
    }
    Jacobi J = make_jacobi(s, params->precond, params->max_block_size, params);
    if (J.kind < 0) return 5;  // ISAI: a row longer than the dense solver handles / no diagonal
    Criterion crit{s, *params, x, b, 0, 1.0, 0.0, 0.0, history, history_cap, 0};
    const auto t0 = std::chrono::steady_clock::now();
    switch (params->solver) {
    case ORC_CG:
        run_cg(s, J, crit, x, b);
        break;
    case ORC_BICGSTAB:
        run_bicgstab(s, J, crit, x, b);
        break;
    case ORC_GMRES:
        run_gmres(s, J, crit, x, b);
        break;
    default:
        return 4;
    }
    const auto t1 = std::chrono::steady_clock::now();
    for (int r = 0; r < n_ranks; ++r)
        std::memcpy(ranks[r].x, x[r].data(), sizeof(orc_scalar) * ranks[r].n);
    result->init_residual = crit.init_res;
    result->final_residual = crit.res;
    result->criterion_calls = crit.iter;
    // GKOBiCGStab.H:112-115 reports calls/2, the others the raw counter
    result->n_iterations = params->solver == ORC_BICGSTAB ? crit.iter / 2 : crit.iter;
    result->norm_factor = crit.norm_factor;
    result->n_history = crit.n_history;
    result->seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}

orc_label orc_bj_find_blocks(orc_label n, const orc_label *row_ptrs,
                             const orc_label *cols, orc_label max_block_size,
                             orc_label *block_ptrs)
{
    return find_blocks(n, row_ptrs, cols, max_block_size < 1 ? 1 : max_block_size,
                       block_ptrs);
}

int orc_isai_generate(orc_label n, const orc_label *row_ptrs, const orc_label *cols, const orc_scalar *vals,
                      int spd, orc_scalar *w, orc_scalar *wt)
{
    return isai_generate(n, row_ptrs, cols, vals, spd != 0, w, wt) ? 0 : 1;
}

void *orc_mg_create(orc_label n, const orc_label *row_ptrs, const orc_label *cols, const orc_scalar *vals,
                    int max_levels, orc_label min_coarse_rows, int coarse_iters)
{
    orc_mg::Csr A;
    A.n = n;
    A.rp.assign(row_ptrs, row_ptrs + n + 1);
    A.cols.assign(cols, cols + row_ptrs[n]);
    A.vals.assign(vals, vals + row_ptrs[n]);
    return new orc_mg::Hierarchy(orc_mg::build(std::move(A), max_levels, min_coarse_rows, coarse_iters));
}

void orc_mg_destroy(void *h) { delete static_cast<orc_mg::Hierarchy *>(h); }

int orc_mg_levels(const void *h) { return static_cast<int>(static_cast<const orc_mg::Hierarchy *>(h)->levels.size()); }

int orc_mg_level_info(const void *h, int level, orc_label *n, orc_label *nnz, orc_label *n_coarse)
{
    const auto &H = *static_cast<const orc_mg::Hierarchy *>(h);
    if (level < 0 || level >= static_cast<int>(H.levels.size())) return 1;
    const auto &L = H.levels[level];
    *n = L.A.n;
    *nnz = static_cast<orc_label>(L.A.vals.size());
    *n_coarse = L.agg.empty() ? 0 : L.n_coarse;
    return 0;
}

int orc_mg_level_get(const void *h, int level, orc_label *row_ptrs, orc_label *cols, orc_scalar *vals,
                     orc_label *agg)
{
    const auto &H = *static_cast<const orc_mg::Hierarchy *>(h);
    if (level < 0 || level >= static_cast<int>(H.levels.size())) return 1;
    const auto &L = H.levels[level];
    std::copy(L.A.rp.begin(), L.A.rp.end(), row_ptrs);
    std::copy(L.A.cols.begin(), L.A.cols.end(), cols);
    std::copy(L.A.vals.begin(), L.A.vals.end(), vals);
    if (agg) std::copy(L.agg.begin(), L.agg.end(), agg);
    return 0;
}

void orc_mg_apply(const void *h, const orc_scalar *r, orc_scalar *z)
{
    const auto &H = *static_cast<const orc_mg::Hierarchy *>(h);
    std::fill(z, z + H.levels[0].A.n, 0.0);
    orc_mg::vcycle(H, 0, r, z);
}

int orc_trifactor(int kind, orc_label n, const orc_label *row_ptrs, const orc_label *cols,
                  const orc_scalar *vals, orc_scalar *factors)
{
    std::vector<orc_label> dp;
    if (!orc_tri::diag_positions(n, row_ptrs, cols, dp)) return 1;
    std::memcpy(factors, vals, sizeof(orc_scalar) * static_cast<size_t>(row_ptrs[n]));
    if (kind == ORC_PRECOND_IC) return orc_tri::ic0(n, row_ptrs, cols, dp.data(), factors) ? 0 : 2;
    if (kind != ORC_PRECOND_ILU && kind != ORC_PRECOND_IRILU) return 3;
    orc_tri::ilu0(n, row_ptrs, cols, dp.data(), factors);
    return 0;
}

int orc_trifactor_apply(int kind, orc_label n, const orc_label *row_ptrs, const orc_label *cols,
                        const orc_scalar *factors, const orc_scalar *r, orc_scalar *z)
{
    std::vector<orc_label> dp;
    if (!orc_tri::diag_positions(n, row_ptrs, cols, dp)) return 1;
    std::vector<orc_scalar> t(r, r + n), scratch;
    if (kind == ORC_PRECOND_IRILU) {
        orc_tri::ir_jacobi(n, row_ptrs, cols, dp.data(), factors, true, 5, r, t.data(), scratch);
        std::memcpy(z, t.data(), sizeof(orc_scalar) * static_cast<size_t>(n));
        orc_tri::ir_jacobi(n, row_ptrs, cols, dp.data(), factors, false, 5, t.data(), z, scratch);
        return 0;
    }
    if (kind != ORC_PRECOND_ILU && kind != ORC_PRECOND_IC) return 3;
    orc_tri::lower_solve(n, row_ptrs, cols, dp.data(), factors, kind == ORC_PRECOND_ILU, r, t.data());
    orc_tri::upper_solve(n, row_ptrs, cols, dp.data(), factors, t.data(), z);
    return 0;
}

void orc_bj_invert_blocks(orc_label n, const orc_label *row_ptrs,
                          const orc_label *cols, const orc_scalar *vals,
                          orc_label n_blocks, const orc_label *block_ptrs,
                          orc_scalar *inv)
{
    (void)n;
    size_t off = 0;
    for (orc_label b = 0; b < n_blocks; ++b) {
        const orc_label lo = block_ptrs[b], hi = block_ptrs[b + 1], sz = hi - lo;
        orc_scalar *m = inv + off;
        for (orc_label i = 0; i < sz * sz; ++i) m[i] = 0.0;
        for (orc_label i = lo; i < hi; ++i)
            for (orc_label e = row_ptrs[i]; e < row_ptrs[i + 1]; ++e)
                if (cols[e] >= lo && cols[e] < hi)
                    m[(i - lo) * sz + (cols[e] - lo)] = vals[e];
        invert_block(sz, m);
        off += static_cast<size_t>(sz) * sz;
    }
}

double orc_time_spmv(orc_label n, const orc_label *row_ptrs,
                     const orc_label *cols, const orc_scalar *vals,
                     const orc_scalar *x, orc_scalar *y, int reps, int threads)
{
    const int nt = threads > 1 ? threads : 1;
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < reps; ++it) {
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1)
        for (orc_label i = 0; i < n; ++i) {
            orc_scalar sum = 0.0;
            for (orc_label e = row_ptrs[i]; e < row_ptrs[i + 1]; ++e)
                sum += vals[e] * x[cols[e]];
            y[i] = sum;
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
