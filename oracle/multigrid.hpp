// TEST INFRASTRUCTURE (oracle) -- the `preconditioner Multigrid` keyword
// (Preconditioner/Preconditioner.H:261-341).  PARITY UNPINNED: the algorithm lives in Ginkgo
// (absent, see krylov.cpp header); this restates the published algorithms of
//   gko::multigrid::Pgm (deterministic)     parallel graph match: size-2 aggregation by mutually
//                                           strongest neighbours, <= 15 matching rounds, leftovers
//                                           joined to the strongest aggregated neighbour
//   gko::solver::Multigrid                  one V cycle per apply from a zero guess; pre- and
//                                           post-smoother = Ir(2 sweeps, relaxation 0.9, scalar
//                                           Jacobi); coarsest solver = 4 (coarseSolverIters)
//                                           unpreconditioned CG iterations from a zero guess;
//                                           at most maxLevels (9) coarsenings while the matrix has
//                                           more than minCoarseRows (10) rows
// as OGL configures them, on the LOCAL block of a rank (wrap_schwarz, :66-82).
//
// One deliberate choice: Ginkgo's find_strongest_neighbor kernel lets a row whose neighbours are all
// aggregated join an aggregate WHILE the other rows of the same round still read the aggregate
// array (sequential on the reference executor, racing on a device).  Restated here with the
// semantics of the data-parallel kernel run without the race: every round reads the aggregates as
// they were when the round started.  The sort that groups the Galerkin product's entries is stable
// (duplicates are added in the fine matrix's storage order).
#pragma once

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "oracle.h"

namespace orc_mg {

struct Csr {
    orc_label n = 0;
    std::vector<orc_label> rp, cols;
    std::vector<orc_scalar> vals;
};

struct Level {
    Csr A;
    std::vector<orc_scalar> inv_diag;   // jacobi::invert_diagonal of A
    std::vector<orc_label> agg;         // fine row -> coarse row (empty on the coarsest level)
    orc_label n_coarse = 0;
};

struct Hierarchy {
    std::vector<Level> levels;   // levels.back(): the coarsest matrix
    int coarse_iters = 4;
};

// weight of entry e = (row, col): (|a_rc| + |a_cr|) / 2 (0.5 |A| + 0.5 |A|^T), a_cr = 0 if absent
inline void weights(const Csr &A, std::vector<orc_scalar> &w, std::vector<orc_scalar> &diag)
{
    w.assign(A.vals.size(), 0.0);
    diag.assign(static_cast<size_t>(A.n), 0.0);
    for (orc_label i = 0; i < A.n; ++i)
        for (orc_label e = A.rp[i]; e < A.rp[i + 1]; ++e) {
            const orc_label c = A.cols[e];
            orc_scalar t = 0.0;
            for (orc_label q = A.rp[c]; q < A.rp[c + 1]; ++q)
                if (A.cols[q] == i) {
                    t = std::fabs(A.vals[q]);
                    break;
                }
            w[e] = 0.5 * std::fabs(A.vals[e]) + 0.5 * t;
        }
    for (orc_label i = 0; i < A.n; ++i)
        for (orc_label e = A.rp[i]; e < A.rp[i + 1]; ++e)
            if (A.cols[e] == i) {
                diag[i] = w[e];
                break;
            }
}

// pgm: returns the number of aggregates, agg[i] in [0, n_agg)
inline orc_label aggregate(const Csr &A, std::vector<orc_label> &agg, int max_iterations = 15,
                           double max_unassigned_ratio = 0.05)
{
    const orc_label n = A.n;
    std::vector<orc_scalar> w, diag;
    weights(A, w, diag);
    agg.assign(static_cast<size_t>(n), -1);
    std::vector<orc_label> strongest(static_cast<size_t>(n), -1), snapshot;
    orc_label num_unagg = 0, num_unagg_prev = 0;
    auto weight_of = [&](orc_label row, orc_label e) {
        const orc_label c = A.cols[e];
        return w[e] / std::max(std::fabs(diag[row]), std::fabs(diag[c]));
    };
    for (int it = 0; it < max_iterations; ++it) {
        // find_strongest_neighbor
        snapshot = agg;
        for (orc_label row = 0; row < n; ++row) {
            if (snapshot[row] != -1) continue;
            orc_scalar max_unagg = 0.0, max_agg = 0.0;
            orc_label s_unagg = -1, s_agg = -1;
            for (orc_label e = A.rp[row]; e < A.rp[row + 1]; ++e) {
                const orc_label c = A.cols[e];
                if (c == row) continue;
                const orc_scalar wt = weight_of(row, e);
                if (snapshot[c] == -1 && (wt > max_unagg || (wt == max_unagg && c > s_unagg))) {
                    max_unagg = wt;
                    s_unagg = c;
                } else if (snapshot[c] != -1 && (wt > max_agg || (wt == max_agg && c > s_agg))) {
                    max_agg = wt;
                    s_agg = c;
                }
            }
            if (s_unagg == -1 && s_agg != -1) agg[row] = snapshot[s_agg];   // all neighbours aggregated
            else if (s_unagg != -1) strongest[row] = s_unagg;
            else strongest[row] = row;                                      // no neighbour
        }
        // match_edge: mutually strongest pairs, the smaller index names the aggregate
        for (orc_label i = 0; i < n; ++i) {
            if (agg[i] != -1) continue;
            const orc_label nb = strongest[i];
            if (nb != -1 && strongest[nb] == i && i <= nb) {
                agg[i] = i;
                agg[nb] = i;
            }
        }
        num_unagg = 0;
        for (orc_label i = 0; i < n; ++i) num_unagg += agg[i] == -1;
        if (num_unagg == 0 || num_unagg == num_unagg_prev || num_unagg < max_unassigned_ratio * n) break;
        num_unagg_prev = num_unagg;
    }
    if (num_unagg != 0) {
        // assign_to_exist_agg (deterministic: reads a copy)
        snapshot = agg;
        for (orc_label row = 0; row < n; ++row) {
            if (snapshot[row] != -1) continue;
            orc_scalar max_agg = 0.0;
            orc_label s_agg = -1;
            for (orc_label e = A.rp[row]; e < A.rp[row + 1]; ++e) {
                const orc_label c = A.cols[e];
                if (c == row) continue;
                const orc_scalar wt = weight_of(row, e);
                if (snapshot[c] != -1 && (wt > max_agg || (wt == max_agg && c > s_agg))) {
                    max_agg = wt;
                    s_agg = c;
                }
            }
            agg[row] = s_agg != -1 ? snapshot[s_agg] : row;
        }
    }
    // renumber: aggregate names (root rows) -> 0 .. n_agg-1 in ascending root order
    std::vector<orc_label> map(static_cast<size_t>(n) + 1, 0);
    for (orc_label i = 0; i < n; ++i) map[agg[i]] = 1;
    orc_label run = 0;
    for (orc_label i = 0; i <= n; ++i) {
        const orc_label v = map[i];
        map[i] = run;
        run += v;
    }
    for (orc_label i = 0; i < n; ++i) agg[i] = map[agg[i]];
    return map[n];
}

// Galerkin product with the piecewise-constant prolongation: A_c(I,J) = sum a_ij, agg[i]=I, agg[j]=J
inline Csr coarsen(const Csr &A, const std::vector<orc_label> &agg, orc_label n_coarse)
{
    const size_t nnz = A.vals.size();
    std::vector<orc_label> row_of(nnz);
    for (orc_label i = 0; i < A.n; ++i)
        for (orc_label e = A.rp[i]; e < A.rp[i + 1]; ++e) row_of[e] = i;
    std::vector<size_t> order(nnz);
    std::iota(order.begin(), order.end(), size_t{0});
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        const orc_label ra = agg[row_of[a]], rb = agg[row_of[b]];
        if (ra != rb) return ra < rb;
        return agg[A.cols[a]] < agg[A.cols[b]];
    });
    Csr C;
    C.n = n_coarse;
    C.rp.assign(static_cast<size_t>(n_coarse) + 1, 0);
    for (size_t k = 0; k < nnz; ++k) {
        const size_t e = order[k];
        const orc_label r = agg[row_of[e]], c = agg[A.cols[e]];
        if (k > 0) {
            const size_t p = order[k - 1];
            if (agg[row_of[p]] == r && agg[A.cols[p]] == c) {
                C.vals.back() = C.vals.back() + A.vals[e];
                continue;
            }
        }
        C.cols.push_back(c);
        C.vals.push_back(A.vals[e]);
        C.rp[r + 1]++;
    }
    for (orc_label i = 0; i < n_coarse; ++i) C.rp[i + 1] += C.rp[i];
    return C;
}

inline void invert_diagonal(Level &L)
{
    L.inv_diag.assign(static_cast<size_t>(L.A.n), 0.0);
    for (orc_label i = 0; i < L.A.n; ++i)
        for (orc_label e = L.A.rp[i]; e < L.A.rp[i + 1]; ++e)
            if (L.A.cols[e] == i) {
                L.inv_diag[i] = 1.0 / L.A.vals[e];
                break;
            }
}

// Multigrid::generate: coarsen while level < max_levels and rows > min_coarse_rows; stop when a
// coarsening does not shrink the matrix
inline Hierarchy build(Csr A, int max_levels, orc_label min_coarse_rows, int coarse_iters)
{
    Hierarchy H;
    H.coarse_iters = coarse_iters;
    H.levels.emplace_back();
    H.levels.back().A = std::move(A);
    invert_diagonal(H.levels.back());
    int level = 0;
    while (level < max_levels && H.levels.back().A.n > min_coarse_rows) {
        Level &L = H.levels.back();
        std::vector<orc_label> agg;
        const orc_label nc = aggregate(L.A, agg);
        if (nc == L.A.n) break;
        Csr C = coarsen(L.A, agg, nc);
        L.agg = std::move(agg);
        L.n_coarse = nc;
        H.levels.emplace_back();
        H.levels.back().A = std::move(C);
        invert_diagonal(H.levels.back());
        ++level;
    }
    return H;
}

inline void spmv(const Csr &A, const orc_scalar *x, orc_scalar *y)
{
    for (orc_label i = 0; i < A.n; ++i) {
        orc_scalar s = 0.0;
        for (orc_label e = A.rp[i]; e < A.rp[i + 1]; ++e) s += A.vals[e] * x[A.cols[e]];
        y[i] = s;
    }
}

// Ir(max 2 iterations, relaxation 0.9, scalar Jacobi): x += 0.9 D^-1 (b - A x), twice; with a zero
// guess the first sweep is x = 0.9 D^-1 b
inline void smooth(const Level &L, const orc_scalar *b, orc_scalar *x, bool x_is_zero, std::vector<orc_scalar> &r)
{
    const orc_label n = L.A.n;
    r.resize(static_cast<size_t>(n));
    for (int it = 0; it < 2; ++it) {
        if (it == 0 && x_is_zero) {
            for (orc_label i = 0; i < n; ++i) x[i] = (0.9 * b[i]) * L.inv_diag[i];   // jacobi scalar_apply: alpha * b * inv_diag
            continue;
        }
        spmv(L.A, x, r.data());
        for (orc_label i = 0; i < n; ++i) {
            const orc_scalar res = b[i] - r[i];
            x[i] = x[i] + (0.9 * res) * L.inv_diag[i];
        }
    }
}

// `iters` CG iterations from x = 0 (Ginkgo cg.cpp operation order, identity preconditioner)
inline void coarse_cg(const Csr &A, const orc_scalar *b, orc_scalar *x, int iters)
{
    const orc_label n = A.n;
    std::vector<orc_scalar> r(b, b + n), p(static_cast<size_t>(n), 0.0), q(static_cast<size_t>(n));
    for (orc_label i = 0; i < n; ++i) x[i] = 0.0;
    orc_scalar rho = 0.0, prev_rho = 1.0;
    for (orc_label i = 0; i < n; ++i) rho += r[i] * r[i];
    for (int it = 0; it < iters; ++it) {
        // step_1: p = z + (rho / prev_rho) p   (p = z when prev_rho == 0)
        const orc_scalar tmp = prev_rho == 0.0 ? 0.0 : rho / prev_rho;
        for (orc_label i = 0; i < n; ++i) p[i] = prev_rho == 0.0 ? r[i] : r[i] + tmp * p[i];
        spmv(A, p.data(), q.data());
        orc_scalar beta = 0.0;
        for (orc_label i = 0; i < n; ++i) beta += p[i] * q[i];
        // step_2: x += (rho / beta) p, r -= (rho / beta) q   (skipped when beta == 0)
        if (beta != 0.0) {
            const orc_scalar a = rho / beta;
            for (orc_label i = 0; i < n; ++i) {
                x[i] = x[i] + a * p[i];
                r[i] = r[i] - a * q[i];
            }
        }
        prev_rho = rho;
        rho = 0.0;
        for (orc_label i = 0; i < n; ++i) rho += r[i] * r[i];
    }
}

// one V cycle on level l: x (zero on entry) <- approximate solution of A_l x = b
inline void vcycle(const Hierarchy &H, size_t l, const orc_scalar *b, orc_scalar *x)
{
    const Level &L = H.levels[l];
    const orc_label n = L.A.n;
    if (l + 1 == H.levels.size()) {
        // no coarsening happened at all: the "coarsest solver" is the whole preconditioner
        coarse_cg(L.A, b, x, H.coarse_iters);
        return;
    }
    std::vector<orc_scalar> r, tmp(static_cast<size_t>(n));
    smooth(L, b, x, true, r);
    spmv(L.A, x, tmp.data());
    for (orc_label i = 0; i < n; ++i) tmp[i] = b[i] - tmp[i];
    // restrict: g_I = sum of the aggregate's residuals, members in ascending fine order
    std::vector<orc_scalar> g(static_cast<size_t>(L.n_coarse), 0.0), e(static_cast<size_t>(L.n_coarse), 0.0);
    for (orc_label i = 0; i < n; ++i) g[L.agg[i]] += tmp[i];
    if (l + 2 == H.levels.size()) coarse_cg(H.levels[l + 1].A, g.data(), e.data(), H.coarse_iters);
    else vcycle(H, l + 1, g.data(), e.data());
    for (orc_label i = 0; i < n; ++i) x[i] = x[i] + e[L.agg[i]];
    smooth(L, b, x, false, r);
}

}  // namespace orc_mg
