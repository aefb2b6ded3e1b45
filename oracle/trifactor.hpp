// TEST INFRASTRUCTURE (oracle) -- incomplete factorisations and triangular solves behind the
// `preconditioner` keywords ILU, IC and IRILU (Preconditioner/Preconditioner.H:106-124, 143-176,
// 177-196).  PARITY UNPINNED: the arithmetic lives in Ginkgo (absent, see krylov.cpp header); this
// restates the published algorithms of
//   gko::factorization::Ilu  (exact ILU(0) on the pattern of A, L unit lower, U upper with diagonal),
//   gko::factorization::Ic   (exact IC(0) on the lower pattern, L L^H ~ A),
//   gko::preconditioner::Ilu<LowerTrs, UpperTrs>  z = U^-1 (L^-1 r),
//   gko::preconditioner::Ic<LowerTrs>             z = L^-H (L^-1 r),
//   gko::preconditioner::Ilu<Ir, Ir> with Ir = 5 sweeps of Richardson iteration preconditioned by
//   scalar Jacobi (relaxation 1, initial guess = the right-hand side).
// Everything is sequential in row order, products rounded before they are added
// (-ffp-contract=off), columns inside a row ascending.  The factors live over the CSR pattern of A
// ("LU in place"): strictly lower part = L (unit diagonal implied) / IC: L with its diagonal, upper
// part = U / IC: L^T mirrored, so that both sweeps read rows.
#pragma once

#include <cmath>
#include <vector>

#include "oracle.h"

namespace orc_tri {

// position of the diagonal entry of every row, -1 if missing; false when a row holds a column twice
inline bool diag_positions(orc_label n, const orc_label *rp, const orc_label *cols, std::vector<orc_label> &dp)
{
    dp.assign(static_cast<size_t>(n), -1);
    for (orc_label i = 0; i < n; ++i) {
        for (orc_label e = rp[i]; e < rp[i + 1]; ++e) {
            if (e > rp[i] && cols[e] <= cols[e - 1]) return false;
            if (cols[e] == i) dp[i] = e;
        }
        if (dp[i] < 0) return false;
    }
    return true;
}

// ILU(0), IKJ order: for every lower entry (i,c) ascending: l = a_ic / u_cc; row_i -= l * U(c, >c)
// restricted to the pattern of row i.
inline void ilu0(orc_label n, const orc_label *rp, const orc_label *cols, const orc_label *dp, orc_scalar *F)
{
    for (orc_label i = 0; i < n; ++i) {
        const orc_label end = rp[i + 1];
        for (orc_label k = rp[i]; k < dp[i]; ++k) {
            const orc_label c = cols[k];
            const orc_scalar l = F[k] / F[dp[c]];
            F[k] = l;
            orc_label p = k + 1;
            for (orc_label j = dp[c] + 1; j < rp[c + 1]; ++j) {
                const orc_label cj = cols[j];
                while (p < end && cols[p] < cj) ++p;
                if (p < end && cols[p] == cj) F[p] = F[p] - l * F[j];
            }
        }
    }
}

// IC(0): l_ij = (a_ij - sum_{k<j} l_ik l_jk) / l_jj, l_ii = sqrt(a_ii - sum_k l_ik^2); the upper
// positions receive the transpose.  false when the pattern is not structurally symmetric.
inline bool ic0(orc_label n, const orc_label *rp, const orc_label *cols, const orc_label *dp, orc_scalar *F)
{
    for (orc_label i = 0; i < n; ++i) {
        for (orc_label k = rp[i]; k <= dp[i]; ++k) {
            const orc_label j = cols[k];
            orc_scalar s = F[k];
            orc_label a = rp[i], b = rp[j];
            const orc_label enda = k, endb = dp[j];
            while (a < enda && b < endb) {
                const orc_label ca = cols[a], cb = cols[b];
                if (ca == cb) {
                    s = s - F[a] * F[b];
                    ++a;
                    ++b;
                } else if (ca < cb) {
                    ++a;
                } else {
                    ++b;
                }
            }
            F[k] = j < i ? s / F[dp[j]] : std::sqrt(s);
        }
        for (orc_label k = rp[i]; k < dp[i]; ++k) {
            const orc_label j = cols[k];
            bool found = false;
            for (orc_label q = dp[j] + 1; q < rp[j + 1]; ++q)
                if (cols[q] == i) {
                    F[q] = F[k];
                    found = true;
                    break;
                }
            if (!found) return false;
        }
    }
    return true;
}

// LowerTrs: x_i = (b_i - sum_{c<i} F_ic x_c) / (unit ? 1 : F_ii)
inline void lower_solve(orc_label n, const orc_label *rp, const orc_label *cols, const orc_label *dp,
                        const orc_scalar *F, bool unit, const orc_scalar *b, orc_scalar *x)
{
    for (orc_label i = 0; i < n; ++i) {
        orc_scalar s = b[i];
        for (orc_label e = rp[i]; e < dp[i]; ++e) s = s - F[e] * x[cols[e]];
        x[i] = unit ? s : s / F[dp[i]];
    }
}

// UpperTrs: x_i = (b_i - sum_{c>i} F_ic x_c) / F_ii, rows descending
inline void upper_solve(orc_label n, const orc_label *rp, const orc_label *cols, const orc_label *dp,
                        const orc_scalar *F, const orc_scalar *b, orc_scalar *x)
{
    for (orc_label i = n - 1; i >= 0; --i) {
        orc_scalar s = b[i];
        for (orc_label e = dp[i] + 1; e < rp[i + 1]; ++e) s = s - F[e] * x[cols[e]];
        x[i] = s / F[dp[i]];
    }
}

// gko::solver::Ir with a scalar-Jacobi inner solver on a triangular factor T (lower: unit diagonal),
// `sweeps` iterations from the initial guess already in x:  x <- x + D^-1 (b - T x).
inline void ir_jacobi(orc_label n, const orc_label *rp, const orc_label *cols, const orc_label *dp,
                      const orc_scalar *F, bool lower, int sweeps, const orc_scalar *b, orc_scalar *x,
                      std::vector<orc_scalar> &scratch)
{
    scratch.resize(static_cast<size_t>(n));
    for (int it = 0; it < sweeps; ++it) {
        for (orc_label i = 0; i < n; ++i) {
            // residual row: b_i - (T x)_i, the row sum left to right as a CSR apply does
            orc_scalar t = 0.0;
            if (lower) {
                for (orc_label e = rp[i]; e < dp[i]; ++e) t = t + F[e] * x[cols[e]];
                t = t + 1.0 * x[i];
                scratch[i] = x[i] + (b[i] - t);              // D = I
            } else {
                for (orc_label e = dp[i]; e < rp[i + 1]; ++e) t = t + F[e] * x[cols[e]];
                scratch[i] = x[i] + (b[i] - t) * (1.0 / F[dp[i]]);   // jacobi::scalar_apply: times the inverted diagonal
            }
        }
        for (orc_label i = 0; i < n; ++i) x[i] = scratch[i];
    }
}

}  // namespace orc_tri
