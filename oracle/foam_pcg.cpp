// TEST INFRASTRUCTURE -- not part of the product (see oracle.h).
//
// "OpenFOAM-native-equivalent" CPU baseline (CPU baseline iii):
// what OpenFOAM itself would run for the same system without OGL -- the face-based
// lduMatrix::Amul and the PCG solver with a diagonal preconditioner -- restated from
// OpenFOAM's published algorithm (OpenFOAM is not in /root/reference and not installed:
// src/OpenFOAM/matrices/lduMatrix/lduMatrix/lduMatrixATmul.C `Amul`, `sumA`;
// solvers/PCG/PCG.C and solvers/PBiCGStab/PBiCGStab.C `scalarSolve`; lduMatrixSolver.C `normFactor`;
// preconditioners/diagonalPreconditioner).  Parity unpinned: no OpenFOAM golden
// vectors exist here; it is cross-checked against the Ginkgo-order CG of krylov.cpp
// (same Krylov method, different operation order) in tests/test_oracle_solvers.py.
//
// Single rank; cyclic (local) interfaces are applied the way
// lduMatrix::updateMatrixInterfaces does for a cyclic patch:
//   result[faceCells[i]] -= bouCoeffs[i] * psi[neighbour faceCells[i]].
#include <chrono>
#include <cmath>
#include <vector>

#include "oracle.h"

namespace {

struct Ldu {
    orc_label n, nf;
    const orc_label *l, *u;
    const orc_scalar *diag, *upper, *lower;   // lower == upper when symmetric
    orc_label n_if;
    const orc_label *if_rows, *if_cols;
    const orc_scalar *if_bou;
};

// lduMatrix::Amul
void amul(const Ldu &a, const orc_scalar *psi, orc_scalar *apsi)
{
    for (orc_label c = 0; c < a.n; ++c) apsi[c] = a.diag[c] * psi[c];
    for (orc_label f = 0; f < a.nf; ++f) {
        apsi[a.u[f]] += a.lower[f] * psi[a.l[f]];
        apsi[a.l[f]] += a.upper[f] * psi[a.u[f]];
    }
    for (orc_label i = 0; i < a.n_if; ++i) apsi[a.if_rows[i]] -= a.if_bou[i] * psi[a.if_cols[i]];
}

// lduMatrix::sumA
void sum_a(const Ldu &a, orc_scalar *s)
{
    for (orc_label c = 0; c < a.n; ++c) s[c] = a.diag[c];
    for (orc_label f = 0; f < a.nf; ++f) {
        s[a.u[f]] += a.lower[f];
        s[a.l[f]] += a.upper[f];
    }
    for (orc_label i = 0; i < a.n_if; ++i) s[a.if_rows[i]] -= a.if_bou[i];
}

}  // namespace

extern "C" int orc_foam_pcg(orc_label n, orc_label n_faces, const orc_label *lower_addr,
                            const orc_label *upper_addr, const orc_scalar *diag,
                            const orc_scalar *upper, const orc_scalar *lower, orc_label n_if,
                            const orc_label *if_rows, const orc_label *if_cols,
                            const orc_scalar *if_bou, const orc_scalar *source, orc_scalar *psi,
                            orc_scalar tolerance, orc_scalar rel_tol, orc_label min_iter,
                            orc_label max_iter, orc_solve_result *result, orc_scalar *history,
                            orc_label history_cap)
{
    if (n < 0 || n_faces < 0 || !result) return 1;
    const Ldu a{n, n_faces, lower_addr, upper_addr, diag, upper, lower ? lower : upper,
                n_if, if_rows, if_cols, if_bou};
    const orc_scalar small = 1e-20, great = 1e20;   // solverPerformance::small_, great_
    std::vector<orc_scalar> pA(n), wA(n), rA(n), rD(n);
    orc_scalar wArA = great, wArAold = wArA;
    const auto t0 = std::chrono::steady_clock::now();

    amul(a, psi, wA.data());                                    // --- Calculate A.psi
    for (orc_label c = 0; c < n; ++c) rA[c] = source[c] - wA[c]; // --- initial residual field
    // --- normalisation factor (lduMatrix::solver::normFactor)
    sum_a(a, pA.data());
    orc_scalar avg = 0;
    for (orc_label c = 0; c < n; ++c) avg += psi[c];
    avg = n > 0 ? avg / n : 0;                                   // gAverage(psi)
    orc_scalar nf = 0;
    for (orc_label c = 0; c < n; ++c) {
        const orc_scalar ref = pA[c] * avg;
        nf += std::fabs(wA[c] - ref) + std::fabs(source[c] - ref);
    }
    nf += small;
    auto sum_mag = [&](const std::vector<orc_scalar> &v) {
        orc_scalar s = 0;
        for (orc_label c = 0; c < n; ++c) s += std::fabs(v[c]);
        return s;
    };
    orc_scalar init_res = sum_mag(rA) / nf, final_res = init_res;
    orc_label n_hist = 0;
    if (history && history_cap > 0) history[n_hist++] = init_res;
    auto converged = [&]() {   // solverPerformance::checkConvergence
        return final_res < tolerance || (rel_tol > small && final_res < rel_tol * init_res);
    };
    orc_label it = 0;
    if (min_iter > 0 || !converged()) {
        for (orc_label c = 0; c < n; ++c) rD[c] = 1.0 / diag[c];   // diagonalPreconditioner
        do {
            wArAold = wArA;
            for (orc_label c = 0; c < n; ++c) wA[c] = rD[c] * rA[c];   // precondition
            wArA = 0;
            for (orc_label c = 0; c < n; ++c) wArA += wA[c] * rA[c];   // gSumProd
            if (it == 0) {
                for (orc_label c = 0; c < n; ++c) pA[c] = wA[c];
            } else {
                const orc_scalar beta = wArA / wArAold;
                for (orc_label c = 0; c < n; ++c) pA[c] = wA[c] + beta * pA[c];
            }
            amul(a, pA.data(), wA.data());
            orc_scalar wApA = 0;
            for (orc_label c = 0; c < n; ++c) wApA += wA[c] * pA[c];
            if (std::fabs(wApA) / nf < small) break;               // checkSingularity
            const orc_scalar alpha = wArA / wApA;
            for (orc_label c = 0; c < n; ++c) {
                psi[c] += alpha * pA[c];
                rA[c] -= alpha * wA[c];
            }
            final_res = sum_mag(rA) / nf;
            if (history && n_hist < history_cap) history[n_hist++] = final_res;
        } while ((++it < max_iter && !converged()) || it < min_iter);
    }
    result->init_residual = init_res;
    result->final_residual = final_res;
    result->criterion_calls = it + 1;
    result->n_iterations = it;
    result->norm_factor = nf;
    result->n_history = n_hist;
    result->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

// PBiCGStab::scalarSolve with the diagonal preconditioner (asymmetric lduMatrix).
// history[k]: residual at every convergence test -- entry 0 the initial one, then per
// iteration the one on s and the one on r (the same two calls per iteration that OGL's
// criterion sees inside Ginkgo's BiCGStab).
extern "C" int orc_foam_pbicgstab(orc_label n, orc_label n_faces, const orc_label *lower_addr,
                                  const orc_label *upper_addr, const orc_scalar *diag,
                                  const orc_scalar *upper, const orc_scalar *lower, orc_label n_if,
                                  const orc_label *if_rows, const orc_label *if_cols,
                                  const orc_scalar *if_bou, const orc_scalar *source,
                                  orc_scalar *psi, orc_scalar tolerance, orc_scalar rel_tol,
                                  orc_label min_iter, orc_label max_iter, orc_solve_result *result,
                                  orc_scalar *history, orc_label history_cap)
{
    if (n < 0 || n_faces < 0 || !result) return 1;
    const Ldu a{n, n_faces, lower_addr, upper_addr, diag, upper, lower ? lower : upper,
                n_if, if_rows, if_cols, if_bou};
    const orc_scalar small = 1e-20;
    std::vector<orc_scalar> pA(n), yA(n), rA(n), rD(n), AyA(n), sA(n), zA(n), tA(n), rA0(n);
    const auto t0 = std::chrono::steady_clock::now();
    amul(a, psi, yA.data());
    for (orc_label c = 0; c < n; ++c) rA[c] = source[c] - yA[c];
    sum_a(a, pA.data());
    orc_scalar avg = 0;
    for (orc_label c = 0; c < n; ++c) avg += psi[c];
    avg = n > 0 ? avg / n : 0;
    orc_scalar nf = 0;
    for (orc_label c = 0; c < n; ++c) {
        const orc_scalar ref = pA[c] * avg;
        nf += std::fabs(yA[c] - ref) + std::fabs(source[c] - ref);
    }
    nf += small;
    auto sum_mag = [&](const std::vector<orc_scalar> &v) {
        orc_scalar s = 0;
        for (orc_label c = 0; c < n; ++c) s += std::fabs(v[c]);
        return s;
    };
    auto dot = [&](const std::vector<orc_scalar> &x, const std::vector<orc_scalar> &y) {
        orc_scalar s = 0;
        for (orc_label c = 0; c < n; ++c) s += x[c] * y[c];
        return s;
    };
    orc_scalar init_res = sum_mag(rA) / nf, final_res = init_res;
    orc_label n_hist = 0, calls = 1;
    if (history && history_cap > 0) history[n_hist++] = init_res;
    auto converged = [&]() {
        return final_res < tolerance || (rel_tol > small && final_res < rel_tol * init_res);
    };
    auto record = [&]() {
        ++calls;
        if (history && n_hist < history_cap) history[n_hist++] = final_res;
    };
    orc_label it = 0;
    bool stopped_on_s = false;
    if (min_iter > 0 || !converged()) {
        rA0 = rA;
        for (orc_label c = 0; c < n; ++c) rD[c] = 1.0 / diag[c];
        orc_scalar rA0rA = 0, alpha = 0, omega = 0;
        do {
            const orc_scalar rA0rAold = rA0rA;
            rA0rA = dot(rA0, rA);
            if (std::fabs(rA0rA) < small) break;            // checkSingularity
            if (it == 0) {
                pA = rA;
            } else {
                if (std::fabs(omega) < small) break;
                const orc_scalar beta = (rA0rA / rA0rAold) * (alpha / omega);
                for (orc_label c = 0; c < n; ++c) pA[c] = rA[c] + beta * (pA[c] - omega * AyA[c]);
            }
            for (orc_label c = 0; c < n; ++c) yA[c] = rD[c] * pA[c];
            amul(a, yA.data(), AyA.data());
            const orc_scalar rA0AyA = dot(rA0, AyA);
            alpha = rA0rA / rA0AyA;
            for (orc_label c = 0; c < n; ++c) sA[c] = rA[c] - alpha * AyA[c];
            final_res = sum_mag(sA) / nf;
            record();
            if (converged()) {
                for (orc_label c = 0; c < n; ++c) psi[c] += alpha * yA[c];
                ++it;
                stopped_on_s = true;
                break;
            }
            for (orc_label c = 0; c < n; ++c) zA[c] = rD[c] * sA[c];
            amul(a, zA.data(), tA.data());
            const orc_scalar tAtA = dot(tA, tA);
            omega = dot(tA, sA) / tAtA;
            for (orc_label c = 0; c < n; ++c) {
                psi[c] += alpha * yA[c] + omega * zA[c];
                rA[c] = sA[c] - omega * tA[c];
            }
            final_res = sum_mag(rA) / nf;
            record();
        } while ((++it < max_iter && !converged()) || it < min_iter);
    }
    (void)stopped_on_s;
    result->init_residual = init_res;
    result->final_residual = final_res;
    result->criterion_calls = calls;
    result->n_iterations = it;
    result->norm_factor = nf;
    result->n_history = n_hist;
    result->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
