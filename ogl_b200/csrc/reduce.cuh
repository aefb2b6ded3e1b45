// Single-pass, run-to-run deterministic grid reductions and the device-side
// scalar epilogues (Krylov coefficients + OGL's stopping criterion).
//
// Every kernel that needs a global sum (dot, L1 norm) reduces in registers ->
// warp shuffles -> shared memory, stores one partial per block, takes a ticket,
// and the LAST block to finish adds the partials in a fixed order.  The result
// therefore does not depend on block scheduling, so iteration counts do not
// jitter between runs.  On one rank the last block then runs the scalar
// epilogue in place; on several ranks the partial sums are all-reduced first
// (comm.cu) and a one-thread kernel runs the same epilogue.
#pragma once

#include "common.cuh"

namespace ogl {

constexpr double kSmall = 1e-15;  // OpenFOAM `SMALL` (double precision)

enum Epilogue : int {
    EPI_NONE = 0,
    EPI_MEAN_LOCAL,      // red[0] = local sum(x)  ->  weighted local mean
    EPI_INIT_CHECK,      // red = {<r,z>, |r|_1, normFactor sum}: first criterion call
    EPI_CG_BETA,         // red = {<p,q>}
    EPI_CG_RHO_CHECK,    // red = {<r,z>, |r|_1}
    EPI_BICG_INIT_CHECK, // red = {<rr,r>, |r|_1, normFactor sum}
    EPI_BICG_ALPHA,      // red = {<rr,v>}
    EPI_BICG_CHECK_S,    // red = {-, |s|_1}
    EPI_BICG_OMEGA,      // red = {<s,t>, <t,t>}
    EPI_BICG_RHO_CHECK,  // red = {<rr,r>, |r|_1}
    EPI_GMRES_INIT,      // red = {<r,r>, |r|_1, normFactor sum}
    EPI_GMRES_RESTART,   // red = {<r,r>, |r|_1}
};

// StoppingCriterion.C:71-151, evaluated by one thread on the device.
// `norm1` is the (already globally summed) L1 norm of the residual handed to
// the criterion.  Returns true when the solver must stop.
__device__ __forceinline__ bool criterion_check(SolveState *s, double norm1,
                                                double *history)
{
    const int it = s->iter;
    if (it > 0 && it < s->min_iter) {  // :77-81
        s->iter = it + 1;
        return false;
    }
    if (it % s->frequency != 0) {      // :84-87
        s->iter = it + 1;
        return false;
    }
    double rn = norm1;
    if (it == 0) s->init_res = rn / s->norm_factor;   // :102-111
    rn /= s->norm_factor;                             // :113
    if (s->export_res && history && it < s->history_cap) {
        history[it] = rn;                             // :115-117
        s->n_history = it + 1;
    }
    s->res = rn;                                      // :119
    bool stop = false;
    if (it >= s->max_iter) stop = true;               // :124-126
    if (rn < s->tolerance) stop = true;               // :128-130
    if (s->rel_tol > 0 && rn < s->rel_tol * s->init_res) stop = true;  // :132-136
    s->iter = it + 1;                                 // :143
    return stop;
}

// extra scalars some epilogues need
struct EpiArgs {
    double inv_n_local;   // 1 / n_local
    double weight;        // n_local / n_global
    double *history;
};

EpiArgs make_epi_args(Context *ctx);

__device__ __forceinline__ void run_epilogue(int epi, SolveState *s, const EpiArgs &a)
{
    switch (epi) {
    case EPI_MEAN_LOCAL:
        // distributed::Vector::compute_mean: local mean times local/global weight
        s->red[0] = (s->red[0] * a.inv_n_local) * a.weight;
        break;
    case EPI_INIT_CHECK:
    case EPI_BICG_INIT_CHECK:
        s->norm_factor = s->red[2] + kSmall;          // StoppingCriterion.C:68
        // fallthrough into the regular rho/check epilogue
    case EPI_CG_RHO_CHECK:
    case EPI_BICG_RHO_CHECK: {
        // `swap(prev_rho, rho)` at the end of the previous iteration followed by
        // the new dot product (Ginkgo cg.cpp / bicgstab.cpp)
        s->prev_rho = s->rho;
        s->rho = s->red[0];
        if (criterion_check(s, s->red[1], a.history)) {
            s->done = 1;
            break;
        }
        if (epi == EPI_CG_RHO_CHECK || epi == EPI_INIT_CHECK) {
            // cg::step_1 : p = z + (rho / prev_rho) p, p = z if prev_rho == 0
            s->flag_p_is_z = (s->prev_rho == 0.0);
            s->coef_p = s->flag_p_is_z ? 0.0 : s->rho / s->prev_rho;
        } else {
            // bicgstab::step_1 : p = r + (rho/prev_rho * alpha/omega)(p - omega v)
            s->flag_p_is_z = !(s->prev_rho * s->omega != 0.0);
            s->coef_p = s->flag_p_is_z ? 0.0 : s->rho / s->prev_rho * s->alpha / s->omega;
        }
        break;
    }
    case EPI_CG_BETA:
        // cg::step_2 : if beta != 0 { t = rho / beta; x += t p; r -= t q }
        s->beta = s->red[0];
        s->coef_x = (s->beta != 0.0) ? s->rho / s->beta : 0.0;
        break;
    case EPI_BICG_ALPHA:
        // bicgstab::step_2 : alpha = rho / beta (0 if beta == 0)
        s->beta = s->red[0];
        s->alpha = (s->beta != 0.0) ? s->rho / s->beta : 0.0;
        break;
    case EPI_BICG_CHECK_S:
        if (criterion_check(s, s->red[1], a.history)) {
            s->done = 1;
            s->stop_half = 1;   // bicgstab::finalize still owes x += alpha y
        }
        break;
    case EPI_BICG_OMEGA:
        // bicgstab::step_3 : omega = gamma / beta (0 if beta == 0)
        s->gamma = s->red[0];
        s->beta = s->red[1];
        s->omega = (s->beta != 0.0) ? s->gamma / s->beta : 0.0;
        break;
    case EPI_GMRES_INIT:
        s->norm_factor = s->red[2] + kSmall;
        // fallthrough
    case EPI_GMRES_RESTART:
        s->res_norm2 = sqrt(s->red[0]);
        // red[1] (|r|_1 of the restart residual) is kept: the criterion sees
        // this vector until the next restart
        break;
    default:
        break;
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of NRED values (all threads must call).  Result valid in
// thread 0.  `sm` needs NRED * 32 doubles.
template <int NRED>
__device__ __forceinline__ void block_sum(double (&v)[NRED], double *sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int j = 0; j < NRED; ++j) {
        v[j] = warp_sum(v[j]);
        if (lane == 0) sm[j * 32 + warp] = v[j];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j) {
            double t = (lane < nwarps) ? sm[j * 32 + lane] : 0.0;
            v[j] = warp_sum(t);
        }
    }
}

// Grid-wide deterministic reduction + optional inline epilogue.
// All threads of all blocks call this exactly once per kernel.
// red_base: first slot of state->red the sums go to.
template <int NRED>
__device__ __forceinline__ void grid_reduce(double (&v)[NRED], double *partials,
                                            unsigned int *ticket, SolveState *state,
                                            int red_base, int epi, bool inline_epi,
                                            const EpiArgs &ea, bool accumulate = false)
{
    __shared__ double sm[NRED * 32];
    __shared__ bool is_last;
    block_sum<NRED>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j) partials[(size_t)blockIdx.x * NRED + j] = v[j];
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc[NRED];
#pragma unroll
    for (int j = 0; j < NRED; ++j) acc[j] = 0.0;
#pragma unroll 4
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int j = 0; j < NRED; ++j)
            acc[j] += __ldcg(&partials[(size_t)b * NRED + j]);
    }
    __syncthreads();   // sm reuse
    block_sum<NRED>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j) {
            // accumulate: corrections of the non-local block on top of the
            // local kernel's sums (multi-rank SpMV with fused reductions)
            state->red[red_base + j] = accumulate ? state->red[red_base + j] + acc[j] : acc[j];
        }
        *ticket = 0u;
        if (inline_epi && epi != EPI_NONE) run_epilogue(epi, state, ea);
    }
}

}  // namespace ogl
