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

#include <cstddef>

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
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool criterion_check_untimed(SolveState *s, double norm1, double *history);

// timed wrapper: an evaluated call records its own duration (skipped calls cost nothing and
// leave the record alone, like the early returns of StoppingCriterion.C:77-87)
__device__ __forceinline__ bool criterion_check(SolveState *s, double norm1, double *history)
{
    const int it = s->iter;
    const bool evaluated = !(it > 0 && it < s->min_iter) && (it % s->frequency == 0);
    const unsigned long long t0 = evaluated ? globaltimer_ns() : 0ull;
    const bool stop = criterion_check_untimed(s, norm1, history);
    if (evaluated) {
        const unsigned long long dt = globaltimer_ns() - t0;
        s->crit_ns = dt > 0 ? dt : 1ull;   // globaltimer ticks in steps of up to 1 us on some parts
    }
    return stop;
}

__device__ __forceinline__ bool criterion_check_untimed(SolveState *s, double norm1,
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

// ---- peer-memory communication window (multi-GPU, one process per GPU) --------
// Every rank exposes one device allocation ("window") to its peers through CUDA
// IPC over NVLink/NVSwitch.  Kernels write straight into the peers' windows:
//   * all-reduce: the last block of a reducing kernel stores its partial sums
//     into slot[my rank] of EVERY rank's mailbox, waits until all slots of its
//     own mailbox carry the current sequence stamp and adds them in rank order
//     (identical order on every rank => bit-identical, deterministic results);
//   * halo exchange: the pack kernel stores x[send_idxs] directly into the
//     neighbours' receive buffers and raises a per-neighbour flag; the kernel
//     that consumes the halo waits for the flags and acknowledges afterwards.
// Buffers are double-buffered by sequence parity; acknowledgements give flow
// control for back-to-back exchanges.  No NCCL call and no extra kernel launch
// is needed per iteration, so a chunk of iterations replays as one CUDA graph.
constexpr int kMaxPeers = 8;       // ranks per NVSwitch domain
constexpr int kMaxTargets = 32;    // neighbour ranks of one rank
constexpr int kSlot = 2 * kMaxReduce;   // 8-byte words per source rank: (low half | seq), (high half | seq) per value
constexpr long long kSpinCycles = 40000000000LL;   // ~20 s (ranks time-slicing ONE device wait that long for each other): fail loudly instead of hanging

struct CommDev {
    int rank, n_ranks, n_targets, pad;
    // all-reduce mailboxes: mbox[r] = rank r's mailbox base, [2][n_ranks][kSlot] 8-byte words
    double *mbox[kMaxPeers];
    unsigned long long ar_seq;          // device-side sequence counters
    unsigned long long halo_seq;
    unsigned int pack_ticket, nl_ticket;
    // halo exchange, per neighbour t
    int send_offs[kMaxTargets + 1];
    double *peer_recv[kMaxTargets];           // neighbour's recv block for me, parity 0
    long long peer_recv_stride[kMaxTargets];  // doubles between the neighbour's two parity buffers
    unsigned long long *peer_data_flag[kMaxTargets];  // neighbour's data flag for me
    unsigned long long *peer_ack_flag[kMaxTargets];   // neighbour's ack flag for me
    unsigned long long *my_data_flag;         // [n_targets] raised by neighbours
    unsigned long long *my_ack_flag;          // [n_targets] raised by neighbours
    double *my_recv;                          // [2][n_halo]
    long long my_recv_stride;
};

// extra scalars some epilogues need
struct EpiArgs {
    double inv_n_local;   // 1 / n_local
    double weight;        // n_local / n_global
    double *history;
    CommDev *comm;        // peer-memory window (nullptr: single rank or NCCL path)
    int ar_count;         // leading state->red[] slots to all-reduce in this launch (0: none)
    int ar_after_epi;     // 1: run the epilogue on the local sums first (mean)
    // device-side timeline (option `trace`): [0] = cursor, then (tag << 48 | globaltimer ns)
    unsigned long long *trace;
    int trace_tag, trace_cap;
};

// one timeline event; called by single threads at a handful of points per launch
__device__ __forceinline__ void trace_event(const EpiArgs &ea, int sub)
{
    if (!ea.trace || ea.trace_tag == 0) return;   // untagged launches (prologue, other solvers) stay silent
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned long long i = atomicAdd(ea.trace, 1ull);
    if (i < (unsigned long long)ea.trace_cap)
        ea.trace[1 + i] = ((unsigned long long)(ea.trace_tag + sub) << 48) | (t & 0xffffffffffffull);
}

EpiArgs make_epi_args(Context *ctx, int ar_count = 0, bool ar_after_epi = false);

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

__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p >= want; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long *p, unsigned long long want)
{
    // relaxed polls (an acquire load per poll would invalidate the SM's L1 each time),
    // one acquire fence once the flag is up
    const long long t0 = clock64();
    unsigned long long v;
    while (true) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= want) break;
        if (clock64() - t0 > kSpinCycles) return false;
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return true;
}

// A double that validates itself: two 8-byte words, each 32 payload bits under a
// 32-bit sequence stamp, written with one 16-byte store.  The reader polls until
// both words carry the stamp it expects -- no flag, no fence, one NVLink traversal.
__device__ __forceinline__ void push_stamped(unsigned long long *dst, double v, unsigned long long stamp)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long lo = (bits & 0xffffffffull) | stamp, hi = (bits >> 32) | stamp;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
// false on timeout (t0: clock64() when the caller started waiting)
__device__ __forceinline__ bool pull_stamped(const unsigned long long *src, unsigned long long stamp,
                                             long long t0, double &v)
{
    unsigned long long lo, hi;
    while (true) {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
        if ((lo & 0xffffffff00000000ull) == stamp && (hi & 0xffffffff00000000ull) == stamp) break;
        if (clock64() - t0 > kSpinCycles) return false;
    }
    v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return true;
}
// stamp of the all-reduce number `seq`
__device__ __forceinline__ unsigned long long stamp_of(unsigned long long seq) { return (seq & 0xffffffffull) << 32; }
__device__ __forceinline__ unsigned long long ld_ar_seq(const CommDev *c)
{
    unsigned long long q;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(q) : "l"(&c->ar_seq) : "memory");
    return q;
}

// All-reduce (sum) of state->red[0..count) over the ranks through the peers'
// mailboxes.  Called by ALL threads of ONE block per rank (the last block of a
// reducing kernel); s->red must already hold the local sums (s may be a
// shared-memory copy of the state).
//
// Low-latency protocol: every value travels as two self-validating 8-byte words
// (32 payload bits | 32-bit sequence stamp), so a contribution is ONE NVLink
// traversal -- no flag behind the data and hence no release fence that would
// wait for the data to be acknowledged first.  The receiver polls the words
// until both carry the current stamp.  Mailboxes are double-buffered by the
// parity of the sequence number; a slot is overwritten every second all-reduce,
// so a stale stamp can never match.  Sums are formed in rank order on every
// rank: bit-identical results everywhere.  Boundary values pushed into peer
// windows (CG ghost-p mode) use the same self-validating words, stamped with the
// number of the all-reduce that follows them, so nothing here has to publish.
__device__ __forceinline__ void p2p_allreduce(SolveState *s, int count, CommDev *c)
{
    __shared__ unsigned long long seq_sh;
    __shared__ double vals_sh[kMaxPeers][kMaxReduce];
    if (threadIdx.x == 0) {
        const unsigned long long q = ld_ar_seq(c) + 1;
        c->ar_seq = q;
        seq_sh = q;
    }
    __syncthreads();
    const unsigned long long seq = seq_sh;
    const unsigned long long stamp = stamp_of(seq);
    const int parity = (int)(seq & 1ull);
    const int t = threadIdx.x;
    const int nw = count > 0 ? count : 1;   // count 0: pure rendezvous
    if (t < c->n_ranks) {
        // my partial sums -> slot[my rank] of rank t's mailbox
        unsigned long long *box = reinterpret_cast<unsigned long long *>(c->mbox[t]) +
                                  ((size_t)(parity * c->n_ranks + c->rank)) * kSlot;
        for (int j = 0; j < nw; ++j) push_stamped(box + 2 * j, s->red[j], stamp);
        // rank t's contribution in my own mailbox
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(c->mbox[c->rank]) +
                                         ((size_t)(parity * c->n_ranks + t)) * kSlot;
        const long long t0 = clock64();
        for (int j = 0; j < nw; ++j) {
            double v = 0.0;
            if (!pull_stamped(mine + 2 * j, stamp, t0, v)) s->comm_error = 1;
            vals_sh[t][j] = v;
        }
    }
    __syncthreads();
    if (t == 0) {
        for (int j = 0; j < count; ++j) {
            double acc = 0.0;
            for (int r = 0; r < c->n_ranks; ++r) acc += vals_sh[r][j];   // rank order: the same on every rank
            s->red[j] = acc;
        }
        if (s->comm_error) s->done = 1;
    }
    __syncthreads();
}

// Grid-wide deterministic reduction + optional inline epilogue.
// All threads of all blocks call this exactly once per kernel.
// red_base: first slot of state->red the sums go to.
//
// The last block to arrive (one acq_rel ticket per block: release of the
// block's partials and acquire of everybody else's in one operation) adds the
// partials in block order and then does ALL the scalar work -- all-reduce over
// the ranks, Krylov coefficients, stopping criterion -- on a shared-memory copy
// of the SolveState, written back in one coalesced pass.  One thread chasing
// the state's fields through L2 one dependent load at a time cost ~2 us per
// reduction on the device timeline (tools/trace_iter.py).
constexpr int kStateWords = (int)(sizeof(SolveState) / sizeof(double));
static_assert(sizeof(SolveState) % sizeof(double) == 0, "SolveState is copied in 8-byte words");
constexpr int kCommErrWord = (int)(offsetof(SolveState, comm_error) / sizeof(double));

// Returns 0 in every block but the last one to arrive; there 1, or 2 when the
// state says the solve is over (criterion fired / peer timeout) after the epilogue.
template <int NRED>
__device__ __forceinline__ int grid_reduce(double (&v)[NRED], double *partials,
                                            unsigned int *ticket, SolveState *state,
                                            int red_base, int epi, bool inline_epi,
                                            const EpiArgs &ea, bool accumulate = false)
{
    __shared__ double sm[NRED * 32];
    __shared__ double st_sh[kStateWords];
    __shared__ bool is_last;
    block_sum<NRED>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j) partials[(size_t)blockIdx.x * NRED + j] = v[j];
        unsigned int t;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(t) : "l"(ticket), "r"(1u) : "memory");
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return 0;
    if (threadIdx.x == 0) trace_event(ea, 1);   // last CTA of the grid has arrived
    for (int w = threadIdx.x; w < kStateWords; w += blockDim.x) {
        double d;
        asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(d) : "l"(reinterpret_cast<const double *>(state) + w) : "memory");
        st_sh[w] = d;
    }
    double acc[NRED];
#pragma unroll
    for (int j = 0; j < NRED; ++j) acc[j] = 0.0;
#pragma unroll 4
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int j = 0; j < NRED; ++j)
            acc[j] += __ldcg(&partials[(size_t)b * NRED + j]);
    }
    __syncthreads();   // sm reuse; st_sh complete
    block_sum<NRED>(acc, sm);
    SolveState *s = reinterpret_cast<SolveState *>(st_sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j) {
            // accumulate: corrections of the non-local block on top of the
            // local kernel's sums (multi-rank SpMV with fused reductions)
            s->red[red_base + j] = accumulate ? s->red[red_base + j] + acc[j] : acc[j];
        }
        *ticket = 0u;
        trace_event(ea, 2);   // local sums done
    }
    const bool ar = ea.comm != nullptr && ea.ar_count > 0;   // block-uniform
    if (ar && !ea.ar_after_epi) {
        __syncthreads();
        p2p_allreduce(s, ea.ar_count, ea.comm);
        if (threadIdx.x == 0) trace_event(ea, 3);   // all-reduced
    }
    if (threadIdx.x == 0 && inline_epi && epi != EPI_NONE) run_epilogue(epi, s, ea);
    if (threadIdx.x == 0) trace_event(ea, 4);   // epilogue done
    if (ar && ea.ar_after_epi) {
        __syncthreads();
        p2p_allreduce(s, ea.ar_count, ea.comm);
    }
    __syncthreads();
    // write the state back; the comm_error word only when it was raised here
    // (waiting CTAs of a halo kernel may be raising it in global memory)
    for (int w = threadIdx.x; w < kStateWords; w += blockDim.x)
        if (w != kCommErrWord) reinterpret_cast<double *>(state)[w] = st_sh[w];
    if (threadIdx.x == 0 && s->comm_error) state->comm_error = 1;
    return s->done ? 2 : 1;
}

}  // namespace ogl
