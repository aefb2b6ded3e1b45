// Krylov iterations as hand-written sm_100a kernels.
//
// Replaces gko::solver::{Cg,Bicgstab}::apply (GMRES: gmres.cu) driven by OGL's
// OpenFOAMDistStoppingCriterion (StoppingCriterion/StoppingCriterion.C:11-151).
// Operation order follows Ginkgo's core/solver/{cg,bicgstab}.cpp; each vector
// update is fused with the preconditioner apply (scalar Jacobi / none) and with
// the reductions the next step needs, every coefficient stays on the device
// (SolveState).  CG runs as the body of a CUDA-graph WHILE node that the
// criterion epilogue ends on the device (run_chunks); the other solvers replay
// chunk graphs while the host polls a `done` flag -- never a per-iteration
// D2H + sync as in StoppingCriterion.C:95-97.
//
// PCG iteration, scalar Jacobi (3 launches on 1..8 GPUs, ~192 n bytes for 7-pt):
//   k_cg_p    p' = z + (rho/prev_rho) p (other buffer; + ghost entries)  24 n
//   spmv      q = A p', beta = <p',q>, alpha = rho/beta     12 nnz + 4 n + 16 n (+8 n p)
//   k_cg_xr   x += alpha p'; r' = r - alpha q (other buffer); z = r'/diag;
//             rho = <r',z>; |r'|_1; criterion (+ push of the boundary z)  64 n
// Small systems: the same loop as one persistent kernel (pcg_fused.cu).
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

int precond_apply(Context *ctx, const double *r, double *z, const double *dot_with,
                  int red_base, bool guard_done, int epi, bool inline_epi, int ar_count);
int solve_gmres(Context *ctx, const ogl_solve_params *p, ogl_solve_result *res);

namespace {

constexpr int kT = kBlas1Threads;

struct VecK {
    label n;
    SolveState *state;
    double *partials;
    unsigned int *ticket;
    int epi, inline_epi, guard_done;
    EpiArgs ea;
    const double *in0, *in1, *in2, *in3, *in4, *in5;
    double *out0, *out1, *out2, *out3;
    const label *send_idx;   // fused halo pack (multi-GPU peer-memory path)
    label n_send;
    int pack;                // k_cg_p: 1 store boundary p', 2 update ghost p; k_cg_xr: 1 push boundary z
    double *const *push_dst;   // k_cg_xr: destination of every send entry (slot 2 of the peer windows)
    cudaGraphConditionalHandle cond;   // k_cg_xr inside a WHILE-node body: cleared when the solve is done
    // block Jacobi (k_cg_xr_bj): block of a row, block row ranges, offsets of the inverse blocks
    const label *bj_row_block, *bj_block_ptrs;
    const int64_t *bj_block_offs;
    const double *bj_inv;
};

#define GRID_STRIDE(i, n)                                                             \
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n);         \
         i += (int64_t)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_sum(const VecK a)
{
    double red[1] = {0.0};
    GRID_STRIDE(i, a.n) red[0] = __dadd_rn(red[0], a.in0[i]);
    grid_reduce<1>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
}

// out0[i] = state->red[0]
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_fill_from_red(const VecK a)
{
    const double v = a.state->red[0];
    GRID_STRIDE(i, a.n) a.out0[i] = v;
}

__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_fill(label n, double *out, double v)
{
    GRID_STRIDE(i, n) out[i] = v;
}

__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_scale(label n, double *v, double s)
{
    GRID_STRIDE(i, n) v[i] = __dmul_rn(v[i], s);
}

// First criterion call (StoppingCriterion.C:102-111) for every solver.
//   in0 = r, in1 = b, in2 = w = A*(mean(x)*1), in3 = inv_diag
//   out0 = z (PK 1), out1 = rr (MODE 1: BiCGStab copies r)
//   red = { MODE 0: <r,z>  MODE 1/2: <r,r> ,  |r|_1 ,  normFactor sum (:32-69) }
// PK: 0 none (z == r), 1 scalar Jacobi fused, 2 deferred (block Jacobi runs after)
template <int PK, int MODE>
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_init_norms(const VecK a)
{
    double red[3] = {0.0, 0.0, 0.0};
    GRID_STRIDE(i, a.n) {
        const double r = a.in0[i];
        const double bs = __dsub_rn(a.in1[i], __dmul_rn(1.0, a.in2[i]));     // :54
        const double part2 = fabs(bs);                                       // :56
        double t = fabs(__dsub_rn(bs, __dmul_rn(1.0, r)));                   // :58-59
        t = __dadd_rn(t, __dmul_rn(1.0, part2));                             // :61
        red[2] = __dadd_rn(red[2], fabs(t));                                 // :63
        red[1] = __dadd_rn(red[1], fabs(r));
        if (MODE == 0) {
            if (PK == 1) {
                const double z = __dmul_rn(r, a.in3[i]);
                a.out0[i] = z;
                red[0] = __dadd_rn(red[0], __dmul_rn(r, z));
            } else if (PK == 0) {
                red[0] = __dadd_rn(red[0], __dmul_rn(r, r));
            }
        } else {
            if (MODE == 1) a.out1[i] = r;
            red[0] = __dadd_rn(red[0], __dmul_rn(r, r));
        }
    }
    grid_reduce<3>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
}

// cg::step_1   in0 = z (or r when unpreconditioned), in1 = p (previous), out0 = p (new)
// Out of place (two p buffers alternate), so that on several GPUs the same
// launch can ALSO compute the new p of the boundary cells and store it straight
// into the neighbours' halo buffers (peer memory): no pack kernel, and the
// stores are on their way before the SpMV starts.
// 128-bit loads/stores (two rows per thread and trip).
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_cg_p(const VecK a)
{
    cudaGridDependencySynchronize();            // PDL: everything below reads the previous kernel's output
    cudaTriggerProgrammaticLaunchCompletion();  // the SpMV may start getting resident
    // the first trip's vector loads need no scalar: in flight while `done` and the coefficient arrive
    const double2 *__restrict__ z2 = reinterpret_cast<const double2 *>(a.in0);
    const double2 *__restrict__ po2 = reinterpret_cast<const double2 *>(a.in1);
    const int64_t n2 = a.n >> 1;
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double2 z_first = make_double2(0.0, 0.0), po_first = z_first;
    if (i0 < n2) {
        z_first = z2[i0];
        po_first = po2[i0];
    }
    if (a.guard_done && a.state->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(a.ea, 0);
    const bool p_is_z = a.state->flag_p_is_z != 0;
    const double t = a.state->coef_p;
    if (a.pack == 1) {
        CommDev *c = a.ea.comm;
        const unsigned long long seq = c->halo_seq + 1;
        const int parity = (int)(seq & 1ull);
        if (seq > 2 && threadIdx.x < c->n_targets) {
            if (!wait_flag(&c->my_ack_flag[threadIdx.x], seq - 2)) a.state->comm_error = 1;
        }
        __syncthreads();
        GRID_STRIDE(k, a.n_send) {
            const label cell = a.send_idx[k];
            const double z = a.in0[cell];
            const double v = p_is_z ? z : __dadd_rn(z, __dmul_rn(t, a.in1[cell]));
            int tg = 0;
            while (k >= c->send_offs[tg + 1]) ++tg;
            double *dst = c->peer_recv[tg] + (size_t)parity * c->peer_recv_stride[tg] + (k - c->send_offs[tg]);
            *dst = v;   // published by the SpMV kernel's release on entry
        }
    }
    double2 *__restrict__ p2 = reinterpret_cast<double2 *>(a.out0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (i0 < n2) {
        double2 p = z_first;
        if (!p_is_z) {
            p.x = __dadd_rn(z_first.x, __dmul_rn(t, po_first.x));
            p.y = __dadd_rn(z_first.y, __dmul_rn(t, po_first.y));
        }
        p2[i0] = p;
    }
#pragma unroll 2
    for (int64_t i = i0 + stride; i < n2; i += stride) {
        const double2 z = z2[i];
        double2 p = z;
        if (!p_is_z) {
            const double2 po = po2[i];
            p.x = __dadd_rn(z.x, __dmul_rn(t, po.x));
            p.y = __dadd_rn(z.y, __dmul_rn(t, po.y));
        }
        p2[i] = p;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        const double z = a.in0[i];
        a.out0[i] = p_is_z ? z : __dadd_rn(z, __dmul_rn(t, a.in1[i]));
    }
    // (behind the main loop: nobody needs the ghost entries before the next launch, and the words were
    // pushed a whole kernel ago -- in front of the loop the extra round trip delayed every thread)
    if (a.pack == 2) {
        // ghost-p mode: the neighbours pushed their boundary z into slot 2 of my window
        // (k_cg_xr / push_boundary) as self-validating words stamped with the number of the
        // all-reduce that gave rho; the ghost entries of p get the same update as the
        // neighbours' own cells -- same operands, same operations, same bits.
        const CommDev *c = a.ea.comm;
        const unsigned long long *zg =
            reinterpret_cast<const unsigned long long *>(c->my_recv + 2 * (size_t)c->my_recv_stride);
        const unsigned long long stamp = stamp_of(ld_ar_seq(c));
        const long long t0 = clock64();
        GRID_STRIDE(k, a.n_send) {
            double z = 0.0;
            if (!pull_stamped(zg + 2 * k, stamp, t0, z)) a.state->comm_error = 1;
            a.out0[a.n + k] = p_is_z ? z : __dadd_rn(z, __dmul_rn(t, a.in1[a.n + k]));
        }
    }
}

// cg::step_2 + preconditioner + <r,z> + |r|_1 (+ criterion on one rank)
//   in0 = p, in1 = q, in2 = inv_diag, in3 = r ; out0 = x, out1 = r' (the other r buffer), out2 = z
template <int PK>
__device__ __forceinline__ void cg_xr_elem(bool upd, double t, double &x, double &r, double p,
                                           double q, double d, double &z, double (&red)[2])
{
    if (upd) {
        x = __dadd_rn(x, __dmul_rn(t, p));
        r = __dsub_rn(r, __dmul_rn(t, q));
    }
    red[1] = __dadd_rn(red[1], fabs(r));
    if (PK == 1) {
        z = __dmul_rn(r, d);
        red[0] = __dadd_rn(red[0], __dmul_rn(r, z));
    } else if (PK == 0) {
        red[0] = __dadd_rn(red[0], __dmul_rn(r, r));
    }
}

template <int PK>
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_cg_xr(const VecK a)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    // the first trip's vector loads need no scalar: in flight while `done`, beta and alpha arrive
    const double2 *__restrict__ p2 = reinterpret_cast<const double2 *>(a.in0);
    const double2 *__restrict__ q2 = reinterpret_cast<const double2 *>(a.in1);
    const double2 *__restrict__ d2 = reinterpret_cast<const double2 *>(a.in2);
    double2 *__restrict__ x2 = reinterpret_cast<double2 *>(a.out0);
    const double2 *__restrict__ ro2 = reinterpret_cast<const double2 *>(a.in3);
    const int64_t n2 = a.n >> 1;
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double2 r_f = make_double2(0.0, 0.0), x_f = r_f, p_f = r_f, q_f = r_f, d_f = r_f;
    if (i0 < n2) {
        r_f = ro2[i0];
        x_f = x2[i0];
        p_f = p2[i0];
        q_f = q2[i0];
        if (PK == 1) d_f = d2[i0];
    }
    if (a.guard_done && a.state->done) {
        // (also reached when the criterion fired before the loop, or in the first half of a body)
        if (a.cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(a.ea, 0);
    const bool upd = a.state->beta != 0.0;
    const double t = a.state->coef_x;
    double red[2] = {0.0, 0.0};
    if (a.pack && PK != 2) {
        // ghost-p mode: the new z (r when unpreconditioned) of the boundary cells goes into
        // slot 2 of the neighbours' windows first -- r is read from the OLD buffer, so any
        // CTA may do this without a race -- and the NVLink stores drain while the main loop
        // runs.  Every value carries the number of the all-reduce this kernel ends with.
        const unsigned long long stamp = stamp_of(ld_ar_seq(a.ea.comm) + 1);
        GRID_STRIDE(k, a.n_send) {
            const label cell = a.send_idx[k];
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.push_dst[k]);
            double r = a.in3[cell];
            if (upd) r = __dsub_rn(r, __dmul_rn(t, a.in1[cell]));
            push_stamped(dst, PK == 1 ? __dmul_rn(r, a.in2[cell]) : r, stamp);
        }
    }
    double2 *__restrict__ r2 = reinterpret_cast<double2 *>(a.out1);
    double2 *__restrict__ z2 = reinterpret_cast<double2 *>(a.out2);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (i0 < n2) {
        double2 z = make_double2(0.0, 0.0);
        cg_xr_elem<PK>(upd, t, x_f.x, r_f.x, p_f.x, q_f.x, d_f.x, z.x, red);
        cg_xr_elem<PK>(upd, t, x_f.y, r_f.y, p_f.y, q_f.y, d_f.y, z.y, red);
        if (upd) x2[i0] = x_f;
        r2[i0] = r_f;
        if (PK == 1) z2[i0] = z;
    }
#pragma unroll 2
    for (int64_t i = i0 + stride; i < n2; i += stride) {
        double2 r = ro2[i];
        double2 x = make_double2(0.0, 0.0), p = x, q = x, d = x, z = x;
        if (upd) {
            x = x2[i];
            p = p2[i];
            q = q2[i];
        }
        if (PK == 1) d = d2[i];
        cg_xr_elem<PK>(upd, t, x.x, r.x, p.x, q.x, d.x, z.x, red);
        cg_xr_elem<PK>(upd, t, x.y, r.y, p.y, q.y, d.y, z.y, red);
        if (upd) x2[i] = x;
        r2[i] = r;
        if (PK == 1) z2[i] = z;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        double x = a.out0[i], r = a.in3[i], z = 0.0;
        cg_xr_elem<PK>(upd, t, x, r, a.in0[i], a.in1[i], PK == 1 ? a.in2[i] : 0.0, z, red);
        if (upd) a.out0[i] = x;
        a.out1[i] = r;
        if (PK == 1) a.out2[i] = z;
    }
    const int last = grid_reduce<2>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
    // device-side loop (CUDA-graph WHILE node): the criterion ends it
    if (a.cond && last == 2 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
}

// cg::step_2 fused with the BLOCK-Jacobi apply (maxBlockSize > 1) + <r,z> + |r|_1 + criterion:
//   z_i = sum_j Minv[i][j] r'_j over the rows j of i's block.  The block-mates' r' = r - alpha q are
//   recomputed from the OLD r buffer and q (adjacent rows: L1 hits) -- the same operation on the same
//   operands as the stored r', so z has the bits of the separate apply kernel -- and r' goes to the
//   other r buffer, so no thread reads what another one writes.  One pass instead of two, and the
//   iteration keeps its in-kernel criterion (device-side loop).
//   in0 = p, in1 = q, in3 = r ; out0 = x, out1 = r', out2 = z
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_cg_xr_bj(const VecK a)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (a.guard_done && a.state->done) {
        if (a.cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(a.ea, 0);
    const bool upd = a.state->beta != 0.0;
    const double t = a.state->coef_x;
    double red[2] = {0.0, 0.0};
    GRID_STRIDE(i, a.n) {
        double r = a.in3[i];
        if (upd) {
            a.out0[i] = __dadd_rn(a.out0[i], __dmul_rn(t, a.in0[i]));
            r = __dsub_rn(r, __dmul_rn(t, a.in1[i]));
        }
        a.out1[i] = r;
        red[1] = __dadd_rn(red[1], fabs(r));
        const label b = a.bj_row_block[i];
        const label lo = a.bj_block_ptrs[b], sz = a.bj_block_ptrs[b + 1] - lo;
        const double *m = a.bj_inv + a.bj_block_offs[b] + (int64_t)(i - lo) * sz;
        double z = 0.0;
        for (label j = 0; j < sz; ++j) {
            double rj = a.in3[lo + j];
            if (upd) rj = __dsub_rn(rj, __dmul_rn(t, a.in1[lo + j]));
            z = __dadd_rn(z, __dmul_rn(m[j], rj));
        }
        a.out2[i] = z;
        red[0] = __dadd_rn(red[0], __dmul_rn(r, z));
    }
    const int last = grid_reduce<2>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
    if (a.cond && last == 2 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
}

// The same for the common case of UNIFORM blocks (every block has exactly BS rows, the last one
// possibly fewer: what find_blocks + agglomeration give for a mesh-ordered stencil matrix): block and
// inverse-row addresses follow from the row index, so all 3 BS loads of a row are independent and
// issued together instead of chasing row -> block -> offset -> entries.
template <int BS>
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_cg_xr_bj_uniform(const VecK a)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (a.guard_done && a.state->done) {
        if (a.cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(a.ea, 0);
    const bool upd = a.state->beta != 0.0;
    const double t = a.state->coef_x;
    double red[2] = {0.0, 0.0};
    const int64_t n_full = (int64_t)a.n / BS * BS;   // rows in complete blocks
    GRID_STRIDE(i, a.n) {
        const int64_t lo = i / BS * BS;
        const int li = (int)(i - lo);
        const int sz = lo < n_full ? BS : (int)(a.n - lo);
        // inverse blocks are stored block after block, row-major: full blocks take BS*BS entries
        const double *m = a.bj_inv + lo * BS + (int64_t)li * sz;
        double mv[BS], rv[BS], qv[BS];
#pragma unroll
        for (int j = 0; j < BS; ++j) mv[j] = j < sz ? __ldcs(&m[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < BS; ++j) rv[j] = j < sz ? a.in3[lo + j] : 0.0;
        if (upd) {
#pragma unroll
            for (int j = 0; j < BS; ++j) qv[j] = j < sz ? a.in1[lo + j] : 0.0;
        }
        double r = 0.0, z = 0.0;
#pragma unroll
        for (int j = 0; j < BS; ++j) {
            if (j < sz) {
                double rj = rv[j];
                if (upd) rj = __dsub_rn(rj, __dmul_rn(t, qv[j]));
                if (j == li) r = rj;
                z = __dadd_rn(z, __dmul_rn(mv[j], rj));
            }
        }
        if (upd) a.out0[i] = __dadd_rn(a.out0[i], __dmul_rn(t, a.in0[i]));
        a.out1[i] = r;
        a.out2[i] = z;
        red[1] = __dadd_rn(red[1], fabs(r));
        red[0] = __dadd_rn(red[0], __dmul_rn(r, z));
    }
    const int last = grid_reduce<2>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
    if (a.cond && last == 2 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
}

// The BiCGStab update kernels move 128 bits per thread and trip (two rows; an odd last row is
// handled by one thread): as scalar 8-byte loops they ran at 3.7-4.9 TB/s (ncu, round 2), far from
// what k_cg_xr reaches on the same kind of stream.
#define GRID_STRIDE2(i, n2) \
    _Pragma("unroll 2") for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n2); \
                             i += (int64_t)gridDim.x * blockDim.x)

// bicgstab::step_1 + y = M^-1 p     in0 = r, in1 = v, in2 = inv_diag ; out0 = p, out1 = y
template <int PK>
__device__ __forceinline__ void bicg_step1_elem(bool p_is_r, double t, double omega, double r, double v,
                                                double d, double &p, double &y)
{
    if (p_is_r) p = r;
    else p = __dadd_rn(r, __dmul_rn(t, __dsub_rn(p, __dmul_rn(omega, v))));
    if (PK == 1) y = __dmul_rn(p, d);
}

template <int PK>
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_bicg_step1(const VecK a)
{
    if (a.guard_done && a.state->done) return;
    const bool p_is_r = a.state->flag_p_is_z != 0;
    const double t = a.state->coef_p, omega = a.state->omega;
    const double2 *__restrict__ r2 = reinterpret_cast<const double2 *>(a.in0);
    const double2 *__restrict__ v2 = reinterpret_cast<const double2 *>(a.in1);
    const double2 *__restrict__ d2 = reinterpret_cast<const double2 *>(a.in2);
    double2 *__restrict__ p2 = reinterpret_cast<double2 *>(a.out0);
    double2 *__restrict__ y2 = reinterpret_cast<double2 *>(a.out1);
    const int64_t n2 = a.n >> 1;
    GRID_STRIDE2(i, n2) {
        const double2 r = r2[i];
        double2 p = make_double2(0.0, 0.0), v = p, d = p, y = p;
        if (!p_is_r) {
            p = p2[i];
            v = v2[i];
        }
        if (PK == 1) d = d2[i];
        bicg_step1_elem<PK>(p_is_r, t, omega, r.x, v.x, d.x, p.x, y.x);
        bicg_step1_elem<PK>(p_is_r, t, omega, r.y, v.y, d.y, p.y, y.y);
        p2[i] = p;
        if (PK == 1) y2[i] = y;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        double p = a.out0[i], y = 0.0;
        bicg_step1_elem<PK>(p_is_r, t, omega, a.in0[i], a.in1[i], PK == 1 ? a.in2[i] : 0.0, p, y);
        a.out0[i] = p;
        if (PK == 1) a.out1[i] = y;
    }
}

// bicgstab::step_2 + |s|_1 + z = M^-1 s   in0 = r, in1 = v, in2 = inv_diag ; out0 = s, out1 = z
template <int PK>
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_bicg_step2(const VecK a)
{
    if (a.guard_done && a.state->done) return;
    const bool upd = a.state->beta != 0.0;
    const double alpha = a.state->alpha;
    double red[2] = {0.0, 0.0};
    const double2 *__restrict__ r2 = reinterpret_cast<const double2 *>(a.in0);
    const double2 *__restrict__ v2 = reinterpret_cast<const double2 *>(a.in1);
    const double2 *__restrict__ d2 = reinterpret_cast<const double2 *>(a.in2);
    double2 *__restrict__ s2 = reinterpret_cast<double2 *>(a.out0);
    double2 *__restrict__ z2 = reinterpret_cast<double2 *>(a.out1);
    const int64_t n2 = a.n >> 1;
    GRID_STRIDE2(i, n2) {
        double2 s = r2[i];
        if (upd) {
            const double2 v = v2[i];
            s.x = __dsub_rn(s.x, __dmul_rn(alpha, v.x));
            s.y = __dsub_rn(s.y, __dmul_rn(alpha, v.y));
        }
        s2[i] = s;
        red[1] = __dadd_rn(__dadd_rn(red[1], fabs(s.x)), fabs(s.y));
        if (PK == 1) {
            const double2 d = d2[i];
            z2[i] = make_double2(__dmul_rn(s.x, d.x), __dmul_rn(s.y, d.y));
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        double s = a.in0[i];
        if (upd) s = __dsub_rn(s, __dmul_rn(alpha, a.in1[i]));
        a.out0[i] = s;
        red[1] = __dadd_rn(red[1], fabs(s));
        if (PK == 1) a.out1[i] = __dmul_rn(s, a.in2[i]);
    }
    const int last = grid_reduce<2>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
    if (a.cond && last == 2 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);   // stopped at the s check
}

// bicgstab::step_3 + <rr,r> + |r|_1
//   in0 = s, in1 = t, in2 = y, in3 = z, in4 = rr ; out0 = x, out1 = r
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_bicg_step3(const VecK a)
{
    if (a.guard_done && a.state->done) {
        // the last kernel of a loop body: the criterion fired earlier (before the loop, at the s
        // check of this iteration, or in the first half of the body)
        if (a.cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
        return;
    }
    const double alpha = a.state->alpha, omega = a.state->omega;
    double red[2] = {0.0, 0.0};
    const double2 *__restrict__ s2 = reinterpret_cast<const double2 *>(a.in0);
    const double2 *__restrict__ t2 = reinterpret_cast<const double2 *>(a.in1);
    const double2 *__restrict__ y2 = reinterpret_cast<const double2 *>(a.in2);
    const double2 *__restrict__ z2 = reinterpret_cast<const double2 *>(a.in3);
    const double2 *__restrict__ rr2 = reinterpret_cast<const double2 *>(a.in4);
    double2 *__restrict__ x2 = reinterpret_cast<double2 *>(a.out0);
    double2 *__restrict__ r2 = reinterpret_cast<double2 *>(a.out1);
    const int64_t n2 = a.n >> 1;
    GRID_STRIDE2(i, n2) {
        const double2 s = s2[i], t = t2[i], y = y2[i], z = z2[i], rr = rr2[i];
        double2 x = x2[i], r;
        x.x = __dadd_rn(x.x, __dadd_rn(__dmul_rn(alpha, y.x), __dmul_rn(omega, z.x)));
        x.y = __dadd_rn(x.y, __dadd_rn(__dmul_rn(alpha, y.y), __dmul_rn(omega, z.y)));
        r.x = __dsub_rn(s.x, __dmul_rn(omega, t.x));
        r.y = __dsub_rn(s.y, __dmul_rn(omega, t.y));
        x2[i] = x;
        r2[i] = r;
        red[0] = __dadd_rn(__dadd_rn(red[0], __dmul_rn(rr.x, r.x)), __dmul_rn(rr.y, r.y));
        red[1] = __dadd_rn(__dadd_rn(red[1], fabs(r.x)), fabs(r.y));
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        a.out0[i] = __dadd_rn(a.out0[i], __dadd_rn(__dmul_rn(alpha, a.in2[i]), __dmul_rn(omega, a.in3[i])));
        const double r = __dsub_rn(a.in0[i], __dmul_rn(omega, a.in1[i]));
        a.out1[i] = r;
        red[0] = __dadd_rn(red[0], __dmul_rn(a.in4[i], r));
        red[1] = __dadd_rn(red[1], fabs(r));
    }
    const int last = grid_reduce<2>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
    if (a.cond && last == 2 && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0);
}

// bicgstab::finalize  x += alpha y when the solver stopped at the first check
__global__ void __launch_bounds__(kT, kBlas1BlocksPerSM) k_bicg_finalize(const VecK a)
{
    if (!a.state->stop_half) return;
    const double alpha = a.state->alpha;
    GRID_STRIDE(i, a.n) a.out0[i] = __dadd_rn(a.out0[i], __dmul_rn(alpha, a.in0[i]));
}

__global__ void k_epilogue(SolveState *state, int epi, EpiArgs ea, int guard_done)
{
    if (guard_done && state->done) return;
    run_epilogue(epi, state, ea);
}

}  // namespace

// ---------------------------------------------------------------------------

bool use_p2p(const Context *ctx);

// ar_count: how many leading state->red[] slots this launch all-reduces on the
// peer-memory path before its epilogue (ignored on one rank / the NCCL path)
static VecK base_args(Context *ctx, int epi, bool guard, int ar_count = 0)
{
    VecK a;
    std::memset(&a, 0, sizeof(a));
    a.n = ctx->n;
    a.state = ctx->d_state;
    a.partials = ctx->d_partials;
    a.ticket = ctx->d_ticket;
    a.epi = epi;
    a.inline_epi = (ctx->n_ranks == 1 || use_p2p(ctx)) ? 1 : 0;
    a.guard_done = guard ? 1 : 0;
    a.ea = make_epi_args(ctx, ar_count);
    return a;
}

static int vec_grid(const Context *ctx)
{
    int64_t g = ((int64_t)ctx->n + kT - 1) / kT;
    if (g > ctx->blas1_blocks) g = ctx->blas1_blocks;
    if (g < 1) g = 1;
    return (int)g;
}

// multi-rank tail of a reduction: all-reduce the partial sums, then run the
// scalar epilogue in a one-thread kernel (one rank: already done in-kernel)
int finish_reduction(Context *ctx, int count, int epi, bool guard)
{
    if (ctx->n_ranks == 1 || use_p2p(ctx)) return OGL_OK;   // done inside the kernel
    OGL_TRY(allreduce_red(ctx, count));
    if (epi != EPI_NONE) {
        k_epilogue<<<1, 1, 0, ctx->stream>>>(ctx->d_state, epi, make_epi_args(ctx), guard ? 1 : 0);
        ctx->launches++;
    }
    return OGL_OK;
}

#define LAUNCH(kernel, args)                                        \
    do {                                                            \
        kernel<<<vec_grid(ctx), kT, 0, ctx->stream>>>(args);        \
        ctx->launches++;                                            \
    } while (0)

int vec_fill(Context *ctx, double *v, double value)
{
    k_fill<<<vec_grid(ctx), kT, 0, ctx->stream>>>(ctx->n, v, value);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

int vec_scale(Context *ctx, double *v, double s)
{
    k_scale<<<vec_grid(ctx), kT, 0, ctx->stream>>>(ctx->n, v, s);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

static int pk_of(const Context *ctx)
{
    if (ctx->precond_kind == OGL_PRECOND_NONE) return 0;
    if (ctx->precond_kind == OGL_PRECOND_ISAI || ctx->precond_kind == OGL_PRECOND_GISAI ||
        is_tri_precond(ctx->precond_kind) || ctx->precond_kind == OGL_PRECOND_MULTIGRID)
        return 3;   // applied by its own launches (precond_apply)
    return ctx->max_block_size == 1 ? 1 : 2;
}

// Everything up to and including the first criterion call:
//   w = A (mean(x) 1)        StoppingCriterion.C:11-30
//   r = b - A x              Ginkgo: r = b; A->apply(-1, x, 1, r)
//   z / rr, rho, |r|_1, normFactor, check(iter 0)
// mode 0: CG (z = M^-1 r, rho = <r,z>), 1: BiCGStab (rr = r, rho = <rr,r>),
// 2: GMRES (<r,r>)
int solve_prologue(Context *ctx, int mode, double *r, double *z, double *rr, double *w,
                   double *tmp, int epi)
{
    const int pk = pk_of(ctx);
    // mean(x): local sum -> weighted local mean -> sum over ranks
    {
        VecK a = base_args(ctx, EPI_MEAN_LOCAL, false);
        a.inline_epi = 1;   // the local transform always runs in-kernel
        a.ea = make_epi_args(ctx, 1, /*ar_after_epi=*/true);
        a.in0 = ctx->d_x;
        LAUNCH(k_sum, a);
        if (ctx->n_ranks > 1) OGL_TRY(allreduce_red(ctx, 1));
        VecK f = base_args(ctx, EPI_NONE, false);
        f.out0 = tmp;
        LAUNCH(k_fill_from_red, f);
    }
    {
        SpmvArgs s;
        s.x = tmp;
        s.y = w;
        OGL_TRY(dist_spmv(ctx, s));
    }
    {
        SpmvArgs s;
        s.x = ctx->d_x;
        s.y = r;
        s.y_in = ctx->d_b;
        s.advanced = true;
        s.alpha = -1.0;
        s.beta = 1.0;
        OGL_TRY(dist_spmv(ctx, s));
    }
    const bool defer = (pk >= 2 && mode == 0);   // block Jacobi / ISAI finish <r,z> afterwards
    VecK a = base_args(ctx, defer ? EPI_NONE : epi, false, defer ? 0 : 3);
    a.in0 = r;
    a.in1 = ctx->d_b;
    a.in2 = w;
    a.in3 = ctx->d_inv_diag;
    a.out0 = z;
    a.out1 = rr;
    if (mode == 0) {
        if (pk == 0) LAUNCH((k_init_norms<0, 0>), a);
        else if (pk == 1) LAUNCH((k_init_norms<1, 0>), a);
        else {
            LAUNCH((k_init_norms<2, 0>), a);
            OGL_TRY(precond_apply(ctx, r, z, r, 0, false, epi, ctx->n_ranks == 1 || use_p2p(ctx), 3));
        }
    } else if (mode == 1) {
        LAUNCH((k_init_norms<0, 1>), a);
    } else {
        LAUNCH((k_init_norms<0, 2>), a);
    }
    OGL_CUDA(ctx, cudaGetLastError());
    return finish_reduction(ctx, 3, epi, false);
}

// CG on several GPUs over peer memory, scalar (or no) preconditioner: "ghost p".
// The only vector the SpMV needs from the neighbours is p = z + beta p_old.  Its
// boundary z is pushed by the kernel that computes z (k_cg_xr) and published by
// the all-reduce of rho that the same kernel ends with; every rank then updates
// the ghost entries of p itself.  The iteration has no halo handshake left --
// its two all-reduces are its only rendezvous -- and the SpMV is the plain
// single-GPU kernel over the ghosted CSR.
static bool ghost_p_mode(const Context *ctx)
{
    return ctx->ghost_p != 0 && ctx->n_ranks > 1 && fused_halo_ok(ctx) && pk_of(ctx) < 2;
}

static int cg_iteration(Context *ctx, const double *r_old, double *r, double *z, const double *p_old,
                        double *p, double *q)
{
    const int pk = pk_of(ctx);
    const double *zz = pk == 0 ? r_old : z;   // the residual this iteration starts from
    const bool ghost = ghost_p_mode(ctx);
    const bool pack = !ghost && fused_halo_ok(ctx) && ctx->n_send > 0;
    // profile_stride > 0 (graphs off): bracket every stride-th SpMV of the real loop with CUDA
    // events on the launching stream
    const bool sample = ctx->profile_stride > 0 && !ctx->capturing &&
                        (ctx->profile_iter++ % ctx->profile_stride) == 0 &&
                        ctx->profile_used + 2 <= (int)ctx->profile_events.size();
    bool fused_p = false;
    if (ctx->fuse_p && (ctx->n_ranks == 1 || ghost)) {
        // p-update inside the ELL SpMV (ell.cu): two launches per iteration instead of three
        if (sample) cudaEventRecord(ctx->profile_events[ctx->profile_used], ctx->stream);
        const int rc = spmv_ell_cgp(ctx, zz, p_old, p, q, ghost);
        if (rc == OGL_OK) {
            fused_p = true;
            if (sample) {
                cudaEventRecord(ctx->profile_events[ctx->profile_used + 1], ctx->stream);
                ctx->profile_used += 2;
            }
        } else if (rc != OGL_ERR_UNSUPPORTED) {
            return rc;
        }
    }
    if (!fused_p) {
    {
        VecK a = base_args(ctx, EPI_NONE, true);
        a.in0 = zz;
        a.in1 = p_old;
        a.out0 = p;
        a.send_idx = ctx->d_send_idxs;
        a.n_send = ctx->n_send;
        a.pack = ghost ? 2 : (pack ? 1 : 0);
        a.ea.trace_tag = 10;
        OGL_CUDA(ctx, launch_pdl(k_cg_p, vec_grid(ctx), kT, 0, ctx->stream, ctx->use_pdl != 0, a));
        ctx->launches++;
    }
    {
        SpmvArgs s;
        s.x = p;
        s.y = q;
        s.dot_with = p;
        s.nred = 1;
        s.guard_done = true;
        s.epi = EPI_CG_BETA;
        s.halo_stored = pack;
        s.ghost_x = ghost;
        if (sample) cudaEventRecord(ctx->profile_events[ctx->profile_used++], ctx->stream);
        OGL_TRY(dist_spmv(ctx, s));
        if (sample) cudaEventRecord(ctx->profile_events[ctx->profile_used++], ctx->stream);
    }
    }
    {
        VecK a = base_args(ctx, pk == 3 ? EPI_NONE : EPI_CG_RHO_CHECK, true, pk == 3 ? 0 : 2);
        a.in0 = p;
        a.in1 = q;
        a.in2 = ctx->d_inv_diag;
        a.in3 = r_old;
        a.out0 = ctx->d_x;
        a.out1 = r;
        a.out2 = z;
        a.cond = ctx->cond_handle;
        a.ea.trace_tag = 30;
        if (ghost) {
            a.pack = 1;
            a.send_idx = ctx->d_send_idxs;
            a.n_send = ctx->n_send;
            a.push_dst = ctx->d_push_dst;
        }
#define LAUNCH_XR(PKV)                                                                             \
    do {                                                                                           \
        OGL_CUDA(ctx, launch_pdl(k_cg_xr<PKV>, vec_grid(ctx), kT, 0, ctx->stream, ctx->use_pdl != 0, a)); \
        ctx->launches++;                                                                           \
    } while (0)
        if (pk == 0) LAUNCH_XR(0);
        else if (pk == 1) LAUNCH_XR(1);
        else if (pk == 3) {
            // ISAI: x, r', |r'|_1 here; z = M^-1 r', <r',z> and the criterion in the apply launches
            LAUNCH_XR(2);
            OGL_TRY(precond_apply(ctx, r, z, r, 0, true, EPI_CG_RHO_CHECK,
                                  ctx->n_ranks == 1 || use_p2p(ctx), 2));
        } else {
            a.bj_row_block = ctx->d_row_block;
            a.bj_block_ptrs = ctx->d_block_ptrs;
            a.bj_block_offs = ctx->d_block_offs;
            a.bj_inv = ctx->d_inv_blocks;
#define LAUNCH_BJ(KERNEL) \
    OGL_CUDA(ctx, launch_pdl(KERNEL, vec_grid(ctx), kT, 0, ctx->stream, ctx->use_pdl != 0, a))
            const int bs = ctx->bj_uniform ? (int)ctx->max_block_size : 0;
            if (bs == 2) LAUNCH_BJ(k_cg_xr_bj_uniform<2>);
            else if (bs == 4) LAUNCH_BJ(k_cg_xr_bj_uniform<4>);
            else if (bs == 8) LAUNCH_BJ(k_cg_xr_bj_uniform<8>);
            else LAUNCH_BJ(k_cg_xr_bj);
#undef LAUNCH_BJ
            ctx->launches++;
        }
    }
    return finish_reduction(ctx, 2, EPI_CG_RHO_CHECK, true);
}

static int bicg_iteration(Context *ctx, double *r, double *rr, double *p, double *v,
                          double *s, double *t, double *y, double *z)
{
    const int pk = pk_of(ctx);
    {
        VecK a = base_args(ctx, EPI_NONE, true);
        a.in0 = r;
        a.in1 = v;
        a.in2 = ctx->d_inv_diag;
        a.out0 = p;
        a.out1 = y;
        if (pk == 1) LAUNCH(k_bicg_step1<1>, a);
        else LAUNCH(k_bicg_step1<0>, a);
        if (pk >= 2) OGL_TRY(precond_apply(ctx, p, y, nullptr, 0, true, EPI_NONE, false, 0));
    }
    const double *yy = pk == 0 ? p : y;
    {
        SpmvArgs sa;
        sa.x = yy;
        sa.y = v;
        sa.dot_with = rr;
        sa.nred = 1;
        sa.guard_done = true;
        sa.epi = EPI_BICG_ALPHA;
        OGL_TRY(dist_spmv(ctx, sa));
    }
    {
        VecK a = base_args(ctx, EPI_BICG_CHECK_S, true, 2);
        a.in0 = r;
        a.in1 = v;
        a.in2 = ctx->d_inv_diag;
        a.out0 = s;
        a.out1 = z;
        a.cond = ctx->cond_handle;
        if (pk == 1) LAUNCH(k_bicg_step2<1>, a);
        else LAUNCH(k_bicg_step2<0>, a);
        OGL_TRY(finish_reduction(ctx, 2, EPI_BICG_CHECK_S, true));
        if (pk >= 2) OGL_TRY(precond_apply(ctx, s, z, nullptr, 0, true, EPI_NONE, false, 0));
    }
    const double *zz = pk == 0 ? s : z;
    {
        SpmvArgs sa;
        sa.x = zz;
        sa.y = t;
        sa.dot_with = s;
        sa.nred = 2;
        sa.guard_done = true;
        sa.epi = EPI_BICG_OMEGA;
        OGL_TRY(dist_spmv(ctx, sa));
    }
    {
        VecK a = base_args(ctx, EPI_BICG_RHO_CHECK, true, 2);
        a.in0 = s;
        a.in1 = t;
        a.in2 = yy;
        a.in3 = zz;
        a.in4 = rr;
        a.out0 = ctx->d_x;
        a.out1 = r;
        a.cond = ctx->cond_handle;
        LAUNCH(k_bicg_step3, a);
    }
    return finish_reduction(ctx, 2, EPI_BICG_RHO_CHECK, true);
}

int init_state(Context *ctx, const ogl_solve_params *p)
{
    SolveState s;
    std::memset(&s, 0, sizeof(s));
    // Ginkgo initialize kernels: prev_rho = 1 (the epilogue shifts rho -> prev_rho
    // before storing the first dot product, so `rho` starts at 1 as well)
    s.rho = 1.0;
    s.prev_rho = 1.0;
    s.alpha = s.beta = s.gamma = s.omega = 1.0;
    s.norm_factor = 1.0;
    s.tolerance = p->tolerance;
    s.rel_tol = p->rel_tol;
    s.min_iter = p->min_iter;
    s.max_iter = p->max_iter;
    s.frequency = p->frequency < 1 ? 1 : p->frequency;
    s.export_res = p->export_res ? 1 : 0;
    s.krylov_dim = p->krylov_dim > 0 ? p->krylov_dim : 100;
    if (p->export_res) {
        const int cap = p->max_iter + 2;
        if (cap > ctx->history_cap) {
            invalidate_graph(ctx);   // a chunk captured without history holds a null pointer
            OGL_TRY(dev_alloc(ctx, &ctx->d_history, cap));
            ctx->history_cap = cap;
        }
        OGL_CUDA(ctx, cudaMemsetAsync(ctx->d_history, 0, sizeof(double) * ctx->history_cap,
                                      ctx->stream));
        s.history_cap = ctx->history_cap;
    }
    *ctx->h_state = s;
    OGL_CUDA(ctx, cudaMemcpyAsync(ctx->d_state, ctx->h_state, sizeof(SolveState),
                                  cudaMemcpyHostToDevice, ctx->stream));
    OGL_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned int), ctx->stream));
    return OGL_OK;
}

// Enqueue chunks of iterations until the device-side criterion reports `done`.
// `enqueue` puts ONE iteration on ctx->stream.  On one rank a chunk is captured
// once into a CUDA graph and replayed (launch-bound at 1 M rows otherwise).
template <typename F>
static int run_chunks(Context *ctx, int solver, int64_t max_criterion_calls, F enqueue)
{
    // an even number of iterations per chunk: CG alternates two p buffers and a
    // replayed chunk must end where it started
    cudaStream_t st = ctx->stream;
    const bool graph_ok = ctx->use_graph && (ctx->n_ranks == 1 || use_p2p(ctx)) &&
                          ctx->profile_stride == 0;
    // CG (its x/r-update kernels carry the criterion): the chunk becomes the body
    // of a WHILE node; k_cg_xr clears the condition when the criterion fires, so a solve is one
    // graph launch with no early-exit launches behind the last iteration and no host polling
    // (preconditioners with their own apply launches keep the host-polled chunks: their criterion
    // call sits in the apply kernels)
    const bool loop = graph_ok && ctx->device_loop &&
                      ((solver == OGL_SOLVER_CG && pk_of(ctx) != 3) || (solver == OGL_SOLVER_BICGSTAB && pk_of(ctx) < 2));
    int chunk = (int)(loop ? ctx->loop_iters : ctx->chunk_iters);
    // a BiCGStab iteration is two SpMVs + three updates: short bodies waste fewer early-exit launches
    if (loop && solver == OGL_SOLVER_BICGSTAB && chunk > 4) chunk = 4;
    if (chunk < 1) chunk = 1;
    chunk += chunk & 1;
    const int64_t sig0 = ((int64_t)solver << 48) ^ ((int64_t)pk_of(ctx) << 40) ^ (ctx->use_pdl << 60) ^
                         ((int64_t)chunk << 32) ^ (int64_t)ctx->n ^ (ctx->spmv_variant << 56);
    const int64_t sig = sig0 ^ ((int64_t)(loop ? 1 : 0) << 62) ^ ((int64_t)(ctx->bj_uniform ? 1 : 0) << 61) ^
                        ((int64_t)ctx->max_block_size << 24) ^ ((int64_t)ctx->precond_kind << 52) ^
                        ((int64_t)ctx->tri_variant << 30);
    if (graph_ok && (!ctx->graph_exec || ctx->graph_sig != sig)) {
        if (ctx->graph_exec) {
            cudaGraphExecDestroy(ctx->graph_exec);
            ctx->graph_exec = nullptr;
        }
        cudaGraph_t graph = nullptr, body = nullptr;
        const int64_t launches_before = ctx->launches;
        if (loop) {
            OGL_CUDA(ctx, cudaGraphCreate(&graph, 0));
            cudaGraphConditionalHandle handle;
            cudaError_t ce = cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault);
            cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = handle;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            if (ce == cudaSuccess) ce = cudaGraphAddNode(&node, graph, nullptr, 0, &np);
            if (ce != cudaSuccess) {
                cudaGraphDestroy(graph);
                return fail(ctx, OGL_ERR_CUDA, std::string("conditional graph node: ") + cudaGetErrorString(ce));
            }
            body = np.conditional.phGraph_out[0];
            ctx->cond_handle = handle;
            OGL_CUDA(ctx, cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        } else {
            OGL_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        }
        int rc = OGL_OK;
        ctx->capturing = true;
        for (int i = 0; i < chunk && rc == OGL_OK; ++i) rc = enqueue();
        ctx->capturing = false;
        ctx->cond_handle = 0;
        cudaGraph_t captured = nullptr;
        cudaError_t e = cudaStreamEndCapture(st, &captured);
        if (!loop) graph = captured;
        const int64_t kernels_per_chunk = ctx->launches - launches_before;
        ctx->launches = launches_before;   // captured, not executed
        if (rc != OGL_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        OGL_CUDA(ctx, e);
        e = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        OGL_CUDA(ctx, e);
        ctx->graph_sig = sig;
        ctx->graph_kernels = kernels_per_chunk;
        ctx->graph_is_loop = loop;
    }
    if (graph_ok && ctx->graph_is_loop) {
        // one launch: the loop ends on the device; solve() reads the state back afterwards
        OGL_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, st));
        ctx->loop_body_iters = chunk;   // solve() counts the launches once it has the final state
        return OGL_OK;
    }
    cudaEvent_t ev[2] = {ctx->ev_poll[0], ctx->ev_poll[1]};
    int64_t enqueued = 0;
    int c = 0;
    while (true) {
        if (graph_ok) {
            OGL_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, st));
            ctx->launches += ctx->graph_kernels;
        } else {
            for (int i = 0; i < chunk; ++i) OGL_TRY(enqueue());
        }
        enqueued += chunk;
        OGL_CUDA(ctx, cudaMemcpyAsync(&ctx->h_state[1 + (c & 1)], ctx->d_state, sizeof(SolveState),
                                      cudaMemcpyDeviceToHost, st));
        OGL_CUDA(ctx, cudaEventRecord(ev[c & 1], st));
        if (c >= 1) {
            OGL_CUDA(ctx, cudaEventSynchronize(ev[(c - 1) & 1]));
            if (ctx->h_state[1 + ((c - 1) & 1)].done) break;
        }
        // safety net: the criterion stops at max_iter at the latest
        if (enqueued > max_criterion_calls + 4 * (int64_t)chunk) break;
        ++c;
    }
    return OGL_OK;
}

int solve(Context *ctx, const ogl_solve_params *p, ogl_solve_result *res)
{
    if (!p || !res) return fail(ctx, OGL_ERR_INVALID, "null params/result");
    if (!ctx->have_pattern || !ctx->have_values)
        return fail(ctx, OGL_ERR_INVALID, "ogl_solve before the matrix is assembled");
    if (!ctx->have_b || !ctx->have_x)
        return fail(ctx, OGL_ERR_INVALID, "ogl_solve before b and x are uploaded");
    if (!ctx->have_precond)
        return fail(ctx, OGL_ERR_INVALID, "ogl_solve before ogl_precond_setup");
    if (ctx->n_ranks > 1 && !ctx->have_partition)
        return fail(ctx, OGL_ERR_INVALID, "ogl_solve on several ranks before ogl_partition_create");
    if (ctx->n_ranks > 1 && !ctx->comm && !use_p2p(ctx))
        return fail(ctx, OGL_ERR_INVALID,
                    "context without NCCL id: ogl_partition_export / ogl_partition_connect must set up the "
                    "peer-memory windows before ogl_solve (comm_mode 1 is not available)");
    if (p->frequency < 1) return fail(ctx, OGL_ERR_INVALID, "frequency must be >= 1");
    if (is_tri_precond(ctx->precond_kind)) OGL_TRY(tri_ensure_structure(ctx));   // + its work vectors, before any capture
    if (ctx->precond_kind == OGL_PRECOND_MULTIGRID) OGL_TRY(mg_ensure(ctx));
    if (p->solver == OGL_SOLVER_GMRES) return solve_gmres(ctx, p, res);
    if (p->solver != OGL_SOLVER_CG && p->solver != OGL_SOLVER_BICGSTAB)
        return fail(ctx, OGL_ERR_UNSUPPORTED, "unknown solver kind");

    const int64_t launches0 = ctx->launches;
    cudaStream_t st = ctx->stream;
    OGL_TRY(init_state(ctx, p));
    ctx->profile_iter = 0;
    ctx->profile_used = 0;
    if (ctx->profile_stride > 0 && ctx->profile_events.empty()) {
        ctx->profile_events.resize(512);
        for (auto &ev : ctx->profile_events) OGL_CUDA(ctx, cudaEventCreate(&ev));
    }
    OGL_CUDA(ctx, cudaEventRecord(ctx->ev_t0, st));
    double *r, *z, *pv, *q, *w, *tmp;
    OGL_TRY(get_work(ctx, 0, &r));
    OGL_TRY(get_work(ctx, 1, &z));
    OGL_TRY(get_work(ctx, 2, &pv));
    OGL_TRY(get_work(ctx, 3, &q));
    w = q;      // w and tmp are only live during the prologue
    tmp = pv;
    if (p->solver == OGL_SOLVER_CG) {
        double *pv2;
        OGL_TRY(get_work(ctx, 9, &pv2));
        OGL_TRY(solve_prologue(ctx, 0, r, z, nullptr, w, tmp, EPI_INIT_CHECK));
        OGL_CUDA(ctx, cudaMemsetAsync(pv, 0, sizeof(double) * ctx->work_len, st));
        OGL_CUDA(ctx, cudaMemsetAsync(pv2, 0, sizeof(double) * ctx->work_len, st));
        const bool fused = pcg_fused_ok(ctx);
        if (!fused) OGL_TRY(ell_prepare_for_loop(ctx, ghost_p_mode(ctx)));
        if (ghost_p_mode(ctx)) OGL_TRY(push_boundary(ctx, pk_of(ctx) == 0 ? r : z));   // boundary z0
        double *r_alt;   // r is ping-ponged like p: the x/r-update writes the other buffer
        OGL_TRY(get_work(ctx, 10, &r_alt));
        if (fused) {
            // the whole loop in one persistent kernel (pcg_fused.cu)
            OGL_TRY(pcg_fused_run(ctx, r, r_alt, z, pv, pv2, q, p->max_iter));
        } else {
            int flip = 0;   // the p and r buffers alternate; a chunk holds an even number of iterations
            OGL_TRY(run_chunks(ctx, OGL_SOLVER_CG, p->max_iter, [&]() {
                double *p_old = flip ? pv2 : pv, *p_new = flip ? pv : pv2;
                double *r_old = flip ? r_alt : r, *r_new = flip ? r : r_alt;
                flip ^= 1;
                return cg_iteration(ctx, r_old, r_new, z, p_old, p_new, q);
            }));
        }
    } else {
        double *rr, *v, *s, *t, *y;
        OGL_TRY(get_work(ctx, 4, &rr));
        OGL_TRY(get_work(ctx, 5, &v));
        OGL_TRY(get_work(ctx, 6, &s));
        OGL_TRY(get_work(ctx, 7, &t));
        OGL_TRY(get_work(ctx, 8, &y));
        OGL_TRY(solve_prologue(ctx, 1, r, nullptr, rr, w, tmp, EPI_BICG_INIT_CHECK));
        OGL_CUDA(ctx, cudaMemsetAsync(pv, 0, sizeof(double) * ctx->n, st));
        OGL_CUDA(ctx, cudaMemsetAsync(v, 0, sizeof(double) * ctx->n, st));
        OGL_TRY(run_chunks(ctx, OGL_SOLVER_BICGSTAB, p->max_iter, [&]() {
            return bicg_iteration(ctx, r, rr, pv, v, s, t, y, z);
        }));
        VecK a = base_args(ctx, EPI_NONE, false);
        a.in0 = pk_of(ctx) == 0 ? pv : y;
        a.out0 = ctx->d_x;
        LAUNCH(k_bicg_finalize, a);
    }
    OGL_CUDA(ctx, cudaEventRecord(ctx->ev_t1, st));
    OGL_CUDA(ctx, cudaMemcpyAsync(&ctx->h_state[0], ctx->d_state, sizeof(SolveState),
                                  cudaMemcpyDeviceToHost, st));
    OGL_CUDA(ctx, cudaStreamSynchronize(st));
    OGL_CUDA(ctx, cudaGetLastError());
    const SolveState &hs = ctx->h_state[0];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1);
    std::memset(res, 0, sizeof(*res));
    res->init_residual = hs.init_res;
    res->final_residual = hs.res;
    res->norm_factor = hs.norm_factor;
    res->criterion_calls = hs.iter;
    // GKOBiCGStab.H:112-115: two criterion calls per iteration
    res->n_iterations = p->solver == OGL_SOLVER_BICGSTAB ? hs.iter / 2 : hs.iter;
    res->solve_us = ms * 1e3;
    // The L1 norm rides inside the fused update kernel; what a criterion evaluation costs here is
    // the scalar epilogue that runs it, timed on the device (reduce.cuh:criterion_check).  The host
    // layers divide the time per iteration by it exactly as lduLduBase.H:286-293 does, so the
    // reference's adaptive minIter / evaluation frequency is live (a cheap evaluation gives
    // frequency 1 and minIter = relaxationFactor * previous iterations).
    res->resnorm_us = hs.crit_ns > 0 ? (double)hs.crit_ns * 1e-3 : 0.0;
    if (ctx->loop_body_iters > 0) {
        // device-side loop: whole bodies ran, up to and including the one the criterion fired in
        int64_t calls = hs.iter > 0 ? hs.iter - 1 : 0;   // minus the prologue's criterion call
        if (p->solver == OGL_SOLVER_BICGSTAB) calls = (calls + 1) / 2;   // two criterion calls per iteration
        int64_t bodies = (calls + ctx->loop_body_iters - 1) / ctx->loop_body_iters;
        if (bodies < 1) bodies = 1;
        ctx->launches += bodies * ctx->graph_kernels;
        ctx->loop_body_iters = 0;
    }
    res->kernel_launches = ctx->launches - launches0;
    if (ctx->profile_used >= 2) {
        double total_ms = 0.0;
        int pairs = 0;
        for (int i = 0; i + 1 < ctx->profile_used; i += 2) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ctx->profile_events[i], ctx->profile_events[i + 1]) == cudaSuccess) {
                total_ms += t;
                ++pairs;
            }
        }
        if (pairs > 0) {
            res->spmv_us_avg = total_ms * 1e3 / pairs;
            res->spmv_samples = pairs;
        }
    }
    if (hs.comm_error == 2)
        return fail(ctx, OGL_ERR_CUDA, "ILU/IC triangular sweep timed out waiting for a row it depends on");
    if (hs.comm_error)
        return fail(ctx, OGL_ERR_NCCL, "peer synchronisation timed out (a rank left the solve?)");
    if (!hs.done)
        return fail(ctx, OGL_ERR_CUDA, "solver loop ended without the criterion firing");
    return OGL_OK;
}

}  // namespace ogl
