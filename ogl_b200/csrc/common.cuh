// Shared declarations of libogl_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "ogl_b200.h"

namespace ogl {

struct CommDev;

using label = int32_t;
using scalar = double;

constexpr int kMaxReduce = 4;          // scalars reduced by one kernel launch
constexpr int kNumSM = 148;            // B200
constexpr int kBlas1Threads = 256;
constexpr int kTraceCap = 1 << 16;   // timeline events kept per context
constexpr int kBlas1BlocksPerSM = 4;   // 1024 resident threads per SM, <= 64 registers each
constexpr int kMaxPartialBlocks = 4096;

// Device-resident scalar state of a solve.  Every kernel of the iteration reads
// its coefficients from here, so the host never has to see alpha/beta/rho
// (StoppingCriterion.C:95-97 costs the reference one D2H + sync per iteration).
struct SolveState {
    // --- Krylov scalars
    double rho, prev_rho, beta, alpha, omega, gamma;
    double coef_p;        // CG: rho/prev_rho (0 => p = z); BiCGStab: step_1 factor
    double coef_x;        // CG: rho/beta (0 => skip update)
    // --- OGL criterion (StoppingCriterion.C:71-151)
    double norm_factor, init_res, res;
    double tolerance, rel_tol;
    int iter, min_iter, max_iter, frequency;
    int done;             // 1 once the criterion fired; kernels early-exit
    int stop_half;        // BiCGStab: stopped at the first (s) check => finalize
    int export_res, history_cap, n_history;
    int flag_p_is_z;      // CG step_1: prev_rho == 0
    int pad;
    // --- reduction mailboxes (local partial results, then all-reduced in place)
    double red[kMaxReduce];
    // --- GMRES
    int restart_iter, final_iter, krylov_dim, need_restart;
    double res_norm2;
    int comm_error, pad2;   // peer synchronisation timed out (multi-GPU)
    // device time of the last EVALUATED criterion call (globaltimer ns around the epilogue that
    // ran it): what the reference measures on the host around its norm1 + D2H
    // (StoppingCriterion.C:89,145-149) and feeds into the adaptive minIter / frequency
    unsigned long long crit_ns;
};

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct Context;

// error plumbing ---------------------------------------------------------------
void set_error(Context *ctx, const std::string &msg);
int fail(Context *ctx, int code, const std::string &msg);

#define OGL_CUDA(ctx, expr)                                                        \
    do {                                                                           \
        cudaError_t e__ = (expr);                                                  \
        if (e__ != cudaSuccess)                                                    \
            return ::ogl::fail(ctx, OGL_ERR_CUDA,                                  \
                               std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define OGL_NCCL(ctx, expr)                                                        \
    do {                                                                           \
        ncclResult_t r__ = (expr);                                                 \
        if (r__ != ncclSuccess)                                                    \
            return ::ogl::fail(ctx, OGL_ERR_NCCL,                                  \
                               std::string(#expr) + ": " + ncclGetErrorString(r__)); \
    } while (0)

#define OGL_TRY(expr)                     \
    do {                                  \
        int rc__ = (expr);                \
        if (rc__ != OGL_OK) return rc__;  \
    } while (0)

struct Context {
    int device = 0, rank = 0, n_ranks = 1;
    cudaStream_t stream = nullptr;       // compute stream
    bool own_stream = false;
    cudaStream_t comm_stream = nullptr;  // halo exchange
    cudaEvent_t ev_pack = nullptr, ev_recv = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_poll[2] = {nullptr, nullptr};
    ncclComm_t comm = nullptr;           // shared per (process, unique id), see capi.cu
    std::string comm_key;
    std::string error;

    // options
    int64_t spmv_variant = 0;    // 0 auto, 1 stream(LDG), 2 thread-per-row, 3 warp-per-row, 4 TMA pipeline, 5 warp tile,
                                 // 6 pipelined stream, 7 ELL, 8 merge-path (entry-balanced slices)
    int64_t chunk_iters = 16;    // iterations enqueued between two host polls
    int64_t use_graph = 1;       // replay a chunk as a CUDA graph
    int64_t profile_stride = 0;  // sample SpMV launch durations every k-th iteration
    int64_t blas1_blocks = kNumSM * kBlas1BlocksPerSM;
    int64_t stream_ctas = 0;     // persistent SpMV grid (0 = 8 CTAs per SM)
    int64_t tma_stages = 3;      // shared-memory ring depth of the TMA SpMV
    int64_t use_pdl = 0;         // programmatic dependent launch between the CG kernels (measured: no gain inside graphs)
    int64_t l2_keep_mb = 0;      // MB of the CSR stream kept L2-resident across iterations (-1 auto, 0 off)
    unsigned long long l2_policies[6] = {0, 0, 0, 0, 0, 0};   // createpolicy results, see spmv.cu:l2_policy
    bool l2_policies_ready = false;
    int64_t tile_blocked = 0;    // persistent SpMV: 1 = contiguous tile range per CTA (measured slower)

    // local pattern (a4/a5) -- resident across solves
    label n = 0, n_faces = 0, n_local_iface = 0;
    bool symmetric = true, have_pattern = false, have_pattern_before = false;
    int64_t nnz = 0;
    label *d_rows = nullptr, *d_cols = nullptr, *d_map = nullptr, *d_row_ptrs = nullptr;
    label max_row_len = 0;
    unsigned long long row_len_hist[16] = {};   // spmv.cu:k_row_len_hist: rows / entries per power-of-two length bucket
    // ELL copies (spmv_variant 7 / `matrixFormat Ell`, ell.cu): `ell` of the local matrix, `gell` of
    // the ghosted one (several ranks, CG ghost-p mode)
    struct EllMatrix {
        label *cols = nullptr;        // slot j of row r at [j * pitch + r], -1 = padding
        double *vals = nullptr;
        int64_t pitch = 0;
        int width = 0;
        bool structure_ready = false; // cols / codes match the current pattern
        bool ready = false;           // vals match the current coefficients
        // pattern-coded columns: rows whose (column - row) tuple is one of <= 255 distinct
        // patterns carry a 1-byte code instead of `width` 4-byte columns
        unsigned char *code = nullptr;   // [pitch], 255 = escape: read `cols`
        label *ptab = nullptr;           // [256 * width] deltas, kPadDelta = no entry
        int n_patterns = 0;
        int64_t n_escape = 0;
        bool coded = false;
    } ell, gell;
    label max_row_len_g = 0;     // longest row of the ghosted CSR
    int64_t ell_coded = 1;       // 0 off, 1 auto (when <= 25% of the rows escape), 2 whenever a table exists
    int64_t ell_chunk = 1;       // consecutive 256-row tiles per CTA visit (ell.cu:TileWalk)
    int64_t ell_tma = 2;         // coded ELL SpMV fed by TMA bulk copies through a shared-memory ring (ell.cu:k_spmv_ell_tma):
                                 // 0 off, 1 on, 2 auto = above 2 M rows (measured: 88 vs 93 us at 8 M rows, 30.5 vs 30.2 at 1 M)
    int64_t ell_minb = 4;        // resident CTAs per SM the coded ELL SpMV is compiled for (3: 85 registers, 4: 64)
    int64_t ell_minb_cgp = 3;    // ... and the fused CG kernel (2, 3 or 4)
    int64_t fuse_p = 0;          // CG: p-update fused into the ELL SpMV (ell.cu:k_spmv_ell_cgp); measured slower than
                                 // the separate k_cg_p at 1 M and 8 M rows on 1 and 2 GPUs (profiles/r02_*probe*), so off
    int64_t ell_auto = 1;        // 1: spmv_variant 0 may pick the ELL kernels (spmv.cu:pick_variant)
    // merge-path SpMV (spmv_merge.cu): first row starting at or behind every 2048-entry slice, carries
    int mp_chunks = 0;
    label *d_mp_chunk_row = nullptr, *d_mp_carry_row = nullptr;
    double *d_mp_carry_val = nullptr;
    int64_t max_block_nnz = 0;   // stream kernel: max nnz of a kRowsPerBlock row block
    int64_t max_warp_nnz = 0;    // warp-tile kernel: max nnz of 32 consecutive rows

    // partition (a6/a11)
    bool have_partition = false;
    int64_t global_n = 0;
    label n_targets = 0, n_send = 0;
    std::vector<label> target_ids, target_sizes, send_offs;
    label *d_send_idxs = nullptr;
    double *d_send_buf = nullptr, *d_recv_buf = nullptr;

    // peer-memory window (multi-GPU P2P path, comm.cu)
    int64_t comm_mode = 0;        // 0 auto, 1 NCCL send/recv + allreduce, 2 peer-memory (P2P)
    int64_t fused_halo = 2;       // P2P: non-local block inside the stream SpMV (0 off, 1 on, 2 auto)
    bool p2p_ready = false;
    void *d_window = nullptr;     // my window (exported through CUDA IPC)
    size_t window_bytes = 0;
    std::vector<void *> peer_windows;   // opened IPC mappings, by rank (nullptr for self)
    struct CommDev *d_commdev = nullptr;

    // non-local pattern (a7)
    label n_halo = 0;
    bool have_nonlocal = false;
    label *d_nl_rows = nullptr, *d_nl_cols = nullptr, *d_nl_map = nullptr;
    // CSR-like grouping of the non-local entries by row (rows touching the halo)
    label n_nl_rows = 0;
    label *d_nl_row_ids = nullptr, *d_nl_row_ptrs = nullptr;
    // CG on several GPUs, "ghost p" mode (solver.cu): boundary z pushed by the x/r-update
    // kernel, p kept with n + n_halo entries; per-CTA lists of the send entries a CTA owns
    int64_t trace = 0;           // 1: kernels log (tag, globaltimer) events into d_trace
    unsigned long long *d_trace = nullptr;
    int64_t ghost_p = 1;
    double **d_push_dst = nullptr;   // [n_send] slot 2 of the neighbour's window, per send entry
    int64_t fused_pcg = 2;       // CG loop as one persistent cooperative kernel (pcg_fused.cu): 0 off, 1 on,
                                 // 2 auto = on for small systems, where launch latencies dominate
    unsigned int *d_bar = nullptr;   // its grid barrier: arrivals, generation
    size_t work_len = 0;
    // ghosted CSR (multi-GPU, halo-fused SpMV): every row = its local entries followed by
    // its non-local ones, whose columns are n + index into the receive window
    label *d_g_row_ptrs = nullptr, *d_g_cols = nullptr, *d_g_map = nullptr;
    double *d_g_vals = nullptr;
    int64_t max_block_nnz_g = 0;
    bool have_ghosted = false;

    // values (a8/a9)
    bool have_values = false;
    double *d_staging = nullptr;   // [upper | lower | diag | iface] as uploaded
    size_t staging_len = 0;
    double *d_vals = nullptr, *d_nl_vals = nullptr, *d_nl_staging = nullptr;

    // vectors (a13) + solver workspace
    double *d_b = nullptr, *d_x = nullptr;
    bool have_b = false, have_x = false;
    std::vector<double *> work;   // n-sized work vectors, allocated on demand
    double *d_krylov = nullptr;   // GMRES basis (krylov_dim+1) * n
    int64_t krylov_cap = 0;
    double *d_hess = nullptr;     // GMRES small dense state
    int64_t hess_cap = 0;

    // preconditioner (a14)
    int precond_kind = OGL_PRECOND_NONE;
    label max_block_size = 1;
    label bj_pattern_mbs = 0;    // block pointers were built for this maxBlockSize
    bool bj_uniform = false;     // every block has exactly maxBlockSize rows (the last one may be shorter)
    bool have_precond = false;
    double *d_inv_diag = nullptr;
    label n_blocks = 0;
    label *d_block_ptrs = nullptr, *d_row_block = nullptr;
    int64_t *d_block_offs = nullptr;
    double *d_inv_blocks = nullptr;
    int64_t inv_blocks_len = 0;
    // ISAI / GISAI: approximate-inverse values over the CSR pattern of the local matrix (w), and
    // for the spd variant the transpose (wt); z = W r  or  z = W^T (W r)
    double *d_isai_w = nullptr, *d_isai_wt = nullptr;
    // ILU / IC / IRILU (trifactor.cu): incomplete factors over the CSR pattern of the local matrix
    // (strictly lower part = L, upper part = U with the diagonal; IC: L with its diagonal, mirrored
    // into the upper part), rows grouped into dependency levels for the two triangular sweeps
    struct TriFactor {
        bool structure_ready = false;   // diag_pos / perm_* / levels match the current pattern
        int64_t nnz = 0;                // entries `vals` was sized for
        label *diag_pos = nullptr;      // [n] position of (i,i) in the CSR
        label *perm_l = nullptr, *perm_u = nullptr;   // [n] rows in level order (forward / backward sweep)
        std::vector<label> lvl_l, lvl_u;              // level offsets into perm_* (host copy, one launch per level)
        double *vals = nullptr;         // [nnz]
        double *dval = nullptr;         // [n] the factor's diagonal (vals[diag_pos[i]]), contiguous
        label dval_n = 0;
        int max_lower = 0, max_upper = 0;   // longest strictly lower / upper part of a row
        int sf_grid = 0, sf_per_sm = 1;     // co-resident grid of the dependency-driven sweeps
    } tri;
    // Multigrid (multigrid.cu): the hierarchy of the local matrix.  Level 0 aliases the context's CSR
    // (and uses the caller's vectors); every coarser level owns its matrix and vectors.
    struct MgLevel {
        label n = 0;
        int64_t nnz = 0;
        label *rp = nullptr, *cols = nullptr, *rows = nullptr;
        double *vals = nullptr;
        double *inv_diag = nullptr;
        label *agg = nullptr;           // [n] fine row -> coarse row (nullptr on the coarsest level)
        label n_coarse = 0;
        label *r_ptr = nullptr, *r_idx = nullptr;   // restriction: members of every aggregate, ascending
        double *b = nullptr, *x = nullptr, *r = nullptr;   // right-hand side, correction, residual / scratch
        double *p = nullptr, *q = nullptr;                 // coarsest level: CG vectors
    };
    struct Multigrid {
        std::vector<MgLevel> levels;
        double *scal = nullptr, *partials = nullptr;   // coarsest CG: rho, prev_rho, beta; dot-product partials
        bool ready = false;
    } mg;
    int64_t mg_max_levels = 9, mg_min_coarse_rows = 10, mg_coarse_iters = 4;   // Preconditioner.H:297,317-320
    int64_t tri_sleep_ns = 0;     // dependency-driven sweep: pause of a waiting warp between two rounds of polls
    int64_t tri_ctas = 0;         // ... and its CTAs per SM (0 = as many as fit)
    int64_t tri_variant = 1;   // triangular sweeps: 0 one launch per dependency level, 1 ONE launch per sweep whose
                               // rows wait for the entries they depend on (trifactor.cu:k_trisolve_sf)

    // reduction scratch
    double *d_partials = nullptr;          // kMaxPartialBlocks * kMaxReduce
    unsigned int *d_ticket = nullptr;
    SolveState *d_state = nullptr;
    SolveState *h_state = nullptr;         // pinned: [0] final, [1..2] polling slots
    double *d_history = nullptr;
    int history_cap = 0;

    // host staging for pageable uploads
    void *h_pinned = nullptr;
    size_t h_pinned_bytes = 0;

    // CUDA graph cache for one chunk of iterations
    cudaGraphExec_t graph_exec = nullptr;
    int graph_solver = -1;
    int64_t graph_sig = 0;
    int64_t graph_kernels = 0;   // kernels inside one replayed chunk
    int64_t device_loop = 1;     // CG: chunk graph = body of a CUDA-graph WHILE node whose condition the
                                 // criterion epilogue clears on the device (one launch per solve)
    int64_t loop_iters = 16;     // iterations per loop body (even): ~3.2 us per body of loop overhead vs
                                 // up to loop_iters-1 early-exit iterations after the criterion fired
    int64_t loop_body_iters = 0; // > 0: a loop graph was launched by this solve
    bool graph_is_loop = false;
    unsigned long long cond_handle = 0;   // cudaGraphConditionalHandle while the loop body is captured

    int64_t gmres_persist = 1;   // GMRES: MGS sweep of an Arnoldi step as one persistent cooperative kernel (<= ~0.6 M rows)
    int64_t launches = 0;
    int64_t precond_setups = 0;   // ogl_precond_setup calls (tests of the `caching` keyword)

    // sampled SpMV timing inside the solve loop (profile_stride > 0)
    std::vector<cudaEvent_t> profile_events;
    int profile_used = 0;
    int64_t profile_iter = 0;
    bool capturing = false;
};

// Launch with (optional) programmatic stream serialisation: the next kernel's
// CTAs may become resident while this one drains (its serial reduction tail);
// kernels call cudaGridDependencySynchronize() before touching anything the
// previous kernel wrote.  Captured into the chunk graph as programmatic edges.
template <typename K, typename A>
inline cudaError_t launch_pdl(K kernel, int grid, int block, size_t smem, cudaStream_t st, bool pdl,
                              const A &arg)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}

// memory helpers ---------------------------------------------------------------
void invalidate_graph(Context *ctx);

// (Re)allocate a device buffer.  Replacing a live buffer drops the cached chunk graph: its
// kernel nodes hold the old address (residual history, block-Jacobi arrays, halo patterns ...).
template <typename T>
int dev_alloc(Context *ctx, T **p, size_t count)
{
    if (*p) {
        invalidate_graph(ctx);
        cudaFree(*p);
        *p = nullptr;
    }
    if (count == 0) count = 1;
    OGL_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T) + 64));
    return OGL_OK;
}

int ensure_pinned(Context *ctx, size_t bytes);
int upload(Context *ctx, void *dst, const void *src, size_t bytes);
int download(Context *ctx, void *dst, const void *src, size_t bytes);
int get_work(Context *ctx, int idx, double **out);

// assembly.cu ------------------------------------------------------------------
int pattern_from_ldu(Context *ctx, label n, label nf, bool sym, const label *lower,
                     const label *upper, label n_if, const label *if_rows,
                     const label *if_cols);
int nonlocal_pattern(Context *ctx, label n_halo, const label *face_cells);
int values_update(Context *ctx, const double *diag, const double *upper,
                  const double *lower, const double *if_bou, const double *nl_bou,
                  double scaling);

// spmv.cu ------------------------------------------------------------------------
// y = A_local x (optionally y = alpha*A*x + beta*y), optional fused partial
// reduction red[0] = <dot_with, y_local_part> left in ctx->d_state->red[0].
struct SpmvArgs {
    const double *x = nullptr;
    double *y = nullptr;
    const double *y_in = nullptr;       // advanced: y = alpha*A*x + beta*y_in (default y)
    bool advanced = false;
    double alpha = 1.0, beta = 0.0;
    const double *dot_with = nullptr;   // nred >= 1: red[0] = <dot_with, y>
    int nred = 0;                       // nred == 2: red[1] = <y, y>
    bool guard_done = false;            // early-exit when state->done
    int epi = 0;                        // scalar epilogue after the reduction (reduce.cuh)
    bool inline_epi = true;             // run it inside the kernel (single rank)
    bool fused_halo = false;            // stream kernel applies the non-local block itself (P2P)
    bool halo_stored = false;           // boundary values already stored by the previous kernel
    bool ghost_x = false;               // x has n + n_halo entries, the ghost part already filled:
                                        // ghosted CSR, no flag handshake (CG ghost-p mode)
    const double *vals_override = nullptr;   // other values over the LOCAL CSR pattern (ISAI apply)
    int ar_count = 0;                   // vals_override on several ranks: red[] slots all-reduced in the launch
};
int spmv_local(Context *ctx, const SpmvArgs &a);
int spmv_nonlocal(Context *ctx, const double *recv, double *y, double alpha,
                  const double *dot_with, int nred, bool guard_done, int epi,
                  bool inline_epi);
int spmv_setup(Context *ctx);
int spmv_merge_setup(Context *ctx);   // spmv_merge.cu
int spmv_ell_cgp(Context *ctx, const double *z, const double *p_old, double *p_new, double *q, bool ghost);
void ell_invalidate(Context *ctx, bool structure);
int ell_prepare_for_loop(Context *ctx, bool ghost);

// comm.cu ------------------------------------------------------------------------
int partition_create(Context *ctx, label n_local, label n_targets, const label *target_ids,
                     const label *target_sizes, const label *send_idxs);
int partition_export(Context *ctx, void *blob, int64_t capacity, int64_t *size);
int partition_connect(Context *ctx, const void *blobs, int64_t n_blobs);
int halo_begin(Context *ctx, const double *x, bool guard_done);   // pack (+ NCCL send/recv)
int halo_end(Context *ctx);                                       // compute stream waits for recv
int allreduce_red(Context *ctx, int count);                      // state->red[0..count) summed over ranks
int dist_spmv(Context *ctx, const SpmvArgs &a);                   // halo + local + non-local

// precond.cu ---------------------------------------------------------------------
int precond_setup(Context *ctx, int kind, label mbs);
int precond_apply(Context *ctx, const double *r, double *z, const double *dot_with,
                  int red_base, bool guard_done, int epi, bool inline_epi, int ar_count);
// trifactor.cu -------------------------------------------------------------------
inline bool is_tri_precond(int kind)
{
    return kind == OGL_PRECOND_ILU || kind == OGL_PRECOND_IC || kind == OGL_PRECOND_IRILU;
}
int tri_setup(Context *ctx, int kind);                 // (analysis once per pattern) + factorisation
int tri_ensure_structure(Context *ctx);                // re-analyse after a pattern rebuild (cached factors)
int tri_apply(Context *ctx, const double *r, double *z, bool guard_done);
void tri_release(Context *ctx);
// multigrid.cu -------------------------------------------------------------------
int mg_setup(Context *ctx);
int mg_ensure(Context *ctx);
int mg_apply(Context *ctx, const double *r, double *z, bool guard_done);
void mg_release(Context *ctx);
int mg_level_info(Context *ctx, int level, label *n, label *nnz, label *n_coarse);
int mg_level_download(Context *ctx, int level, label *rp, label *cols, double *vals, label *agg);
bool use_p2p(const Context *ctx);
void comm_teardown(Context *ctx);
int comm_bench(Context *ctx, int mode, int reps, double *us);
int spmv_variant_in_use(const Context *ctx);
int l2_keep_level(const Context *ctx);
bool fused_halo_ok(const Context *ctx);
int pack_stores(Context *ctx, const double *x, bool guard_done);
int push_boundary(Context *ctx, const double *v);
bool pcg_fused_ok(const Context *ctx);
int pcg_fused_run(Context *ctx, double *r0, double *r1, double *z, double *p0, double *p1,
                  double *q, int64_t max_iter);
unsigned long long spmv_l2_policy(Context *ctx);

// solver.cu ----------------------------------------------------------------------
int solve(Context *ctx, const ogl_solve_params *p, ogl_solve_result *res);
int finish_reduction(Context *ctx, int count, int epi, bool guard);
int vec_fill(Context *ctx, double *v, double value);
int vec_scale(Context *ctx, double *v, double s);
int init_state(Context *ctx, const ogl_solve_params *p);
int solve_prologue(Context *ctx, int mode, double *r, double *z, double *rr, double *w,
                   double *tmp, int epi);

}  // namespace ogl

struct ogl_ctx : public ogl::Context {};
