// Multi-GPU plumbing: one process (rank) per GPU.
//
// Replaces, on this path, Ginkgo's distributed::Matrix::communicate
// (MPI_Ineighbor_alltoallv of gathered boundary values, overlapped with the
// local SpMV) and the MPI_Allreduce behind every distributed dot / norm.  The
// partition (who sends what to whom) is OGL's create_communication_pattern
// (HostMatrix/HostMatrix.C:251-306) fed through PartitionInitFunctor
// (DevicePersistent/Partition/Partition.H:57-70).
//
// Two data paths:
//   comm_mode 2 (default when every peer is reachable over NVLink P2P):
//     peer-memory windows (reduce.cuh).  Reducing kernels all-reduce their
//     partial sums through the peers' mailboxes in their own last block
//     (self-validating stamped words).  Boundary values: the CG loop pushes the
//     boundary z from its x/r-update kernel and keeps ghost entries of p itself
//     (solver.cu "ghost p"); every other distributed SpMV stores x[send_idxs]
//     into the neighbours' receive buffers and shakes hands with flags.
//     No NCCL call, no extra launch per iteration.
//   comm_mode 1: NCCL -- grouped ncclSend/ncclRecv on a side stream overlapped
//     with the local SpMV, ncclAllReduce of the packed scalars in place in the
//     device-resident SolveState, then a one-thread epilogue kernel.
// NCCL is also what bootstraps the windows (all-gather of the IPC handles and
// of each rank's neighbour directory) and sums the global size.
#include <cstring>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

namespace {

// send_buf[k] = x[send_idxs[k]]  (Ginkgo row_gather by the partition's send indices)
__global__ void k_pack(label n_send, const label *__restrict__ idx,
                       const double *__restrict__ x, double *__restrict__ buf)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_send) buf[k] = x[idx[k]];
}

// P2P pack: x[send_idxs[k]] goes directly into the receive buffer of the
// neighbour that owns block k; the last block raises the neighbours' data flags.
__global__ void __launch_bounds__(256)
k_pack_p2p(label n_send, const label *__restrict__ idx, const double *__restrict__ x,
           CommDev *c, SolveState *state, int guard_done)
{
    if (guard_done && state->done) return;
    __shared__ bool is_last;
    const unsigned long long seq = c->halo_seq + 1;
    const int parity = (int)(seq & 1ull);
    // flow control: the buffer of this parity was last used by exchange seq-2;
    // wait until every neighbour has acknowledged consuming it
    if (seq > 2 && threadIdx.x < c->n_targets) {
        if (!wait_flag(&c->my_ack_flag[threadIdx.x], seq - 2)) state->comm_error = 1;
    }
    __syncthreads();
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_send) {
        int t = 0;
        while (k >= c->send_offs[t + 1]) ++t;
        double *dst = c->peer_recv[t] + (size_t)parity * c->peer_recv_stride[t] + (k - c->send_offs[t]);
        const double v = x[idx[k]];
        asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
    }
    // one system-scope fence per block (cumulative over the block's stores after
    // the barrier) instead of one per thread: membar.sys is the expensive part
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int tk = atomicAdd(&c->pack_ticket, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x < c->n_targets) st_flag(c->peer_data_flag[threadIdx.x], seq);
    if (threadIdx.x == 0) {
        c->halo_seq = seq;
        c->pack_ticket = 0u;
    }
}

// Stores only: the consumer kernel (halo-fused stream SpMV) publishes the
// exchange on entry -- the kernel boundary orders these stores before its flag.
__global__ void __launch_bounds__(256)
k_pack_stores(label n_send, const label *__restrict__ idx, const double *__restrict__ x,
              CommDev *c, SolveState *state, int guard_done)
{
    if (guard_done && state->done) return;
    const unsigned long long seq = c->halo_seq + 1;
    const int parity = (int)(seq & 1ull);
    if (seq > 2 && threadIdx.x < c->n_targets) {
        if (!wait_flag(&c->my_ack_flag[threadIdx.x], seq - 2)) state->comm_error = 1;
    }
    __syncthreads();
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_send) {
        int t = 0;
        while (k >= c->send_offs[t + 1]) ++t;
        double *dst = c->peer_recv[t] + (size_t)parity * c->peer_recv_stride[t] + (k - c->send_offs[t]);
        *dst = x[idx[k]];   // weak, coalesced per warp; published by the next kernel's release
    }
}

// CG ghost-p mode, before the first iteration: boundary values of v into slot 2
// of the neighbours' windows (afterwards the x/r-update kernel does this itself)
__global__ void __launch_bounds__(256)
k_push_boundary(label n_send, const label *__restrict__ idx, const double *__restrict__ v,
                double *const *__restrict__ dst, const CommDev *c)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_send)
        push_stamped(reinterpret_cast<unsigned long long *>(dst[k]), v[idx[k]], stamp_of(ld_ar_seq(c) + 1));
}

// rendezvous of all ranks (an all-reduce of nothing): keeps the rule "a push is
// stamped with the number of the all-reduce that follows it" for the first push
__global__ void k_rank_barrier(SolveState *state, CommDev *c) { p2p_allreduce(state, 0, c); }

constexpr int kDirectoryMagic = 0x4f474c44;   // "OGLD"

struct Directory {
    int magic, pad;
    long long n_local;
    int n_targets;
    int n_halo;
    int target_ids[kMaxTargets];
    int send_offs[kMaxTargets + 1];
    long long off_mbox, off_data_flag, off_ack_flag, off_recv;   // bytes inside the window
    cudaIpcMemHandle_t handle;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

static void p2p_teardown(Context *ctx)
{
    for (void *p : ctx->peer_windows)
        if (p) cudaIpcCloseMemHandle(p);
    ctx->peer_windows.clear();
    if (ctx->d_window) cudaFree(ctx->d_window);
    ctx->d_window = nullptr;
    if (ctx->d_commdev) cudaFree(ctx->d_commdev);
    ctx->d_commdev = nullptr;
    ctx->p2p_ready = false;
}

void comm_teardown(Context *ctx) { p2p_teardown(ctx); }

// Step 1 of the window bootstrap (rank-local): allocate and zero this rank's window, describe it.
// Returns false when the allocation / IPC export fails (the caller agrees with its peers on that).
static bool p2p_make_window(Context *ctx, Directory &mine)
{
    const int R = ctx->n_ranks;
    std::memset(&mine, 0, sizeof(mine));
    mine.magic = kDirectoryMagic;
    mine.n_local = ctx->n;
    mine.n_targets = ctx->n_targets;
    mine.n_halo = ctx->n_send;
    for (int t = 0; t < ctx->n_targets; ++t) mine.target_ids[t] = ctx->target_ids[t];
    for (int t = 0; t <= ctx->n_targets; ++t) mine.send_offs[t] = ctx->send_offs[t];
    size_t off = 0;
    mine.off_mbox = (long long)off;
    off = align_up(off + sizeof(unsigned long long) * 2 * R * kSlot, 256);
    mine.off_data_flag = (long long)off;
    off = align_up(off + sizeof(unsigned long long) * kMaxTargets, 256);
    mine.off_ack_flag = (long long)off;
    off = align_up(off + sizeof(unsigned long long) * kMaxTargets, 256);
    mine.off_recv = (long long)off;
    // parity 0/1 + slot 2: boundary z of the CG ghost-p mode, 16 B per (stamped) value
    off = align_up(off + sizeof(double) * 4 * (size_t)(ctx->n_send > 0 ? ctx->n_send : 1), 256);
    ctx->window_bytes = off;
    bool ok = cudaMalloc(&ctx->d_window, off) == cudaSuccess;
    if (ok) ok = cudaMemset(ctx->d_window, 0, off) == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&mine.handle, ctx->d_window) == cudaSuccess;
    cudaGetLastError();
    return ok;
}

// Step 2a (rank-local): map the peers' windows.  false when a peer is not reachable.
static bool p2p_map_peers(Context *ctx, const Directory *all)
{
    const int R = ctx->n_ranks;
    ctx->peer_windows.assign(R, nullptr);
    for (int q = 0; q < R; ++q) {
        if (q == ctx->rank) continue;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        ctx->peer_windows[q] = p;
    }
    return true;
}

// Step 2b (rank-local): the device-side description of the windows.  *sym_ok = 0 when the
// neighbour lists of the ranks do not match (the caller agrees with its peers on the verdict).
static int p2p_describe(Context *ctx, const Directory *all, int *sym_ok)
{
    const int R = ctx->n_ranks;
    const Directory &mine = all[ctx->rank];
    CommDev h;
    std::memset(&h, 0, sizeof(h));
    h.rank = ctx->rank;
    h.n_ranks = R;
    h.n_targets = ctx->n_targets;
    auto base_of = [&](int q) -> char * {
        return static_cast<char *>(q == ctx->rank ? ctx->d_window : ctx->peer_windows[q]);
    };
    for (int q = 0; q < R; ++q) h.mbox[q] = reinterpret_cast<double *>(base_of(q) + all[q].off_mbox);
    for (int t = 0; t <= ctx->n_targets; ++t) h.send_offs[t] = ctx->send_offs[t];
    std::vector<long long> peer_block_off((size_t)ctx->n_targets, 0);   // my block inside q's recv buffer
    *sym_ok = 1;
    for (int t = 0; t < ctx->n_targets; ++t) {
        const int q = ctx->target_ids[t];
        const Directory &dq = all[q];
        int u = -1;
        for (int k = 0; k < dq.n_targets; ++k)
            if (dq.target_ids[k] == ctx->rank) u = k;
        if (u < 0 || dq.send_offs[u + 1] - dq.send_offs[u] != ctx->target_sizes[t]) {
            // no rank-local return inside a collective sequence: the verdict is agreed on by
            // the caller, so that every rank leaves together
            *sym_ok = 0;
            continue;
        }
        // my values land in the neighbour's recv block reserved for me: same offset
        // as the neighbour's own send block towards me (blocked by ascending rank)
        h.peer_recv[t] = reinterpret_cast<double *>(base_of(q) + dq.off_recv) + dq.send_offs[u];
        peer_block_off[(size_t)t] = dq.send_offs[u];
        h.peer_recv_stride[t] = dq.n_halo > 0 ? dq.n_halo : 1;
        h.peer_data_flag[t] = reinterpret_cast<unsigned long long *>(base_of(q) + dq.off_data_flag) + u;
        h.peer_ack_flag[t] = reinterpret_cast<unsigned long long *>(base_of(q) + dq.off_ack_flag) + u;
    }
    char *me = static_cast<char *>(ctx->d_window);
    h.my_data_flag = reinterpret_cast<unsigned long long *>(me + mine.off_data_flag);
    h.my_ack_flag = reinterpret_cast<unsigned long long *>(me + mine.off_ack_flag);
    h.my_recv = reinterpret_cast<double *>(me + mine.off_recv);
    h.my_recv_stride = ctx->n_send > 0 ? ctx->n_send : 1;
    OGL_TRY(dev_alloc(ctx, &ctx->d_commdev, 1));
    OGL_CUDA(ctx, cudaMemcpy(ctx->d_commdev, &h, sizeof(h), cudaMemcpyHostToDevice));
    // CG ghost-p mode: destination of every send entry in slot 2 of its neighbour's window
    std::vector<double *> dst((size_t)ctx->n_send, nullptr);
    for (label t = 0; t < ctx->n_targets; ++t)
        for (label k = ctx->send_offs[t]; k < ctx->send_offs[t + 1]; ++k)
            // slot 2 holds 16 B per entry: my block starts at 2 * (its offset in q's buffer);
            // peer_recv[t] already contains that offset once
            dst[(size_t)k] = h.peer_recv[t] + 2 * h.peer_recv_stride[t] + peer_block_off[(size_t)t] +
                             2 * (size_t)(k - ctx->send_offs[t]);
    OGL_TRY(dev_alloc(ctx, &ctx->d_push_dst, (size_t)ctx->n_send));
    OGL_CUDA(ctx, cudaMemcpy(ctx->d_push_dst, dst.data(), sizeof(double *) * dst.size(),
                             cudaMemcpyHostToDevice));
    return OGL_OK;
}

// Build the peer-memory windows with NCCL as the bootstrap channel.  Collective.  Returns OGL_OK
// with ctx->p2p_ready == false when the topology does not allow it (caller falls back to NCCL).
static int p2p_setup(Context *ctx)
{
    p2p_teardown(ctx);
    if (ctx->n_ranks > kMaxPeers || ctx->n_targets > kMaxTargets) return OGL_OK;
    const int R = ctx->n_ranks;
    Directory mine;
    int ok_local = p2p_make_window(ctx, mine) ? 1 : 0;
    // ---- all-gather the directories (and whether everyone got this far)
    Directory *d_all = nullptr;
    OGL_TRY(dev_alloc(ctx, &d_all, (size_t)R));
    int *d_ok = nullptr;
    OGL_TRY(dev_alloc(ctx, &d_ok, 1));
    cudaMemcpyAsync(d_all + ctx->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_ok, &ok_local, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r1 = ncclAllGather(d_all + ctx->rank, d_all, sizeof(Directory), ncclChar, ctx->comm,
                                    ctx->stream);
    ncclResult_t r2 = ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, ctx->comm, ctx->stream);
    std::vector<Directory> all(R);
    int ok_all = 0;
    cudaMemcpyAsync(all.data(), d_all, sizeof(Directory) * R, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&ok_all, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_all);
    cudaFree(d_ok);
    if (r1 != ncclSuccess || r2 != ncclSuccess)
        return fail(ctx, OGL_ERR_NCCL, "window bootstrap: NCCL all-gather failed");
    if (e != cudaSuccess)
        return fail(ctx, OGL_ERR_CUDA, std::string("window bootstrap: ") + cudaGetErrorString(e));
    // ---- map the peers' windows; everyone must agree before the first kernel relies on it
    int mapped = ok_all && p2p_map_peers(ctx, all.data()) ? 1 : 0;
    int *d_flag = nullptr;
    OGL_TRY(dev_alloc(ctx, &d_flag, 1));
    cudaMemcpyAsync(d_flag, &mapped, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r3 = ncclAllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, ctx->comm, ctx->stream);
    int mapped_all = 0;
    cudaMemcpyAsync(&mapped_all, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_flag);
    if (r3 != ncclSuccess || e != cudaSuccess)
        return fail(ctx, OGL_ERR_NCCL, "window bootstrap: agreement all-reduce failed");
    if (!mapped_all) {
        p2p_teardown(ctx);
        return OGL_OK;   // NCCL path
    }
    int sym_ok = 1;
    OGL_TRY(p2p_describe(ctx, all.data(), &sym_ok));
    // nobody may touch a window before every rank has finished zeroing/mapping; the same
    // all-reduce carries the verdict on the neighbour lists
    int *d_bar = nullptr;
    OGL_TRY(dev_alloc(ctx, &d_bar, 1));
    cudaMemcpyAsync(d_bar, &sym_ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r4 = ncclAllReduce(d_bar, d_bar, 1, ncclInt, ncclMin, ctx->comm, ctx->stream);
    int sym_all = 0;
    cudaMemcpyAsync(&sym_all, d_bar, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bar);
    if (r4 != ncclSuccess || e != cudaSuccess)
        return fail(ctx, OGL_ERR_NCCL, "window bootstrap: final barrier failed");
    if (!sym_all) {
        p2p_teardown(ctx);
        return fail(ctx, OGL_ERR_INVALID,
                    "neighbour lists of the ranks are not symmetric (processor patches)");
    }
    ctx->p2p_ready = true;
    return OGL_OK;
}

// ---- host-driven bootstrap (no NCCL): the caller moves the directories with whatever
// communicator it has (MPI in an OpenFOAM run, gloo in the tests).  This is also what lets several
// ranks share ONE device (NCCL refuses duplicate GPUs; CUDA IPC does not care).
int partition_export(Context *ctx, void *blob, int64_t capacity, int64_t *size)
{
    if (size) *size = (int64_t)sizeof(Directory);
    if (!blob) return OGL_OK;   // size query
    if (capacity < (int64_t)sizeof(Directory)) return fail(ctx, OGL_ERR_INVALID, "directory buffer too small");
    if (!ctx->have_partition) return fail(ctx, OGL_ERR_INVALID, "ogl_partition_export before ogl_partition_create");
    if (ctx->n_ranks > kMaxPeers || ctx->n_targets > kMaxTargets)
        return fail(ctx, OGL_ERR_UNSUPPORTED, "too many ranks / neighbours for the peer-memory windows");
    p2p_teardown(ctx);
    Directory mine;
    if (!p2p_make_window(ctx, mine)) {
        p2p_teardown(ctx);
        return fail(ctx, OGL_ERR_CUDA, "could not allocate / export the peer-memory window");
    }
    std::memcpy(blob, &mine, sizeof(mine));
    return OGL_OK;
}

int partition_connect(Context *ctx, const void *blobs, int64_t n_blobs)
{
    if (!blobs || n_blobs != ctx->n_ranks) return fail(ctx, OGL_ERR_INVALID, "need one directory per rank");
    if (!ctx->d_window) return fail(ctx, OGL_ERR_INVALID, "ogl_partition_connect before ogl_partition_export");
    std::vector<Directory> all((size_t)ctx->n_ranks);
    std::memcpy(all.data(), blobs, sizeof(Directory) * all.size());
    long long global_n = 0;
    for (const Directory &d : all) {
        if (d.magic != kDirectoryMagic) return fail(ctx, OGL_ERR_INVALID, "bad directory blob");
        global_n += d.n_local;
    }
    if (!p2p_map_peers(ctx, all.data())) {
        p2p_teardown(ctx);
        return fail(ctx, OGL_ERR_CUDA, "a peer's window cannot be mapped (no P2P path between the devices)");
    }
    int sym_ok = 1;
    OGL_TRY(p2p_describe(ctx, all.data(), &sym_ok));
    if (!sym_ok) {
        p2p_teardown(ctx);
        return fail(ctx, OGL_ERR_INVALID, "neighbour lists of the ranks are not symmetric (processor patches)");
    }
    ctx->global_n = global_n;
    ctx->p2p_ready = true;   // the caller runs a barrier of its own before the first solve
    invalidate_graph(ctx);
    return OGL_OK;
}

namespace {

// reps all-reduces inside ONE launch: pure device-side latency of the primitive
__global__ void k_ar_loop(SolveState *state, CommDev *c, int reps)
{
    for (int i = 0; i < reps; ++i) {
        if (threadIdx.x == 0) state->red[0] = 1.0;
        __syncthreads();
        p2p_allreduce(state, 1, c);
    }
}

// one all-reduce per launch
__global__ void k_ar_once(SolveState *state, CommDev *c)
{
    if (threadIdx.x == 0) state->red[0] = 1.0;
    __syncthreads();
    p2p_allreduce(state, 1, c);
}

}  // namespace

int comm_bench(Context *ctx, int mode, int reps, double *us)
{
    if (!use_p2p(ctx)) return fail(ctx, OGL_ERR_INVALID, "comm_bench needs the peer-memory path");
    cudaStream_t st = ctx->stream;
    double *x = nullptr;
    OGL_TRY(get_work(ctx, 2, &x));
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) cudaEventRecord(ctx->ev_t0, st);
        const int n = pass == 0 ? 3 : reps;
        if (mode == 0) {
            for (int i = 0; i < n; ++i) k_ar_once<<<1, 32, 0, st>>>(ctx->d_state, ctx->d_commdev);
        } else if (mode == 1) {
            k_ar_loop<<<1, 32, 0, st>>>(ctx->d_state, ctx->d_commdev, n);
        } else {
            for (int i = 0; i < n; ++i) {
                OGL_TRY(halo_begin(ctx, x, false));
                OGL_TRY(spmv_nonlocal(ctx, ctx->d_recv_buf, x, 0.0, nullptr, 0, false, EPI_NONE, true));
            }
        }
    }
    cudaEventRecord(ctx->ev_t1, st);
    OGL_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1);
    *us = ms * 1e3 / reps;
    return OGL_OK;
}

// the halo-fused stream kernels need the peer-memory path and SpMV variant 1 or 6
bool fused_halo_ok(const Context *ctx)
{
    const int v = spmv_variant_in_use(ctx);
    if (!use_p2p(ctx) || !(v == 1 || v == 6 || v == 7) || ctx->fused_halo == 0) return false;
    // With the ghosted CSR (assembly.cu:build_ghosted) the halo costs the kernel
    // nothing per tile, so auto (2) == on (1).
    return true;
}

int push_boundary(Context *ctx, const double *v)
{
    if (!use_p2p(ctx)) return fail(ctx, OGL_ERR_INVALID, "push_boundary needs the peer-memory path");
    if (ctx->n_send > 0) {
        k_push_boundary<<<(ctx->n_send + 255) / 256, 256, 0, ctx->stream>>>(
            ctx->n_send, ctx->d_send_idxs, v, ctx->d_push_dst, ctx->d_commdev);
        ctx->launches++;
    }
    k_rank_barrier<<<1, 32, 0, ctx->stream>>>(ctx->d_state, ctx->d_commdev);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

int pack_stores(Context *ctx, const double *x, bool guard_done)
{
    if (ctx->n_send == 0) return OGL_OK;
    k_pack_stores<<<(ctx->n_send + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->n_send, ctx->d_send_idxs, x, ctx->d_commdev, ctx->d_state, guard_done ? 1 : 0);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

bool use_p2p(const Context *ctx)
{
    return ctx->n_ranks > 1 && ctx->p2p_ready && ctx->comm_mode != 1;
}

int partition_create(Context *ctx, label n_local, label n_targets, const label *target_ids,
                     const label *target_sizes, const label *send_idxs)
{
    if (!ctx->have_pattern)
        return fail(ctx, OGL_ERR_INVALID, "ogl_partition_create before ogl_pattern_from_ldu");
    if (n_local != ctx->n)
        return fail(ctx, OGL_ERR_INVALID, "partition local size differs from the matrix rows");
    if (n_targets < 0 || (n_targets > 0 && (!target_ids || !target_sizes)))
        return fail(ctx, OGL_ERR_INVALID, "bad partition arguments");
    ctx->target_ids.assign(target_ids, target_ids + n_targets);
    ctx->target_sizes.assign(target_sizes, target_sizes + n_targets);
    ctx->send_offs.assign(n_targets + 1, 0);
    for (label t = 0; t < n_targets; ++t) {
        if (target_ids[t] < 0 || target_ids[t] >= ctx->n_ranks || target_ids[t] == ctx->rank)
            return fail(ctx, OGL_ERR_INVALID, "partition target rank out of range");
        if (t > 0 && target_ids[t] <= target_ids[t - 1])
            return fail(ctx, OGL_ERR_INVALID, "partition targets must be ascending and unique");
        if (target_sizes[t] < 0) return fail(ctx, OGL_ERR_INVALID, "negative target size");
        ctx->send_offs[t + 1] = ctx->send_offs[t] + target_sizes[t];
    }
    ctx->n_targets = n_targets;
    ctx->n_send = ctx->send_offs[n_targets];
    if (ctx->n_send > 0 && !send_idxs) return fail(ctx, OGL_ERR_INVALID, "null send_idxs");
    for (label k = 0; k < ctx->n_send; ++k)
        if (send_idxs[k] < 0 || send_idxs[k] >= ctx->n)
            return fail(ctx, OGL_ERR_INVALID, "send index out of range");
    OGL_TRY(dev_alloc(ctx, &ctx->d_send_idxs, ctx->n_send));
    OGL_TRY(dev_alloc(ctx, &ctx->d_send_buf, ctx->n_send));
    OGL_TRY(dev_alloc(ctx, &ctx->d_recv_buf, ctx->n_send));
    OGL_TRY(upload(ctx, ctx->d_send_idxs, send_idxs, sizeof(label) * ctx->n_send));
    // global size (Partition.H:118-121): sum of the local sizes over all ranks
    ctx->global_n = ctx->n;
    if (ctx->n_ranks > 1 && !ctx->comm) {
        // host-driven bootstrap: ogl_partition_export / ogl_partition_connect finish the job
        p2p_teardown(ctx);
        ctx->global_n = -1;
    } else if (ctx->n_ranks > 1) {
        long long *d_cnt = nullptr;
        OGL_TRY(dev_alloc(ctx, &d_cnt, 1));
        long long h = ctx->n;
        cudaMemcpyAsync(d_cnt, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
        ncclResult_t r = ncclAllReduce(d_cnt, d_cnt, 1, ncclInt64, ncclSum, ctx->comm, ctx->stream);
        cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_cnt);
        if (r != ncclSuccess)
            return fail(ctx, OGL_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
        if (e != cudaSuccess)
            return fail(ctx, OGL_ERR_CUDA, std::string("partition: ") + cudaGetErrorString(e));
        ctx->global_n = h;
        if (ctx->comm_mode != 1) OGL_TRY(p2p_setup(ctx));
    }
    ctx->have_partition = true;
    if (ctx->graph_exec) {
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
    }
    return OGL_OK;
}

int halo_begin(Context *ctx, const double *x, bool guard_done)
{
    if (ctx->n_send == 0) return OGL_OK;
    if (ctx->n_ranks == 1)
        return fail(ctx, OGL_ERR_INVALID, "halo exchange requested on a single rank");
    if (use_p2p(ctx)) {
        k_pack_p2p<<<(ctx->n_send + 255) / 256, 256, 0, ctx->stream>>>(
            ctx->n_send, ctx->d_send_idxs, x, ctx->d_commdev, ctx->d_state, guard_done ? 1 : 0);
        ctx->launches++;
        OGL_CUDA(ctx, cudaGetLastError());
        return OGL_OK;
    }
    k_pack<<<(ctx->n_send + 255) / 256, 256, 0, ctx->stream>>>(ctx->n_send, ctx->d_send_idxs, x,
                                                              ctx->d_send_buf);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    OGL_CUDA(ctx, cudaEventRecord(ctx->ev_pack, ctx->stream));
    OGL_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_pack, 0));
    OGL_NCCL(ctx, ncclGroupStart());
    for (label t = 0; t < ctx->n_targets; ++t) {
        const label off = ctx->send_offs[t], cnt = ctx->target_sizes[t];
        if (cnt == 0) continue;
        // the recv buffer is blocked by ascending neighbour rank with the same
        // block sizes (the shared faces), Partition.H:66-67 build_from_blocked_recv
        OGL_NCCL(ctx, ncclSend(ctx->d_send_buf + off, cnt, ncclDouble, ctx->target_ids[t],
                               ctx->comm, ctx->comm_stream));
        OGL_NCCL(ctx, ncclRecv(ctx->d_recv_buf + off, cnt, ncclDouble, ctx->target_ids[t],
                               ctx->comm, ctx->comm_stream));
    }
    OGL_NCCL(ctx, ncclGroupEnd());
    OGL_CUDA(ctx, cudaEventRecord(ctx->ev_recv, ctx->comm_stream));
    return OGL_OK;
}

int halo_end(Context *ctx)
{
    if (ctx->n_send == 0 || use_p2p(ctx)) return OGL_OK;   // P2P: the consumer kernel waits on flags
    OGL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_recv, 0));
    return OGL_OK;
}

int allreduce_red(Context *ctx, int count)
{
    if (ctx->n_ranks == 1 || use_p2p(ctx)) return OGL_OK;
    double *red = &ctx->d_state->red[0];
    OGL_NCCL(ctx, ncclAllReduce(red, red, count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    return OGL_OK;
}

// y = A x through the distributed operator: boundary values travel to the
// neighbours while the local block runs, then the non-local block is applied
// to the received values (and finishes the fused reductions).
int dist_spmv(Context *ctx, const SpmvArgs &a)
{
    if (ctx->n_ranks == 1) {
        if (ctx->n_halo != 0)
            return fail(ctx, OGL_ERR_INVALID, "non-local entries on a single rank");
        SpmvArgs s = a;
        s.inline_epi = true;
        return spmv_local(ctx, s);
    }
    if (ctx->n_halo != ctx->n_send)
        return fail(ctx, OGL_ERR_INVALID, "halo pattern and partition disagree on the halo size");
    if (a.ghost_x) {
        // the caller keeps x with its ghost entries up to date: no exchange here
        SpmvArgs s = a;
        s.inline_epi = true;
        return spmv_local(ctx, s);
    }
    if (fused_halo_ok(ctx)) {
        // one kernel: local block + halo rows + all-reduce of the fused sums + epilogue
        if (!a.halo_stored) OGL_TRY(pack_stores(ctx, a.x, a.guard_done));
        SpmvArgs s = a;
        s.inline_epi = true;
        s.fused_halo = true;
        return spmv_local(ctx, s);
    }
    OGL_TRY(halo_begin(ctx, a.x, a.guard_done));
    SpmvArgs s = a;
    s.inline_epi = false;
    s.epi = EPI_NONE;
    OGL_TRY(spmv_local(ctx, s));
    OGL_TRY(halo_end(ctx));
    const bool p2p = use_p2p(ctx);
    OGL_TRY(spmv_nonlocal(ctx, ctx->d_recv_buf, a.y, a.advanced ? a.alpha : 1.0, a.dot_with,
                          a.nred, a.guard_done, p2p ? a.epi : EPI_NONE, p2p));
    if (a.nred > 0) OGL_TRY(finish_reduction(ctx, a.nred, a.epi, a.guard_done));
    return OGL_OK;
}

}  // namespace ogl
