// extern "C" entry points of libogl_b200.so (include/ogl_b200.h) and the small
// host-side helpers shared by the kernels' launchers.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <mutex>

#include "common.cuh"
#include "reduce.cuh"

namespace ogl {

static thread_local std::string g_last_error;

void set_error(Context *ctx, const std::string &msg)
{
    if (ctx) ctx->error = msg;
    g_last_error = msg;
}

int fail(Context *ctx, int code, const std::string &msg)
{
    set_error(ctx, msg);
    return code;
}

void invalidate_graph(Context *ctx)
{
    if (ctx && ctx->graph_exec) {
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
    }
}

int upload(Context *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return OGL_OK;
    if (!src) return fail(ctx, OGL_ERR_INVALID, "null host pointer in upload");
    // pinned sources are DMA'd directly; pageable ones are staged by the driver
    OGL_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return OGL_OK;
}

int download(Context *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return OGL_OK;
    if (!dst) return fail(ctx, OGL_ERR_INVALID, "null host pointer in download");
    OGL_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return OGL_OK;
}

int get_work(Context *ctx, int idx, double **out)
{
    if ((int)ctx->work.size() <= idx) ctx->work.resize(idx + 1, nullptr);
    // n rows + room for the ghost entries (CG ghost-p mode keeps p with its halo)
    const size_t need = (size_t)ctx->n + (size_t)(ctx->n_halo > ctx->n_send ? ctx->n_halo : ctx->n_send);
    if (ctx->work_len < need) {
        for (auto &w : ctx->work) {
            if (w) cudaFree(w);
            w = nullptr;
        }
        if (ctx->graph_exec) {   // a captured chunk holds the old addresses
            cudaGraphExecDestroy(ctx->graph_exec);
            ctx->graph_exec = nullptr;
        }
        ctx->work_len = need;
    }
    if (!ctx->work[idx]) OGL_TRY(dev_alloc(ctx, &ctx->work[idx], ctx->work_len));
    *out = ctx->work[idx];
    return OGL_OK;
}

// One NCCL communicator per (process, unique id), shared by every context
// created with that id -- as all fields share MPI_COMM_WORLD in the reference
// (ExecutorHandler.H:140-144).  Calls are serialised by the caller, so sharing
// it across the contexts' streams is safe.
struct SharedComm {
    ncclComm_t comm = nullptr;
    int refs = 0;
};
static std::mutex g_comm_mutex;
static std::map<std::string, SharedComm> g_comms;

static int acquire_comm(const void *id_bytes, int n_ranks, int rank, ncclComm_t *out,
                        std::string *key, std::string *err)
{
    std::lock_guard<std::mutex> lock(g_comm_mutex);
    key->assign(static_cast<const char *>(id_bytes), OGL_NCCL_ID_BYTES);
    auto it = g_comms.find(*key);
    if (it == g_comms.end()) {
        ncclUniqueId id;
        std::memcpy(&id, id_bytes, sizeof(id));
        ncclComm_t comm = nullptr;
        ncclResult_t r = ncclCommInitRank(&comm, n_ranks, id, rank);
        if (r != ncclSuccess) {
            *err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
            return OGL_ERR_NCCL;
        }
        it = g_comms.emplace(*key, SharedComm{comm, 0}).first;
    }
    it->second.refs++;
    *out = it->second.comm;
    return OGL_OK;
}

static void release_comm(const std::string &key)
{
    std::lock_guard<std::mutex> lock(g_comm_mutex);
    auto it = g_comms.find(key);
    if (it == g_comms.end()) return;
    // the communicator stays cached for the life of the process: an NCCL unique
    // id cannot be used for a second ncclCommInitRank, and later contexts (new
    // fields, new time steps) arrive with the same id
    if (it->second.refs > 0) --it->second.refs;
}

static void destroy(Context *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    comm_teardown(c);
    if (c->comm) release_comm(c->comm_key);
    void *ptrs[] = {c->d_rows,       c->d_cols,       c->d_map,        c->d_row_ptrs,
                    c->d_send_idxs,  c->d_send_buf,   c->d_recv_buf,   c->d_nl_rows,
                    c->d_nl_cols,    c->d_nl_map,     c->d_nl_row_ids, c->d_nl_row_ptrs,
                    c->d_staging,    c->d_vals,       c->d_nl_vals,    c->d_nl_staging,
                    c->d_b,          c->d_x,          c->d_krylov,     c->d_hess,
                    c->d_inv_diag,   c->d_block_ptrs, c->d_row_block,  c->d_block_offs,
                    c->d_inv_blocks, c->d_partials,   c->d_ticket,     c->d_state,
                    c->d_history,    c->d_g_row_ptrs, c->d_g_cols,     c->d_g_map,
                    c->d_g_vals,     c->d_trace,      c->d_push_dst,   c->d_bar,
                    c->ell.cols,     c->ell.vals,     c->ell.code,     c->ell.ptab,
                    c->gell.cols,    c->gell.vals,    c->gell.code,    c->gell.ptab,
                    c->d_isai_w,     c->d_isai_wt,    c->d_mp_chunk_row, c->d_mp_carry_row,
                    c->d_mp_carry_val};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    tri_release(c);
    mg_release(c);
    for (double *w : c->work)
        if (w) cudaFree(w);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (cudaEvent_t e : c->profile_events)
        if (e) cudaEventDestroy(e);
    cudaEvent_t evs[] = {c->ev_pack, c->ev_recv, c->ev_t0, c->ev_t1, c->ev_poll[0], c->ev_poll[1]};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete static_cast<ogl_ctx *>(c);
}

}  // namespace ogl

using namespace ogl;

namespace {

__global__ void __launch_bounds__(256, 8) k_mb_copy(const double2 *__restrict__ in,
                                                    double2 *__restrict__ out, int64_t n2)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

__global__ void __launch_bounds__(256, 8) k_mb_read(const double2 *__restrict__ in, int64_t n2,
                                                    double *sink)
{
    double acc = 0.0;
#pragma unroll 4
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = in[i];
        acc += v.x + v.y;
    }
    if (acc == 1.2345e-300) *sink = acc;
}

__global__ void __launch_bounds__(256, 8) k_mb_read2(const double *__restrict__ a,
                                                     const int *__restrict__ b, int64_t n,
                                                     double *sink)
{
    double acc = 0.0;
#pragma unroll 4
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        acc += a[i] + (double)b[i];
    if (acc == 1.2345e-300) *sink = acc;
}

}  // namespace

#define CHECK_CTX(ctx)                                                      \
    do {                                                                    \
        if (!(ctx)) {                                                       \
            ogl::set_error(nullptr, "null context");                        \
            return OGL_ERR_INVALID;                                         \
        }                                                                   \
        cudaError_t e__ = cudaSetDevice((ctx)->device);                     \
        if (e__ != cudaSuccess)                                             \
            return ogl::fail(ctx, OGL_ERR_CUDA,                             \
                             std::string("cudaSetDevice: ") + cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

int ogl_nccl_unique_id(void *out_id)
{
    if (!out_id) return OGL_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) <= OGL_NCCL_ID_BYTES, "ncclUniqueId grew");
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) {
        ogl::set_error(nullptr, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
        return OGL_ERR_NCCL;
    }
    std::memset(out_id, 0, OGL_NCCL_ID_BYTES);
    std::memcpy(out_id, &id, sizeof(id));
    return OGL_OK;
}

int ogl_device_count(int *count)
{
    if (!count) return OGL_ERR_INVALID;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        ogl::set_error(nullptr, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        return OGL_ERR_CUDA;
    }
    return OGL_OK;
}

int ogl_ctx_create(int device_id, int rank, int n_ranks, const void *nccl_id, void *stream,
                   ogl_ctx **out)
{
    if (!out) return OGL_ERR_INVALID;
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return ogl::fail(nullptr, OGL_ERR_INVALID, "bad rank / n_ranks");
    // n_ranks > 1 without an NCCL id: host-driven bootstrap of the peer-memory windows
    // (ogl_partition_export / ogl_partition_connect), no NCCL fallback path
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ogl::fail(nullptr, OGL_ERR_CUDA,
                         std::string("no CUDA device usable (there is no CPU fallback): ") +
                             (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices"));
    // ExecutorHandler.H:57-58: device_id % num_devices
    const int dev = ((device_id % ndev) + ndev) % ndev;
    e = cudaSetDevice(dev);
    if (e != cudaSuccess)
        return ogl::fail(nullptr, OGL_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    if (prop.major < 10)
        return ogl::fail(nullptr, OGL_ERR_CUDA,
                         "device is not Blackwell-class (sm_100a cubin only, no fallback)");
    ogl_ctx *c = new ogl_ctx();
    c->device = dev;
    c->rank = rank;
    c->n_ranks = n_ranks;
    if (const char *e = std::getenv("OGL_B200_PDL")) c->use_pdl = std::atoi(e) != 0;   // A/B switch
    auto bail = [&](int code, const std::string &msg) {
        ogl::set_error(nullptr, msg);
        ogl::destroy(c);
        return code;
    };
    if (stream) {
        c->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
            return bail(OGL_ERR_CUDA, "cudaStreamCreate failed");
        c->own_stream = true;
    }
    bool ok = cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_recv, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->ev_t0) == cudaSuccess && cudaEventCreate(&c->ev_t1) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_poll[0], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_poll[1], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_state, sizeof(SolveState)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_ticket, sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_state, 3 * sizeof(SolveState)) == cudaSuccess;
    if (!ok) return bail(OGL_ERR_CUDA, std::string("context allocation failed: ") +
                                           cudaGetErrorString(cudaGetLastError()));
    cudaMemset(c->d_state, 0, sizeof(SolveState));
    cudaMemset(c->d_ticket, 0, sizeof(unsigned int));
    std::memset(c->h_state, 0, 3 * sizeof(SolveState));
    if (n_ranks > 1 && nccl_id) {
        std::string err;
        const int rc = ogl::acquire_comm(nccl_id, n_ranks, rank, &c->comm, &c->comm_key, &err);
        if (rc != OGL_OK) {
            c->comm = nullptr;
            return bail(rc, err);
        }
    }
    *out = c;
    return OGL_OK;
}

int ogl_ctx_destroy(ogl_ctx *ctx)
{
    if (!ctx) return OGL_ERR_INVALID;
    ogl::destroy(ctx);
    return OGL_OK;
}

const char *ogl_last_error(const ogl_ctx *ctx)
{
    return ctx ? ctx->error.c_str() : ogl::g_last_error.c_str();
}

int ogl_set_option(ogl_ctx *ctx, const char *key, int64_t value)
{
    CHECK_CTX(ctx);
    if (!key) return fail(ctx, OGL_ERR_INVALID, "null option key");
    const std::string k(key);
    // a cached chunk graph survives a call that does not change anything (the host layers
    // re-apply their keywords on every solver construction, i.e. once per solve)
    int64_t before = 0;
    const bool known = ogl_get_option(ctx, key, &before) == OGL_OK;
    if (known && before == value && k != "trace") return OGL_OK;
    if (k == "spmv_variant") {
        if (value < 0 || value > 8) return fail(ctx, OGL_ERR_INVALID, "spmv_variant in [0,8]");
        ctx->spmv_variant = value;
    } else if (k == "chunk_iters") {
        if (value < 1) return fail(ctx, OGL_ERR_INVALID, "chunk_iters >= 1");
        ctx->chunk_iters = value;
    } else if (k == "use_graph") {
        ctx->use_graph = value != 0;
    } else if (k == "use_pdl") {
        ctx->use_pdl = value != 0;
    } else if (k == "profile_stride") {
        ctx->profile_stride = value < 0 ? 0 : value;
    } else if (k == "comm_mode") {
        if (value < 0 || value > 2) return fail(ctx, OGL_ERR_INVALID, "comm_mode in {0,1,2}");
        ctx->comm_mode = value;   // takes effect at the next ogl_partition_create / solve
    } else if (k == "fused_halo") {
        if (value < 0 || value > 2) return fail(ctx, OGL_ERR_INVALID, "fused_halo in {0,1,2}");
        ctx->fused_halo = value;
    } else if (k == "ghost_p") {
        ctx->ghost_p = value != 0;
    } else if (k == "fuse_p") {
        ctx->fuse_p = value != 0;
    } else if (k == "ell_auto") {
        ctx->ell_auto = value != 0;
    } else if (k == "gmres_persist") {
        ctx->gmres_persist = value != 0;
    } else if (k == "tri_variant") {
        if (value < 0 || value > 1) return fail(ctx, OGL_ERR_INVALID, "tri_variant in {0,1}");
        ctx->tri_variant = value;
    } else if (k == "mg_max_levels") {
        if (value < 1 || value > 64) return fail(ctx, OGL_ERR_INVALID, "mg_max_levels in [1,64]");
        ctx->mg_max_levels = value;
    } else if (k == "mg_min_coarse_rows") {
        if (value < 1) return fail(ctx, OGL_ERR_INVALID, "mg_min_coarse_rows >= 1");
        ctx->mg_min_coarse_rows = value;
    } else if (k == "mg_coarse_iters") {
        if (value < 1 || value > 10000) return fail(ctx, OGL_ERR_INVALID, "mg_coarse_iters in [1,10000]");
        ctx->mg_coarse_iters = value;
    } else if (k == "tri_sleep_ns") {
        if (value < 0 || value > 100000) return fail(ctx, OGL_ERR_INVALID, "tri_sleep_ns in [0,100000]");
        ctx->tri_sleep_ns = value;
    } else if (k == "tri_ctas") {
        if (value < 0 || value > 32) return fail(ctx, OGL_ERR_INVALID, "tri_ctas in [0,32]");
        ctx->tri_ctas = value;
    } else if (k == "ell_chunk") {
        if (value < 1 || value > 4096) return fail(ctx, OGL_ERR_INVALID, "ell_chunk in [1,4096]");
        ctx->ell_chunk = value;
    } else if (k == "ell_tma") {
        if (value < 0 || value > 2) return fail(ctx, OGL_ERR_INVALID, "ell_tma in {0,1,2}");
        ctx->ell_tma = value;
    } else if (k == "ell_minb") {
        if (value < 3 || value > 4) return fail(ctx, OGL_ERR_INVALID, "ell_minb in {3,4}");
        ctx->ell_minb = value;
    } else if (k == "ell_minb_cgp") {
        if (value < 2 || value > 4) return fail(ctx, OGL_ERR_INVALID, "ell_minb_cgp in {2,3,4}");
        ctx->ell_minb_cgp = value;
    } else if (k == "ell_coded") {
        if (value < 0 || value > 2) return fail(ctx, OGL_ERR_INVALID, "ell_coded in {0,1,2}");
        ctx->ell_coded = value;
        ell_invalidate(ctx, /*structure=*/true);
    } else if (k == "device_loop") {
        ctx->device_loop = value != 0;
    } else if (k == "loop_iters") {
        if (value < 1 || value > 1024) return fail(ctx, OGL_ERR_INVALID, "loop_iters in [1,1024]");
        ctx->loop_iters = value;
    } else if (k == "fused_pcg") {
        if (value < 0 || value > 2) return fail(ctx, OGL_ERR_INVALID, "fused_pcg in {0,1,2}");
        ctx->fused_pcg = value;
    } else if (k == "trace") {
        ctx->trace = value != 0;
        if (ctx->trace) {
            if (!ctx->d_trace) OGL_TRY(dev_alloc(ctx, &ctx->d_trace, (size_t)ogl::kTraceCap + 1));
            OGL_CUDA(ctx, cudaMemsetAsync(ctx->d_trace, 0, sizeof(unsigned long long), ctx->stream));
        }
    } else if (k == "l2_keep_mb") {
        if (value < -1 || value > 4096) return fail(ctx, OGL_ERR_INVALID, "l2_keep_mb out of range");
        ctx->l2_keep_mb = value;
    } else if (k == "tile_blocked") {
        ctx->tile_blocked = value != 0;
    } else if (k == "tma_stages") {
        if (value < 1 || value > 4) return fail(ctx, OGL_ERR_INVALID, "tma_stages in [1,4]");
        ctx->tma_stages = value;
    } else if (k == "stream_ctas") {
        if (value < 0 || value > 65535) return fail(ctx, OGL_ERR_INVALID, "stream_ctas in [0,65535]");
        ctx->stream_ctas = value;
    } else if (k == "blas1_blocks") {
        if (value < 1 || value > 65535) return fail(ctx, OGL_ERR_INVALID, "blas1_blocks in [1,65535]");
        ctx->blas1_blocks = value;
        if (ctx->have_pattern) OGL_TRY(spmv_setup(ctx));
    } else {
        return fail(ctx, OGL_ERR_INVALID, "unknown option: " + k);
    }
    if (ctx->graph_exec) {
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
    }
    return OGL_OK;
}

int ogl_get_option(ogl_ctx *ctx, const char *key, int64_t *value)
{
    CHECK_CTX(ctx);
    if (!key || !value) return fail(ctx, OGL_ERR_INVALID, "null argument");
    const std::string k(key);
    if (k == "spmv_variant") *value = ctx->spmv_variant;
    else if (k == "chunk_iters") *value = ctx->chunk_iters;
    else if (k == "use_graph") *value = ctx->use_graph;
    else if (k == "use_pdl") *value = ctx->use_pdl;
    else if (k == "profile_stride") *value = ctx->profile_stride;
    else if (k == "blas1_blocks") *value = ctx->blas1_blocks;
    else if (k == "stream_ctas") *value = ctx->stream_ctas;
    else if (k == "tma_stages") *value = ctx->tma_stages;
    else if (k == "tile_blocked") *value = ctx->tile_blocked;
    else if (k == "l2_keep_mb") *value = ctx->l2_keep_mb;
    else if (k == "ghost_p") *value = ctx->ghost_p;
    else if (k == "fused_pcg") *value = ctx->fused_pcg;
    else if (k == "device_loop") *value = ctx->device_loop;
    else if (k == "ell_auto") *value = ctx->ell_auto;
    else if (k == "ell_coded") *value = ctx->ell_coded;
    else if (k == "ell_chunk") *value = ctx->ell_chunk;
    else if (k == "gmres_persist") *value = ctx->gmres_persist;
    else if (k == "tri_variant") *value = ctx->tri_variant;
    else if (k == "tri_sleep_ns") *value = ctx->tri_sleep_ns;
    else if (k == "mg_max_levels") *value = ctx->mg_max_levels;
    else if (k == "mg_min_coarse_rows") *value = ctx->mg_min_coarse_rows;
    else if (k == "mg_coarse_iters") *value = ctx->mg_coarse_iters;
    else if (k == "tri_ctas") *value = ctx->tri_ctas;
    else if (k == "tri_levels_lower") *value = ctx->tri.structure_ready ? (int64_t)ctx->tri.lvl_l.size() - 1 : 0;
    else if (k == "tri_levels_upper") *value = ctx->tri.structure_ready ? (int64_t)ctx->tri.lvl_u.size() - 1 : 0;
    else if (k == "ell_tma") *value = ctx->ell_tma;
    else if (k == "ell_minb") *value = ctx->ell_minb;
    else if (k == "ell_minb_cgp") *value = ctx->ell_minb_cgp;
    else if (k == "ell_coded_active") *value = (ctx->ell.coded ? 1 : 0) | (ctx->gell.coded ? 2 : 0);
    else if (k == "ell_patterns") *value = ctx->ell.n_patterns;
    else if (k == "ell_escape_rows") *value = ctx->ell.n_escape;
    else if (k == "gell_patterns") *value = ctx->gell.n_patterns;
    else if (k == "gell_escape_rows") *value = ctx->gell.n_escape;
    else if (k == "fuse_p") *value = ctx->fuse_p;
    else if (k == "spmv_variant_in_use") *value = spmv_variant_in_use(ctx);
    else if (k == "loop_iters") *value = ctx->loop_iters;
    else if (k == "device_loop_active") *value = ctx->graph_exec && ctx->graph_is_loop ? 1 : 0;
    else if (k == "fused_pcg_active") *value = pcg_fused_ok(ctx) ? 1 : 0;
    else if (k == "l2_keep_level") *value = l2_keep_level(ctx);
    else if (k == "comm_mode") *value = ctx->comm_mode;
    else if (k == "fused_halo") *value = ctx->fused_halo;
    else if (k == "p2p_active") *value = use_p2p(ctx) ? 1 : 0;
    else if (k == "max_row_len") *value = ctx->max_row_len;
    else if (k.rfind("row_len_hist_", 0) == 0 && k.size() == 14 && k[13] >= '0' && k[13] <= '7')
        *value = (int64_t)ctx->row_len_hist[k[13] - '0'];
    else if (k == "max_block_nnz") *value = ctx->max_block_nnz;
    else if (k == "launches") *value = ctx->launches;
    else if (k == "precond_setups") *value = ctx->precond_setups;
    else return fail(ctx, OGL_ERR_INVALID, "unknown option: " + k);
    return OGL_OK;
}

int ogl_pattern_from_ldu(ogl_ctx *ctx, int32_t n_rows, int32_t n_faces, int symmetric,
                         const int32_t *lower_addr, const int32_t *upper_addr,
                         int32_t n_local_iface, const int32_t *iface_rows,
                         const int32_t *iface_cols)
{
    CHECK_CTX(ctx);
    return pattern_from_ldu(ctx, n_rows, n_faces, symmetric != 0, lower_addr, upper_addr,
                            n_local_iface, iface_rows, iface_cols);
}

int ogl_pattern_nnz(ogl_ctx *ctx, int64_t *local_nnz, int64_t *nonlocal_nnz)
{
    CHECK_CTX(ctx);
    if (!ctx->have_pattern) return fail(ctx, OGL_ERR_INVALID, "no pattern");
    if (local_nnz) *local_nnz = ctx->nnz;
    if (nonlocal_nnz) *nonlocal_nnz = ctx->n_halo;
    return OGL_OK;
}

int ogl_pattern_download(ogl_ctx *ctx, int32_t *rows, int32_t *cols, int32_t *ldu_mapping,
                         int32_t *row_ptrs)
{
    CHECK_CTX(ctx);
    if (!ctx->have_pattern) return fail(ctx, OGL_ERR_INVALID, "no pattern");
    if (rows) OGL_TRY(download(ctx, rows, ctx->d_rows, sizeof(label) * ctx->nnz));
    if (cols) OGL_TRY(download(ctx, cols, ctx->d_cols, sizeof(label) * ctx->nnz));
    if (ldu_mapping) OGL_TRY(download(ctx, ldu_mapping, ctx->d_map, sizeof(label) * ctx->nnz));
    if (row_ptrs) OGL_TRY(download(ctx, row_ptrs, ctx->d_row_ptrs, sizeof(label) * (ctx->n + 1)));
    return OGL_OK;
}

int ogl_partition_create(ogl_ctx *ctx, int32_t n_local, int32_t n_targets,
                         const int32_t *target_ids, const int32_t *target_sizes,
                         const int32_t *send_idxs)
{
    CHECK_CTX(ctx);
    return partition_create(ctx, n_local, n_targets, target_ids, target_sizes, send_idxs);
}

int ogl_partition_export(ogl_ctx *ctx, void *blob, int64_t capacity, int64_t *size)
{
    CHECK_CTX(ctx);
    return partition_export(ctx, blob, capacity, size);
}

int ogl_partition_connect(ogl_ctx *ctx, const void *blobs, int64_t n_blobs)
{
    CHECK_CTX(ctx);
    return partition_connect(ctx, blobs, n_blobs);
}

int ogl_partition_sizes(ogl_ctx *ctx, int64_t *local_size, int64_t *global_size)
{
    CHECK_CTX(ctx);
    if (!ctx->have_pattern) return fail(ctx, OGL_ERR_INVALID, "no pattern");
    if (local_size) *local_size = ctx->n;
    if (global_size) *global_size = ctx->global_n;
    return OGL_OK;
}

int ogl_nonlocal_pattern(ogl_ctx *ctx, int32_t n_halo, const int32_t *face_cells)
{
    CHECK_CTX(ctx);
    return nonlocal_pattern(ctx, n_halo, face_cells);
}

int ogl_nonlocal_pattern_download(ogl_ctx *ctx, int32_t *rows, int32_t *cols, int32_t *mapping)
{
    CHECK_CTX(ctx);
    if (!ctx->have_nonlocal) return fail(ctx, OGL_ERR_INVALID, "no non-local pattern");
    if (rows) OGL_TRY(download(ctx, rows, ctx->d_nl_rows, sizeof(label) * ctx->n_halo));
    if (cols) OGL_TRY(download(ctx, cols, ctx->d_nl_cols, sizeof(label) * ctx->n_halo));
    if (mapping) OGL_TRY(download(ctx, mapping, ctx->d_nl_map, sizeof(label) * ctx->n_halo));
    return OGL_OK;
}

int ogl_values_update(ogl_ctx *ctx, const double *diag, const double *upper, const double *lower,
                      const double *local_iface_bou, const double *nonlocal_bou, double scaling)
{
    CHECK_CTX(ctx);
    return values_update(ctx, diag, upper, lower, local_iface_bou, nonlocal_bou, scaling);
}

int ogl_values_download(ogl_ctx *ctx, double *local_vals, double *nonlocal_vals)
{
    CHECK_CTX(ctx);
    if (!ctx->have_values) return fail(ctx, OGL_ERR_INVALID, "no values");
    if (local_vals) OGL_TRY(download(ctx, local_vals, ctx->d_vals, sizeof(double) * ctx->nnz));
    if (nonlocal_vals && ctx->n_halo)
        OGL_TRY(download(ctx, nonlocal_vals, ctx->d_nl_vals, sizeof(double) * ctx->n_halo));
    return OGL_OK;
}

int ogl_vector_upload(ogl_ctx *ctx, int which, const double *host, double scale)
{
    CHECK_CTX(ctx);
    if (!ctx->have_pattern) return fail(ctx, OGL_ERR_INVALID, "vector upload before the pattern");
    if (which != OGL_VEC_B && which != OGL_VEC_X) return fail(ctx, OGL_ERR_INVALID, "bad vector id");
    double *dst = which == OGL_VEC_B ? ctx->d_b : ctx->d_x;
    OGL_TRY(upload(ctx, dst, host, sizeof(double) * ctx->n));
    // lduLduBase.H:242-252: b is scaled when `scaling` != 1
    if (scale != 1.0 && ctx->n > 0) OGL_TRY(vec_scale(ctx, dst, scale));
    (which == OGL_VEC_B ? ctx->have_b : ctx->have_x) = true;
    return OGL_OK;
}

int ogl_vector_download(ogl_ctx *ctx, int which, double *host)
{
    CHECK_CTX(ctx);
    if (which != OGL_VEC_B && which != OGL_VEC_X) return fail(ctx, OGL_ERR_INVALID, "bad vector id");
    if (!(which == OGL_VEC_B ? ctx->have_b : ctx->have_x))
        return fail(ctx, OGL_ERR_INVALID, "vector was never set");
    return download(ctx, host, which == OGL_VEC_B ? ctx->d_b : ctx->d_x, sizeof(double) * ctx->n);
}

int ogl_vector_fill(ogl_ctx *ctx, int which, double value)
{
    CHECK_CTX(ctx);
    if (!ctx->have_pattern) return fail(ctx, OGL_ERR_INVALID, "vector fill before the pattern");
    if (which != OGL_VEC_B && which != OGL_VEC_X) return fail(ctx, OGL_ERR_INVALID, "bad vector id");
    if (ctx->n > 0) OGL_TRY(vec_fill(ctx, which == OGL_VEC_B ? ctx->d_b : ctx->d_x, value));
    (which == OGL_VEC_B ? ctx->have_b : ctx->have_x) = true;
    return OGL_OK;
}

int ogl_precond_setup(ogl_ctx *ctx, int kind, int32_t max_block_size, int skip_sorting)
{
    CHECK_CTX(ctx);
    (void)skip_sorting;   // the assembled pattern is always row-sorted (skipSorting true)
    return precond_setup(ctx, kind, max_block_size);
}

int ogl_precond_download(ogl_ctx *ctx, int32_t *n_blocks, int32_t *block_ptrs, double *inv_blocks)
{
    CHECK_CTX(ctx);
    if (!ctx->have_precond || ctx->precond_kind != OGL_PRECOND_BJ)
        return fail(ctx, OGL_ERR_INVALID, "no BJ preconditioner");
    if (ctx->max_block_size == 1) {
        if (n_blocks) *n_blocks = ctx->n;
        if (inv_blocks) OGL_TRY(download(ctx, inv_blocks, ctx->d_inv_diag, sizeof(double) * ctx->n));
        if (block_ptrs) {
            std::vector<label> bp(ctx->n + 1);
            for (label i = 0; i <= ctx->n; ++i) bp[i] = i;
            std::memcpy(block_ptrs, bp.data(), sizeof(label) * (ctx->n + 1));
        }
        return OGL_OK;
    }
    if (n_blocks) *n_blocks = ctx->n_blocks;
    if (block_ptrs)
        OGL_TRY(download(ctx, block_ptrs, ctx->d_block_ptrs, sizeof(label) * (ctx->n_blocks + 1)));
    if (inv_blocks)
        OGL_TRY(download(ctx, inv_blocks, ctx->d_inv_blocks, sizeof(double) * ctx->inv_blocks_len));
    return OGL_OK;
}

int ogl_precond_factors_download(ogl_ctx *ctx, double *factors)
{
    CHECK_CTX(ctx);
    if (!factors) return fail(ctx, OGL_ERR_INVALID, "null argument");
    if (!ctx->have_precond || !is_tri_precond(ctx->precond_kind) || !ctx->tri.vals)
        return fail(ctx, OGL_ERR_INVALID, "no ILU / IC / IRILU preconditioner");
    return download(ctx, factors, ctx->tri.vals, sizeof(double) * ctx->nnz);
}

int ogl_mg_levels(ogl_ctx *ctx, int32_t *n_levels)
{
    CHECK_CTX(ctx);
    if (!n_levels) return fail(ctx, OGL_ERR_INVALID, "null argument");
    if (!ctx->have_precond || ctx->precond_kind != OGL_PRECOND_MULTIGRID || !ctx->mg.ready)
        return fail(ctx, OGL_ERR_INVALID, "no Multigrid preconditioner");
    *n_levels = (int32_t)ctx->mg.levels.size();
    return OGL_OK;
}

int ogl_mg_level_info(ogl_ctx *ctx, int32_t level, int32_t *n, int32_t *nnz, int32_t *n_coarse)
{
    CHECK_CTX(ctx);
    if (!n || !nnz || !n_coarse) return fail(ctx, OGL_ERR_INVALID, "null argument");
    return mg_level_info(ctx, level, n, nnz, n_coarse);
}

int ogl_mg_level_download(ogl_ctx *ctx, int32_t level, int32_t *row_ptrs, int32_t *cols, double *vals,
                          int32_t *agg)
{
    CHECK_CTX(ctx);
    return mg_level_download(ctx, level, row_ptrs, cols, vals, agg);
}

int ogl_precond_apply(ogl_ctx *ctx, const double *r_host, double *z_host)
{
    CHECK_CTX(ctx);
    if (!r_host || !z_host) return fail(ctx, OGL_ERR_INVALID, "null argument");
    if (!ctx->have_pattern || !ctx->have_precond)
        return fail(ctx, OGL_ERR_INVALID, "ogl_precond_apply before ogl_precond_setup");
    if (is_tri_precond(ctx->precond_kind)) OGL_TRY(tri_ensure_structure(ctx));
    if (ctx->precond_kind == OGL_PRECOND_MULTIGRID) OGL_TRY(mg_ensure(ctx));
    double *r, *z;
    OGL_TRY(get_work(ctx, 0, &r));
    OGL_TRY(get_work(ctx, 1, &z));
    OGL_TRY(upload(ctx, r, r_host, sizeof(double) * ctx->n));
    OGL_CUDA(ctx, cudaMemsetAsync(&ctx->d_state->comm_error, 0, sizeof(int), ctx->stream));
    OGL_TRY(precond_apply(ctx, r, z, nullptr, 0, false, 0, false, 0));
    int err = 0;
    OGL_TRY(download(ctx, &err, &ctx->d_state->comm_error, sizeof(int)));
    if (err) return fail(ctx, OGL_ERR_CUDA, "ILU/IC triangular sweep timed out waiting for a row it depends on");
    return download(ctx, z_host, z, sizeof(double) * ctx->n);
}

int ogl_solve(ogl_ctx *ctx, const ogl_solve_params *params, ogl_solve_result *result)
{
    CHECK_CTX(ctx);
    return solve(ctx, params, result);
}

int ogl_residual_history(ogl_ctx *ctx, double *out, int32_t capacity, int32_t *written)
{
    CHECK_CTX(ctx);
    if (!out || capacity < 0) return fail(ctx, OGL_ERR_INVALID, "bad history buffer");
    int n = ctx->h_state[0].n_history;
    if (n > capacity) n = capacity;
    if (n > 0 && ctx->d_history) OGL_TRY(download(ctx, out, ctx->d_history, sizeof(double) * n));
    if (written) *written = n;
    return OGL_OK;
}

int ogl_spmv(ogl_ctx *ctx, const double *x_host, double *y_host)
{
    CHECK_CTX(ctx);
    if (!ctx->have_values) return fail(ctx, OGL_ERR_INVALID, "ogl_spmv before the matrix is assembled");
    double *x, *y;
    OGL_TRY(get_work(ctx, 0, &x));
    OGL_TRY(get_work(ctx, 1, &y));
    OGL_TRY(upload(ctx, x, x_host, sizeof(double) * ctx->n));
    SpmvArgs s;
    s.x = x;
    s.y = y;
    OGL_TRY(dist_spmv(ctx, s));
    return download(ctx, y_host, y, sizeof(double) * ctx->n);
}

int ogl_spmv_bench(ogl_ctx *ctx, int32_t reps, int fused_dot, float *ms)
{
    CHECK_CTX(ctx);
    if (!ctx->have_values || reps < 1 || !ms)
        return fail(ctx, OGL_ERR_INVALID, "ogl_spmv_bench: bad state or arguments");
    double *x, *y;
    OGL_TRY(get_work(ctx, 2, &x));
    OGL_TRY(get_work(ctx, 3, &y));
    OGL_TRY(vec_fill(ctx, x, 1.0));
    ogl_solve_params p;
    std::memset(&p, 0, sizeof(p));
    p.frequency = 1;
    p.max_iter = 1;
    OGL_TRY(init_state(ctx, &p));
    SpmvArgs s;
    s.x = x;
    s.y = y;
    if (fused_dot) {
        s.dot_with = x;
        s.nred = 1;
        s.epi = EPI_CG_BETA;
    }
    for (int i = 0; i < 3; ++i) OGL_TRY(dist_spmv(ctx, s));
    OGL_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    for (int i = 0; i < reps; ++i) OGL_TRY(dist_spmv(ctx, s));
    OGL_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    OGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    OGL_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
    return OGL_OK;
}

int ogl_pcg_bench(ogl_ctx *ctx, int32_t iters, float *ms)
{
    CHECK_CTX(ctx);
    if (iters < 1 || !ms) return fail(ctx, OGL_ERR_INVALID, "ogl_pcg_bench: bad arguments");
    // a solve that can only stop on maxIter: tolerance 0 is never undercut
    ogl_solve_params p;
    std::memset(&p, 0, sizeof(p));
    p.solver = OGL_SOLVER_CG;
    p.tolerance = 0.0;
    p.rel_tol = 0.0;
    p.max_iter = iters;
    p.frequency = 1;
    ogl_solve_result r;
    OGL_TRY(solve(ctx, &p, &r));
    *ms = (float)(r.solve_us * 1e-3);
    return OGL_OK;
}

int ogl_trace_download(ogl_ctx *ctx, uint64_t *events, int64_t cap, int64_t *n_events)
{
    CHECK_CTX(ctx);
    if (!n_events || cap < 0 || (cap > 0 && !events))
        return fail(ctx, OGL_ERR_INVALID, "ogl_trace_download: bad arguments");
    *n_events = 0;
    if (!ctx->d_trace) return OGL_OK;
    unsigned long long cursor = 0;
    OGL_TRY(download(ctx, &cursor, ctx->d_trace, sizeof(cursor)));
    int64_t n = (int64_t)(cursor < (unsigned long long)ogl::kTraceCap ? cursor : ogl::kTraceCap);
    if (n > cap) n = cap;
    if (n > 0) OGL_TRY(download(ctx, events, ctx->d_trace + 1, sizeof(uint64_t) * (size_t)n));
    *n_events = n;
    // restart the timeline
    OGL_CUDA(ctx, cudaMemsetAsync(ctx->d_trace, 0, sizeof(unsigned long long), ctx->stream));
    return OGL_OK;
}

int ogl_membench(ogl_ctx *ctx, int mode, int64_t n_doubles, int32_t reps, double *gbs)
{
    CHECK_CTX(ctx);
    if (mode < 0 || mode > 2 || n_doubles < 1024 || reps < 1 || !gbs)
        return fail(ctx, OGL_ERR_INVALID, "ogl_membench: bad arguments");
    double *a = nullptr, *b = nullptr;
    OGL_TRY(dev_alloc(ctx, &a, (size_t)n_doubles));
    OGL_TRY(dev_alloc(ctx, &b, (size_t)n_doubles));
    cudaMemsetAsync(a, 0, sizeof(double) * n_doubles, ctx->stream);
    cudaMemsetAsync(b, 0, sizeof(double) * n_doubles, ctx->stream);
    const int grid = kNumSM * 8;
    double bytes = 0;
    for (int r = -2; r < reps; ++r) {
        if (r == 0) cudaEventRecord(ctx->ev_t0, ctx->stream);
        if (mode == 0) {
            k_mb_copy<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const double2 *>(a),
                                                     reinterpret_cast<double2 *>(b), n_doubles / 2);
            bytes = 16.0 * (double)(n_doubles / 2) * 2;
        } else if (mode == 1) {
            k_mb_read<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const double2 *>(a),
                                                     n_doubles / 2, b);
            bytes = 16.0 * (double)(n_doubles / 2);
        } else {
            k_mb_read2<<<grid, 256, 0, ctx->stream>>>(a, reinterpret_cast<const int *>(b), n_doubles,
                                                      b + n_doubles / 2 + 8);
            bytes = 12.0 * (double)n_doubles;
        }
    }
    cudaEventRecord(ctx->ev_t1, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1);
    cudaFree(a);
    cudaFree(b);
    if (e != cudaSuccess) return fail(ctx, OGL_ERR_CUDA, std::string("membench: ") + cudaGetErrorString(e));
    *gbs = bytes * reps / (ms * 1e-3) / 1e9;
    return OGL_OK;
}

int ogl_commbench(ogl_ctx *ctx, int mode, int32_t reps, double *us)
{
    CHECK_CTX(ctx);
    if (mode < 0 || mode > 2 || reps < 1 || !us) return fail(ctx, OGL_ERR_INVALID, "bad arguments");
    if (!ctx->have_partition) return fail(ctx, OGL_ERR_INVALID, "no partition");
    return comm_bench(ctx, mode, reps, us);
}

int ogl_synchronize(ogl_ctx *ctx)
{
    CHECK_CTX(ctx);
    OGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    OGL_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    return OGL_OK;
}

int ogl_export_mtx(ogl_ctx *ctx, int which, const char *path)
{
    CHECK_CTX(ctx);
    if (!path) return fail(ctx, OGL_ERR_INVALID, "null path");
    if (which < 0 || which > 3) return fail(ctx, OGL_ERR_INVALID, "which in {0,1,2,3}");
    std::ofstream os(path);
    if (!os) return fail(ctx, OGL_ERR_INVALID, std::string("cannot open ") + path);
    os << std::setprecision(15);   // common.C:48
    if (which == 3) {
        // partition side-car (JSON): the communication pattern the matrices do not carry
        os << "{\"rank\": " << ctx->rank << ", \"n_ranks\": " << ctx->n_ranks << ", \"n_local\": " << ctx->n
           << ", \"target_ids\": [";
        for (size_t t = 0; t < ctx->target_ids.size(); ++t) os << (t ? ", " : "") << ctx->target_ids[t];
        os << "], \"target_sizes\": [";
        for (size_t t = 0; t < ctx->target_sizes.size(); ++t) os << (t ? ", " : "") << ctx->target_sizes[t];
        os << "]}\n";
        return OGL_OK;
    }
    if (which == 2) {
        if (!ctx->have_b) return fail(ctx, OGL_ERR_INVALID, "no rhs");
        std::vector<double> b(ctx->n);
        OGL_TRY(download(ctx, b.data(), ctx->d_b, sizeof(double) * ctx->n));
        // gko::write of a Dense: array layout
        os << "%%MatrixMarket matrix array real general\n" << ctx->n << " 1\n";
        for (double v : b) os << v << "\n";
        return OGL_OK;
    }
    if (!ctx->have_values) return fail(ctx, OGL_ERR_INVALID, "no matrix values");
    const int64_t nnz = which == 0 ? ctx->nnz : ctx->n_halo;
    std::vector<label> rows(nnz), cols(nnz);
    std::vector<double> vals(nnz);
    if (nnz > 0) {
        OGL_TRY(download(ctx, rows.data(), which == 0 ? ctx->d_rows : ctx->d_nl_rows, sizeof(label) * nnz));
        OGL_TRY(download(ctx, cols.data(), which == 0 ? ctx->d_cols : ctx->d_nl_cols, sizeof(label) * nnz));
        OGL_TRY(download(ctx, vals.data(), which == 0 ? ctx->d_vals : ctx->d_nl_vals, sizeof(double) * nnz));
    }
    // gko::write(stream, mtx): coordinate layout, 1-based, row-major order
    os << "%%MatrixMarket matrix coordinate real general\n"
       << ctx->n << " " << (which == 0 ? (int64_t)ctx->n : (int64_t)ctx->n_halo) << " " << nnz << "\n";
    for (int64_t k = 0; k < nnz; ++k)
        os << rows[k] + 1 << " " << cols[k] + 1 << " " << vals[k] << "\n";
    return OGL_OK;
}

}  // extern "C"
