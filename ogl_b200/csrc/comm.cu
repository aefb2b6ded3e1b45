// Multi-GPU plumbing: one process (rank) per GPU, NCCL over NVLink/NVSwitch
// (SURVEY.md section 8e).  Replaces, on this path, Ginkgo's
// distributed::Matrix::communicate + MPI_Allreduce calls:
//   halo exchange   MPI_Ineighbor_alltoallv of gathered boundary values
//                   -> pack kernel + grouped ncclSend/ncclRecv on a side stream,
//                      overlapped with the interior (local-block) SpMV
//   reductions      MPI_Allreduce(SUM) of 1 scalar, 2-3 times per iteration
//                   -> one ncclAllReduce of the packed scalars, in place in the
//                      device-resident SolveState
// The partition itself (who sends what to whom) is OGL's
// create_communication_pattern (HostMatrix/HostMatrix.C:251-306) fed through
// PartitionInitFunctor (DevicePersistent/Partition/Partition.H:57-70).
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

}  // namespace

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
    if (ctx->n_ranks > 1) {
        if (!ctx->comm) return fail(ctx, OGL_ERR_NCCL, "context has no NCCL communicator");
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
    }
    ctx->have_partition = true;
    return OGL_OK;
}

int halo_begin(Context *ctx, const double *x)
{
    if (ctx->n_send == 0) return OGL_OK;
    k_pack<<<(ctx->n_send + 255) / 256, 256, 0, ctx->stream>>>(ctx->n_send, ctx->d_send_idxs, x,
                                                              ctx->d_send_buf);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    if (ctx->n_ranks == 1) {
        // no peers (cannot happen with a valid partition): loop back for safety
        return fail(ctx, OGL_ERR_INVALID, "halo exchange requested on a single rank");
    }
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
    if (ctx->n_send == 0) return OGL_OK;
    OGL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_recv, 0));
    return OGL_OK;
}

int allreduce_red(Context *ctx, int count)
{
    if (ctx->n_ranks == 1) return OGL_OK;
    double *red = &ctx->d_state->red[0];
    OGL_NCCL(ctx, ncclAllReduce(red, red, count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    return OGL_OK;
}

// y = A x through the distributed operator: halo exchange on the side stream
// while the local block runs, then the non-local block on the received values.
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
    OGL_TRY(halo_begin(ctx, a.x));
    SpmvArgs s = a;
    s.inline_epi = false;
    s.epi = EPI_NONE;
    OGL_TRY(spmv_local(ctx, s));
    OGL_TRY(halo_end(ctx));
    OGL_TRY(spmv_nonlocal(ctx, ctx->d_recv_buf, a.y, a.advanced ? a.alpha : 1.0, a.dot_with,
                          a.nred, a.guard_done, EPI_NONE, false));
    if (a.nred > 0) OGL_TRY(finish_reduction(ctx, a.nred, a.epi, a.guard_done));
    return OGL_OK;
}

}  // namespace ogl
