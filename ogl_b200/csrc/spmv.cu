// FP64 CSR SpMV for sm_100a, optionally fused with the reductions the Krylov
// step needs next (a20, kernel table row "Coo/Csr::apply").
//
// The path is HBM-bound: 12 B per stored entry (8 B value + 4 B column) plus
// 4 B row pointer and two 8 B vector elements per row; x is re-read ~7x per row
// on hex meshes but from L1/L2.  Tensor cores are irrelevant (0.17 flop/B).
//
// Variants (chosen in spmv_setup from the row-length statistics, or forced
// with ogl_set_option("spmv_variant")):
//   1 "stream"  kRowsPerBlock rows per CTA.  The CTA's contiguous slice of
//               values/columns is streamed with fully coalesced loads
//               (evict-first), each entry multiplied with its gathered x
//               (read-only path, L1/L2 hits) and parked in shared memory; then
//               one thread per row adds its products left to right.  Sum
//               order == storage order, products rounded before the add
//               (__dmul_rn/__dadd_rn), i.e. bit-identical to the sequential
//               reference-executor kernel.  For short, regular rows (FV meshes).
//   2 "scalar"  one thread per row, same arithmetic order; no shared memory.
//   3 "vector"  one warp per row (long / irregular rows); lane-strided partial
//               sums + shuffle reduction (not bit-identical, deterministic).
//   4 "tma"     persistent CTAs, slices staged by cp.async.bulk (TMA) into a
//               multi-stage shared-memory ring guarded by mbarriers.
//   5 "warp"    warp x 32-row tiles, no CTA barrier.
//   6 "pipe"    (default for short rows) variant 1 software-pipelined: the next
//               tile's (column, value) loads are in flight while this tile's
//               products are gathered, parked and added.
//
// Several GPUs: the same kernels run over the ghosted CSR (assembly.cu:
// build_ghosted; every row = local entries, then its non-local entries with
// column n + halo slot), either on x = [p | ghost p] kept current by the CG
// loop itself (ghost_x, no handshake) or, HALO, reading the ghost operands from
// the receive window after a flag handshake with the neighbours.
//
// Fused reductions (NRED): red[0] = <dot_with, y>, red[1] = <y, y>; reduced
// deterministically (reduce.cuh) and, on one rank, followed in the same launch
// by the scalar epilogue (e.g. CG: beta = <p,q>, alpha = rho / beta).
#include <cstring>

#include "common.cuh"
#include "reduce.cuh"
#include "spmv.cuh"

namespace ogl {

constexpr int kRowsPerBlock = 256;
constexpr int kStreamThreads = 256;
constexpr int kStreamSmemMax = 96 * 1024;
constexpr int kBatch = 8;            // entries per thread requested at once (7-pt rows: 7)
constexpr int kBatchStream = 7;      // stream kernel: 48-register budget (5 CTAs per SM)
constexpr int kStreamCtasPerSM = 5;  // <= 51 registers per thread
constexpr int kMaxTilesPerCta = 1024; // row-block extents cached in shared memory

namespace {

// x operand of one entry.  HALO (ghosted CSR, multi-GPU): columns >= n address
// the receive window the neighbours store into over NVLink -- read with a
// system-scope load, never through L1.
template <bool HALO>
__device__ __forceinline__ double gather_x(const double *__restrict__ x, label n, const double *recv,
                                           label c)
{
    if (HALO && c >= n) {
        double h;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(h) : "l"(recv + (c - n)) : "memory");
        return h;
    }
    return __ldg(&x[c]);
}

// tile range of a persistent CTA (or warp): strided over the grid, or one
// contiguous chunk per worker so that consecutive tiles reuse x lines in L1
__device__ __forceinline__ void tile_range(int blocked, label worker, label n_workers,
                                           label n_tiles, label &first, label &last, label &step)
{
    if (blocked) {
        const label per = (n_tiles + n_workers - 1) / n_workers;
        first = worker * per;
        last = min(first + per, n_tiles);
        step = 1;
    } else {
        first = worker;
        last = n_tiles;
        step = n_workers;
    }
}

// HALO (multi-GPU, peer-memory path): the kernel is the whole distributed
// operator.  The boundary values were stored into the neighbours' windows by
// the PREVIOUS kernel on the stream (the fused p-update or k_pack_stores); block
// 0 publishes them (data flags) on entry, tiles that own halo rows wait for the
// neighbours' flags -- which arrive while interior tiles are being processed --
// and add the non-local entries after the local row sums (Ginkgo's order:
// local apply, then `y += A_nl * recv`).  The last CTA acknowledges the
// exchange; the fused sums are all-reduced in grid_reduce's last block.
template <bool ADV, int NRED, bool HALO>
__global__ void __launch_bounds__(kStreamThreads, kStreamCtasPerSM)
k_spmv_stream(const SpmvK a)
{
    if (a.guard_done && a.state->done) return;
    extern __shared__ double prod[];
    const int tid = threadIdx.x;
    CommDev *c = HALO ? a.ea.comm : nullptr;
    unsigned long long seq = 0;
    const double *recv = nullptr;
    if (HALO) {
        seq = c->halo_seq + 1;
        if (blockIdx.x == 0 && tid < c->n_targets) st_flag(c->peer_data_flag[tid], seq);
        recv = c->my_recv + (size_t)(seq & 1ull) * c->my_recv_stride;
        // The neighbours stored their boundary values during THEIR previous kernel
        // and publish them on entry of this one, so this wait is the rank skew plus
        // one NVLink flag latency; afterwards halo operands can be prefetched with
        // the rest of a tile's loads.
        if (tid < c->n_targets && !wait_flag(&c->my_data_flag[tid], seq)) a.state->comm_error = 1;
        __syncthreads();
    }
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    // persistent CTAs: a fixed grid (8 per SM) walks the row blocks, so a fused
    // reduction leaves gridDim.x partials whatever the matrix size
    label t_first, t_last, t_step;
    tile_range(a.blocked, blockIdx.x, gridDim.x, a.n_row_blocks, t_first, t_last, t_step);
    for (label rb = t_first; rb < t_last; rb += t_step) {
        const label r0 = rb * kRowsPerBlock;
        const label nr = min((label)kRowsPerBlock, a.n - r0);
        const label s = __ldg(&a.row_ptrs[r0]);
        const label e = __ldg(&a.row_ptrs[r0 + nr]);
        // row extents of "my" row: issued early, consumed after the barrier
        label rs = 0, re = 0;
        if (tid < nr) {
            rs = __ldg(&a.row_ptrs[r0 + tid]);
            re = __ldg(&a.row_ptrs[r0 + tid + 1]);
        }
        // ---- stream the slice: coalesced value/column loads, gathered x.
        // All of a thread's entries of the slice are requested in ONE batch
        // (kBatchStream independent loads of columns, of values, then of x), so a row
        // block costs one HBM round trip plus one L2 round trip instead of one
        // pair per entry.
        const label len = e - s;
        for (label base = 0; base < len; base += kBatchStream * kStreamThreads) {
            label c[kBatchStream];
            double v[kBatchStream], xv[kBatchStream];
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) {
                const label q = base + tid + u * kStreamThreads;
                c[u] = q < len ? __ldcs(&a.cols[s + q]) : -1;
            }
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) {
                const label q = base + tid + u * kStreamThreads;
                v[u] = q < len ? __ldcs(&a.vals[s + q]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) xv[u] = c[u] >= 0 ? gather_x<HALO>(a.x, a.n, recv, c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) {
                const label q = base + tid + u * kStreamThreads;
                if (q < len) prod[q] = prod_of(v[u], xv[u], a.alpha, ADV);
            }
        }
        __syncthreads();
        // ---- one thread per row: left-to-right sum of its products
        if (tid < nr) {
            const label row = r0 + tid;
            double sum = ADV ? __dmul_rn(a.beta, a.y_in[row]) : 0.0;
            for (label q = rs - s; q < re - s; ++q) sum = __dadd_rn(sum, prod[q]);
            a.y[row] = sum;
            if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(a.dot_with[row], sum));
            if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
        }
        __syncthreads();   // prod is overwritten by the next row block
    }
    if (HALO) {
        // the last CTA to get here closes the exchange: acknowledge to the
        // neighbours (their buffer of this parity may be reused) and advance seq
        __shared__ bool last_cta;
        if (tid == 0) {
            __threadfence();
            last_cta = (atomicAdd(&c->nl_ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (last_cta) {
            if (tid < c->n_targets) st_flag(c->peer_ack_flag[tid], seq);
            if (tid == 0) {
                c->halo_seq = seq;
                c->nl_ticket = 0u;
            }
        }
        __syncthreads();
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// ---------------------------------------------------------------------------
// Variant 6: software-pipelined stream.  Same tiles and arithmetic as variant 1,
// but the (column, value) loads of the NEXT tile are issued -- into the very
// registers the current tile has just finished with -- before the CTA goes
// into its barrier / row-sum phase, and the next tile's extents one tile
// earlier still.  The HBM stream of a CTA therefore never pauses while it
// gathers x and adds rows: bytes in flight per SM stay at the level that
// variant 1 only reaches during its load phase.
// ---------------------------------------------------------------------------
template <bool ADV, int NRED, bool HALO>
__global__ void __launch_bounds__(kStreamThreads, kStreamCtasPerSM)
k_spmv_pipe(const SpmvK a)
{
    extern __shared__ double prod[];
    const int tid = threadIdx.x;
    CommDev *cm = HALO ? a.ea.comm : nullptr;
    unsigned long long seq = 0;
    const double *recv = nullptr;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    label rb, t_last, t_step;
    tile_range(a.blocked, blockIdx.x, gridDim.x, a.n_row_blocks, rb, t_last, t_step);
    label s = 0, e = 0;
    if (rb < t_last) {
        s = __ldg(&a.row_ptrs[rb * kRowsPerBlock]);
        e = __ldg(&a.row_ptrs[min((rb + 1) * kRowsPerBlock, a.n)]);
    }
    label c[kBatchStream];
    double v[kBatchStream];
    // prologue: first batch of the first tile
#pragma unroll
    for (int u = 0; u < kBatchStream; ++u) {
        const label q = tid + u * kStreamThreads;
        c[u] = q < e - s ? ld_mat(&a.cols[s + q], a.mat_policy) : -1;
    }
#pragma unroll
    for (int u = 0; u < kBatchStream; ++u) {
        const label q = tid + u * kStreamThreads;
        v[u] = q < e - s ? ld_mat(&a.vals[s + q], a.mat_policy) : 0.0;
    }
    // PDL: the matrix stream above does not depend on the previous kernel (the
    // p-update); everything from here on does.
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (a.guard_done && a.state->done) return;
    if (blockIdx.x == 0 && tid == 0) trace_event(a.ea, 0);
    if (HALO) {
        // HALO (multi-GPU, peer-memory path): see k_spmv_stream
        seq = cm->halo_seq + 1;
        if (blockIdx.x == 0 && tid < cm->n_targets) st_flag(cm->peer_data_flag[tid], seq);
        recv = cm->my_recv + (size_t)(seq & 1ull) * cm->my_recv_stride;
    }
    if (HALO) {
        // the first tile's loads are in flight; now make sure the neighbours'
        // boundary values (stored during their previous kernel) are published
        if (tid < cm->n_targets && !wait_flag(&cm->my_data_flag[tid], seq)) a.state->comm_error = 1;
        __syncthreads();
    }
    for (; rb < t_last; rb += t_step) {
        const label r0 = rb * kRowsPerBlock;
        const label nr = min((label)kRowsPerBlock, a.n - r0);
        const label len = e - s;
        // extents of the next tile: needed when this tile's products are parked
        const label rbn = rb + t_step;
        label s2 = 0, e2 = 0;
        if (rbn < t_last) {
            s2 = __ldg(&a.row_ptrs[rbn * kRowsPerBlock]);
            e2 = __ldg(&a.row_ptrs[min((rbn + 1) * kRowsPerBlock, a.n)]);
        }
        label rs = 0, re = 0;
        if (tid < nr) {
            rs = __ldg(&a.row_ptrs[r0 + tid]);
            re = __ldg(&a.row_ptrs[r0 + tid + 1]);
        }
        // ---- gather x for the batch already in registers, park the products
        {
            double xv[kBatchStream];
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) xv[u] = c[u] >= 0 ? gather_x<HALO>(a.x, a.n, recv, c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) {
                const label q = tid + u * kStreamThreads;
                if (q < len) prod[q] = prod_of(v[u], xv[u], a.alpha, ADV);
            }
        }
        // ---- tiles longer than one batch: the rest without prefetch
        for (label base = kBatchStream * kStreamThreads; base < len; base += kBatchStream * kStreamThreads) {
#pragma unroll
            for (int u = 0; u < kBatchStream; ++u) {
                const label q = base + tid + u * kStreamThreads;
                if (q < len)
                    prod[q] = prod_of(ld_mat(&a.vals[s + q], a.mat_policy),
                                      gather_x<HALO>(a.x, a.n, recv, ld_mat(&a.cols[s + q], a.mat_policy)), a.alpha, ADV);
            }
        }
        // ---- prefetch the next tile's first batch into the freed registers
#pragma unroll
        for (int u = 0; u < kBatchStream; ++u) {
            const label q = tid + u * kStreamThreads;
            c[u] = q < e2 - s2 ? ld_mat(&a.cols[s2 + q], a.mat_policy) : -1;
        }
#pragma unroll
        for (int u = 0; u < kBatchStream; ++u) {
            const label q = tid + u * kStreamThreads;
            v[u] = q < e2 - s2 ? ld_mat(&a.vals[s2 + q], a.mat_policy) : 0.0;
        }
        __syncthreads();
        // ---- one thread per row: left-to-right sum of its products
        if (tid < nr) {
            const label row = r0 + tid;
            double sum = ADV ? __dmul_rn(a.beta, a.y_in[row]) : 0.0;
            for (label q = rs - s; q < re - s; ++q) sum = __dadd_rn(sum, prod[q]);
            a.y[row] = sum;
            if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(a.dot_with[row], sum));
            if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
        }
        __syncthreads();   // prod is overwritten by the next row block
        s = s2;
        e = e2;
    }
    if (HALO) {
        // the last CTA closes the exchange: acknowledge to the neighbours, advance seq
        __shared__ bool last_cta;
        if (tid == 0) {
            __threadfence();
            last_cta = (atomicAdd(&cm->nl_ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (last_cta) {
            if (tid < cm->n_targets) st_flag(cm->peer_ack_flag[tid], seq);
            if (tid == 0) {
                cm->halo_seq = seq;
                cm->nl_ticket = 0u;
            }
        }
        __syncthreads();
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// ---------------------------------------------------------------------------
// Variant 4: TMA-fed persistent pipeline.
//
// One producer lane per CTA walks the CTA's row blocks ahead of the consumers
// and issues three 1-D bulk copies (cp.async.bulk, the TMA engine) per block --
// values, columns, row pointers -- into a ring of kStages shared-memory stages;
// completion is tracked by an mbarrier per stage (expect_tx / complete_tx).
// The 8 consumer warps never wait on HBM: they wait on the stage's mbarrier,
// gather x (L1/L2), overwrite the staged values with the products in place,
// meet at a named barrier, add each row's products left to right (same order
// as the stream kernel => bit-identical results), and hand the stage back
// through a second mbarrier.  The matrix stream carries an L2 evict-first
// policy so that x keeps its place in the 126 MB L2.
// ---------------------------------------------------------------------------

constexpr int kTmaThreads = kStreamThreads + 32;   // 8 consumer warps + 1 producer warp
constexpr int kTmaMaxStages = 4;

struct TmaHdr {
    label s, e, s_al, r0, nr, pad0, pad1, pad2;
};

template <bool ADV, int NRED>
__global__ void __launch_bounds__(kTmaThreads)
k_spmv_tma(const SpmvK a, const int cap, const int stages)
{
    if (a.guard_done && a.state->done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // carve: [vals | cols | row ptrs] per stage, then headers, then barriers
    const size_t vals_bytes = (size_t)cap * sizeof(double);
    const size_t cols_bytes = (size_t)cap * sizeof(label);
    const size_t rp_bytes_max = (size_t)(kRowsPerBlock + 4) * sizeof(label);
    const size_t stage_bytes = vals_bytes + cols_bytes + rp_bytes_max;
    TmaHdr *hdr = reinterpret_cast<TmaHdr *>(smem_raw + stage_bytes * stages);
    uint64_t *full = reinterpret_cast<uint64_t *>(hdr + kTmaMaxStages);
    uint64_t *empty = full + kTmaMaxStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int st = 0; st < stages; ++st) {
            tma::mbar_init(&full[st], 1);
            tma::mbar_init(&empty[st], kStreamThreads / 32);
        }
        tma::fence_barrier_init();
        tma::fence_proxy_async();
    }
    // extents of this CTA's row blocks, fetched once by all threads: the single
    // producer lane must not pay a global round trip per block
    __shared__ label ext[2 * kMaxTilesPerCta];
    label t_first, t_last, t_step;
    tile_range(a.blocked, blockIdx.x, gridDim.x, a.n_row_blocks, t_first, t_last, t_step);
    {
        int i = tid;
        for (label rb = t_first + (label)tid * t_step; rb < t_last && i < kMaxTilesPerCta;
             rb += (label)kTmaThreads * t_step, i += kTmaThreads) {
            const label r0 = rb * kRowsPerBlock;
            ext[2 * i] = __ldg(&a.row_ptrs[r0]);
            ext[2 * i + 1] = __ldg(&a.row_ptrs[min(r0 + (label)kRowsPerBlock, a.n)]);
        }
    }
    __syncthreads();
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;

    if (warp == kStreamThreads / 32) {
        // ===== producer: one lane feeds the ring =====
        if (lane == 0) {
            const uint64_t pol = tma::policy_evict_first();
            int i = 0;
            for (label rb = t_first; rb < t_last; rb += t_step, ++i) {
                const int st = i % stages;
                const uint32_t round = (uint32_t)(i / stages);
                if (round > 0) tma::mbar_wait(&empty[st], (round - 1) & 1);
                const label r0 = rb * kRowsPerBlock;
                const label nr = min((label)kRowsPerBlock, a.n - r0);
                const label s = i < kMaxTilesPerCta ? ext[2 * i] : __ldg(&a.row_ptrs[r0]);
                const label e = i < kMaxTilesPerCta ? ext[2 * i + 1] : __ldg(&a.row_ptrs[r0 + nr]);
                const label s_al = s & ~3;                  // 32 B (values) / 16 B (columns) aligned
                const label cnt = ((e + 3) & ~3) - s_al;
                const uint32_t rp_bytes = (uint32_t)(((nr + 1) * sizeof(label) + 15) & ~15u);
                unsigned char *base = smem_raw + stage_bytes * st;
                hdr[st].s = s;
                hdr[st].e = e;
                hdr[st].s_al = s_al;
                hdr[st].r0 = r0;
                hdr[st].nr = nr;
                const uint32_t bytes = (uint32_t)cnt * 12u + rp_bytes;
                tma::mbar_expect_tx(&full[st], bytes);
                if (cnt > 0) {
                    tma::bulk_load(base, a.vals + s_al, (uint32_t)cnt * 8u, &full[st], pol);
                    tma::bulk_load(base + vals_bytes, a.cols + s_al, (uint32_t)cnt * 4u, &full[st], pol);
                }
                tma::bulk_load(base + vals_bytes + cols_bytes, a.row_ptrs + r0, rp_bytes, &full[st], pol);
            }
        }
        __syncwarp();   // reconverge the producer warp before the block-wide reduction
    } else {
        // ===== consumers: 256 threads =====
        int i = 0;
        for (label rb = t_first; rb < t_last; rb += t_step, ++i) {
            const int st = i % stages;
            const uint32_t round = (uint32_t)(i / stages);
            tma::mbar_wait(&full[st], round & 1);
            unsigned char *base = smem_raw + stage_bytes * st;
            double *v = reinterpret_cast<double *>(base);
            const label *c = reinterpret_cast<const label *>(base + vals_bytes);
            const label *rp = reinterpret_cast<const label *>(base + vals_bytes + cols_bytes);
            const label s = hdr[st].s, e = hdr[st].e, s_al = hdr[st].s_al;
            const label r0 = hdr[st].r0, nr = hdr[st].nr;
            const label off = s - s_al, len = e - s;
            double dw = 0.0, yin = 0.0;
            if (tid < nr) {
                if (NRED >= 1) dw = a.dot_with[r0 + tid];
                if (ADV) yin = a.y_in[r0 + tid];
            }
            // products in place: v[k] <- v[k] * x[c[k]]; all x gathers of the
            // thread are in flight together (one L2 round trip per row block)
            for (label base = 0; base < len; base += kBatch * kStreamThreads) {
                double xv[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const label q = base + tid + u * kStreamThreads;
                    xv[u] = q < len ? __ldg(&a.x[c[off + q]]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const label q = base + tid + u * kStreamThreads;
                    if (q < len) v[off + q] = prod_of(v[off + q], xv[u], a.alpha, ADV);
                }
            }
            tma::consumer_sync();
            if (tid < nr) {
                const label row = r0 + tid;
                double sum = ADV ? __dmul_rn(a.beta, yin) : 0.0;
                const label qe = rp[tid + 1] - s_al;
                for (label q = rp[tid] - s_al; q < qe; ++q) sum = __dadd_rn(sum, v[q]);
                a.y[row] = sum;
                if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
                if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
            }
            // generic-proxy writes to the stage must be ordered before the next
            // bulk copy (async proxy) overwrites it
            tma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&empty[st]);
        }
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// ---------------------------------------------------------------------------
// Variant 5: warp-synchronous stream.  Same arithmetic as variant 1, but the
// unit of work is one WARP x 32 rows: no CTA-wide barrier, so the 40 resident
// warps of an SM drift apart and overlap each other's load / gather / add
// phases; the row pointers of the NEXT tile are prefetched while the current
// one is processed, so a tile costs one HBM round trip (columns + values, all
// requested at once) plus one L2 round trip (x).
// ---------------------------------------------------------------------------
constexpr int kWarpRows = 32;
constexpr int kWarpCtaThreads = 128;   // 4 independent warps per CTA
constexpr int kWarpCtasPerSM = 8;      // 32 warps per SM, <= 64 registers per thread

template <bool ADV, int NRED>
__global__ void __launch_bounds__(kWarpCtaThreads, kWarpCtasPerSM)
k_spmv_warp(const SpmvK a, const int warp_cap)
{
    if (a.guard_done && a.state->done) return;
    extern __shared__ double prod_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *prod = prod_all + (size_t)warp * warp_cap;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    const label n_tiles = (a.n + kWarpRows - 1) / kWarpRows;
    label tile, t_last, stride;
    tile_range(a.blocked, blockIdx.x * (kWarpCtaThreads / 32) + warp,
               gridDim.x * (kWarpCtaThreads / 32), n_tiles, tile, t_last, stride);
    // row pointers of the first tile
    label rs = 0, re = 0;
    if (tile < t_last) {
        const label row = min(tile * kWarpRows + lane, a.n - 1);
        rs = __ldg(&a.row_ptrs[row]);
        re = __ldg(&a.row_ptrs[row + 1]);
    }
    for (; tile < t_last; tile += stride) {
        const label r0 = tile * kWarpRows;
        const label nr = min((label)kWarpRows, a.n - r0);
        const label s = __shfl_sync(0xffffffffu, rs, 0);
        const label e = __shfl_sync(0xffffffffu, re, nr - 1);
        const label my_rs = rs, my_re = re;
        // prefetch the next tile's row pointers
        const label nxt = tile + stride;
        if (nxt < t_last) {
            const label row = min(nxt * kWarpRows + lane, a.n - 1);
            rs = __ldg(&a.row_ptrs[row]);
            re = __ldg(&a.row_ptrs[row + 1]);
        }
        double dw = 0.0, yin = 0.0;
        if (lane < nr) {
            if (NRED >= 1) dw = a.dot_with[r0 + lane];
            if (ADV) yin = a.y_in[r0 + lane];
        }
        const label len = e - s;
        for (label base = 0; base < len; base += kBatch * 32) {
            label c[kBatch];
            double v[kBatch], xv[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const label q = base + lane + u * 32;
                c[u] = q < len ? __ldcs(&a.cols[s + q]) : -1;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const label q = base + lane + u * 32;
                v[u] = q < len ? __ldcs(&a.vals[s + q]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) xv[u] = c[u] >= 0 ? __ldg(&a.x[c[u]]) : 0.0;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const label q = base + lane + u * 32;
                if (q < len) prod[q] = prod_of(v[u], xv[u], a.alpha, ADV);
            }
        }
        __syncwarp();
        if (lane < nr) {
            double sum = ADV ? __dmul_rn(a.beta, yin) : 0.0;
            for (label q = my_rs - s; q < my_re - s; ++q) sum = __dadd_rn(sum, prod[q]);
            a.y[r0 + lane] = sum;
            if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
            if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
        }
        __syncwarp();
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

template <bool ADV, int NRED>
__global__ void __launch_bounds__(256) k_spmv_scalar(const SpmvK a)
{
    if (a.guard_done && a.state->done) return;
    const label row = blockIdx.x * blockDim.x + threadIdx.x;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    if (row < a.n) {
        const label rs = __ldg(&a.row_ptrs[row]), re = __ldg(&a.row_ptrs[row + 1]);
        double sum = ADV ? __dmul_rn(a.beta, a.y_in[row]) : 0.0;
        for (label q = rs; q < re; ++q)
            sum = __dadd_rn(sum, prod_of(__ldcs(&a.vals[q]), __ldg(&a.x[__ldcs(&a.cols[q])]),
                                         a.alpha, ADV));
        a.y[row] = sum;
        if (NRED >= 1) red[0] = __dmul_rn(a.dot_with[row], sum);
        if (NRED >= 2) red[1] = __dmul_rn(sum, sum);
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

template <bool ADV, int NRED>
__global__ void __launch_bounds__(256) k_spmv_vector(const SpmvK a)
{
    if (a.guard_done && a.state->done) return;
    const int lane = threadIdx.x & 31;
    const label row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    if (row < a.n) {
        const label rs = __ldg(&a.row_ptrs[row]), re = __ldg(&a.row_ptrs[row + 1]);
        double sum = 0.0;
        for (label q = rs + lane; q < re; q += 32)
            sum = __dadd_rn(sum, prod_of(__ldcs(&a.vals[q]), __ldg(&a.x[__ldcs(&a.cols[q])]),
                                         a.alpha, ADV));
        sum = warp_sum(sum);
        if (lane == 0) {
            if (ADV) sum = __dadd_rn(__dmul_rn(a.beta, a.y_in[row]), sum);
            a.y[row] = sum;
            if (NRED >= 1) red[0] = __dmul_rn(a.dot_with[row], sum);
            if (NRED >= 2) red[1] = __dmul_rn(sum, sum);
        }
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// y[row] += alpha * A_nl[row, :] * recv for the rows that touch the halo
// (distributed::Matrix::apply: non_local_mtx->apply(alpha, recv, one, y)), one
// thread per such row, entries added one at a time in storage order.  With
// fused reductions the kernel adds the CHANGE of <d,y> and <y,y> caused by the
// halo terms on top of the local kernel's sums.
struct NonLocalK {
    label n_rows;
    const label *row_ids, *row_ptrs, *cols;
    const double *vals, *recv;
    double *y;
    double alpha;
    const double *dot_with;
    double *partials;
    unsigned int *ticket;
    SolveState *state;
    int epi, inline_epi, guard_done;
    EpiArgs ea;
};

template <int NRED>
__global__ void __launch_bounds__(256) k_spmv_nonlocal(const NonLocalK a)
{
    if (a.guard_done && a.state->done) return;
    const double *recv = a.recv;
    CommDev *c = a.ea.comm;
    unsigned long long seq = 0;
    if (c != nullptr && c->n_targets > 0) {
        // peer-memory path: the neighbours' pack kernels store into my window;
        // wait for their data flags of the current exchange
        seq = c->halo_seq;
        if (threadIdx.x < c->n_targets) {
            if (!wait_flag(&c->my_data_flag[threadIdx.x], seq)) a.state->comm_error = 1;
        }
        __syncthreads();
        recv = c->my_recv + (size_t)(seq & 1ull) * c->my_recv_stride;
    }
    const label u = blockIdx.x * blockDim.x + threadIdx.x;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    if (u < a.n_rows) {
        const label row = a.row_ids[u];
        const double y_old = a.y[row];
        double acc = y_old;
        for (label q = a.row_ptrs[u]; q < a.row_ptrs[u + 1]; ++q) {
            double h;
            if (c != nullptr)
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(h) : "l"(recv + a.cols[q]) : "memory");
            else
                h = recv[a.cols[q]];
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(a.alpha, a.vals[q]), h));
        }
        a.y[row] = acc;
        if (NRED >= 1) {
            const double d = a.dot_with[row];
            red[0] = __dmul_rn(d, acc) - __dmul_rn(d, y_old);
        }
        if (NRED >= 2) red[1] = __dmul_rn(acc, acc) - __dmul_rn(y_old, y_old);
    }
    if (c != nullptr && c->n_targets > 0) {
        // acknowledge: this rank is done reading the buffer of exchange `seq`
        __shared__ bool last_nl;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int tk = atomicAdd(&c->nl_ticket, 1u);
            last_nl = (tk == gridDim.x - 1);
        }
        __syncthreads();
        if (last_nl) {
            if (threadIdx.x < c->n_targets) st_flag(c->peer_ack_flag[threadIdx.x], seq);
            if (threadIdx.x == 0) c->nl_ticket = 0u;
        }
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea, /*accumulate=*/true);
}

__global__ void k_block_nnz_max(label n, const label *__restrict__ row_ptrs, int rows_per_block,
                                int *out)
{
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r0 = b * rows_per_block;
    if (r0 < n) {
        const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
        atomicMax(out, row_ptrs[r1] - row_ptrs[r0]);
    }
}

// row-length histogram in powers of two: bucket b counts the rows (hist[b]) and their entries
// (hist[8 + b]) with length in (2^(b-1), 2^b] for b = 1..6, b = 0: empty rows and length 1,
// b = 7: longer than 64
__global__ void k_row_len_hist(label n, const label *__restrict__ row_ptrs, unsigned long long *hist)
{
    __shared__ unsigned long long sh[16];
    if (threadIdx.x < 16) sh[threadIdx.x] = 0ull;
    __syncthreads();
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int len = row_ptrs[r + 1] - row_ptrs[r];
        int b = 0;
        while (b < 7 && (1 << b) < len) ++b;
        atomicAdd(&sh[b], 1ull);
        atomicAdd(&sh[8 + b], (unsigned long long)len);
    }
    __syncthreads();
    if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

template <typename K, typename A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const A &args)
{
    kernel<<<grid, block, smem, st>>>(args);
}

}  // namespace

bool use_p2p(const Context *ctx);

EpiArgs make_epi_args(Context *ctx, int ar_count, bool ar_after_epi)
{
    EpiArgs ea;
    ea.inv_n_local = ctx->n > 0 ? 1.0 / (double)ctx->n : 0.0;
    const double g = (double)(ctx->global_n > 0 ? ctx->global_n : ctx->n);
    ea.weight = g > 0 ? (double)ctx->n / g : 0.0;
    ea.history = ctx->d_history;
    const bool p2p = use_p2p(ctx);
    ea.comm = p2p ? ctx->d_commdev : nullptr;
    ea.ar_count = p2p ? ar_count : 0;
    ea.ar_after_epi = ar_after_epi ? 1 : 0;
    ea.trace = ctx->trace ? ctx->d_trace : nullptr;
    ea.trace_tag = 0;
    ea.trace_cap = kTraceCap;
    return ea;
}

int spmv_setup(Context *ctx)
{
    // largest slice any kRowsPerBlock-row CTA would have to park in shared memory
    int *d_max = nullptr;
    OGL_CUDA(ctx, cudaMalloc(&d_max, 2 * sizeof(int)));
    cudaMemsetAsync(d_max, 0, 2 * sizeof(int), ctx->stream);
    const int64_t nblk = (ctx->n + kRowsPerBlock - 1) / kRowsPerBlock;
    const int64_t nwt = (ctx->n + kWarpRows - 1) / kWarpRows;
    if (nblk > 0) {
        k_block_nnz_max<<<(int)((nblk + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->n, ctx->d_row_ptrs, kRowsPerBlock, d_max);
        k_block_nnz_max<<<(int)((nwt + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->n, ctx->d_row_ptrs, kWarpRows, d_max + 1);
    }
    unsigned long long *d_hist = nullptr;
    if (cudaMalloc(&d_hist, 16 * sizeof(unsigned long long)) == cudaSuccess) {
        cudaMemsetAsync(d_hist, 0, 16 * sizeof(unsigned long long), ctx->stream);
        if (ctx->n > 0)
            k_row_len_hist<<<kNumSM * 4, 256, 0, ctx->stream>>>(ctx->n, ctx->d_row_ptrs, d_hist);
        cudaMemcpyAsync(ctx->row_len_hist, d_hist, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                        ctx->stream);
    }
    int mx2[2] = {0, 0};
    int &mx = mx2[0];
    cudaMemcpyAsync(mx2, d_max, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_max);
    cudaFree(d_hist);
    if (e != cudaSuccess)
        return fail(ctx, OGL_ERR_CUDA, std::string("spmv_setup: ") + cudaGetErrorString(e));
    ctx->max_block_nnz = mx;
    ctx->max_warp_nnz = mx2[1];
    // reduction scratch sized for the largest grid any kernel of the library uses
    int64_t max_grid = nblk;
    const int64_t vec_grid = ((int64_t)ctx->n * 32 + 255) / 256;
    if (vec_grid > max_grid) max_grid = vec_grid;
    if (ctx->blas1_blocks > max_grid) max_grid = ctx->blas1_blocks;
    OGL_TRY(dev_alloc(ctx, &ctx->d_partials, (size_t)(max_grid + 1) * kMaxReduce));
    (void)spmv_l2_policy(ctx);   // create the policy words outside any graph capture
    return spmv_merge_setup(ctx);
}

// ---------------------------------------------------------------------------
// L2 residency of the matrix stream.  The 126 MB L2 of a B200 holds the whole
// CSR stream of a 1 M-cell system (87 MB): marking those lines evict-last keeps
// them on chip from one Krylov iteration to the next, so that only the vectors
// travel over HBM.  Larger matrices keep an address-hashed fraction
// (1/2 .. 1/16) resident and stream the rest evict-first.  The 64-bit policy
// words are produced once per context by createpolicy on the device.
// ---------------------------------------------------------------------------
__global__ void k_make_policies(unsigned long long *out)
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    out[0] = p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    out[1] = p;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.5;" : "=l"(p));
    out[2] = p;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.25;" : "=l"(p));
    out[3] = p;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.125;" : "=l"(p));
    out[4] = p;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.0625;" : "=l"(p));
    out[5] = p;
}

// 0: stream everything; 1..5: keep 1, 1/2, 1/4, 1/8, 1/16 of the lines
int l2_keep_level(const Context *ctx)
{
    // auto: leave ~1/3 of the L2 to the vectors of the iteration
    const double budget = (ctx->l2_keep_mb < 0 ? 88.0 : (double)ctx->l2_keep_mb) * 1.0e6;
    if (budget <= 0.0 || ctx->nnz <= 0) return 0;
    const double bytes = 12.0 * (double)ctx->nnz;
    double frac = 1.0;
    for (int level = 1; level <= 5; ++level, frac *= 0.5)
        if (bytes * frac <= budget) return level;
    return 0;
}

unsigned long long spmv_l2_policy(Context *ctx)
{
    if (!ctx->l2_policies_ready) {
        unsigned long long *d = nullptr;
        if (cudaMalloc(&d, sizeof(ctx->l2_policies)) != cudaSuccess) return 0;   // caught by the launch check
        k_make_policies<<<1, 1, 0, ctx->stream>>>(d);
        cudaMemcpyAsync(ctx->l2_policies, d, sizeof(ctx->l2_policies), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d);
        ctx->l2_policies_ready = true;
    }
    return ctx->l2_policies[l2_keep_level(ctx)];
}

int spmv_variant_in_use(const Context *ctx);
static int pick_variant(const Context *ctx)
{
    if (ctx->spmv_variant >= 1 && ctx->spmv_variant <= 8) return (int)ctx->spmv_variant;
    const size_t smem = (size_t)ctx->max_block_nnz * sizeof(double);
    const double mean_len = ctx->n > 0 ? (double)ctx->nnz / ctx->n : 0.0;
    // ell_auto (default off): short regular rows on one rank -> ELL.  The plain ELL SpMV is the
    // fastest kernel here (profiles/r01_ell_probe.jsonl: 124.6 us = 6.65 TB/s at 8 M rows against
    // 142.9 us CSR; 16.4 vs 20.5 us at 1 M), but its instantiation with the fused <p,q> is not
    // (167 vs 155 us; 24.6-26.6 vs 24.6 us), so the PCG iteration loses (39.9 vs 38.6 us).
    // Several ranks: the CG loop's SpMV (ghost_x) runs over an ELL copy of the ghosted matrix; the
    // flag-handshake SpMV of the other solvers stays on the pipelined CSR kernel.
    if (ctx->ell_auto && ctx->n > 262144 && ctx->max_row_len <= 16 &&
        (double)ctx->max_row_len * ctx->n <= 1.25 * (double)ctx->nnz)
        return 7;
    // row-length histogram (spmv_setup): the share of the entries that sit in long rows decides.
    // Short rows (finite-volume meshes: 5..30 entries) -> the pipelined stream kernel, whose tiles
    // park every product in shared memory; a matrix whose entries are mostly in rows longer than 32
    // -> one warp per row; in between, or when a tile would not fit in shared memory -> thread per row.
    const double nnz = ctx->nnz > 0 ? (double)ctx->nnz : 1.0;
    const double share_long = (double)(ctx->row_len_hist[8 + 6] + ctx->row_len_hist[8 + 7]) / nnz;   // rows > 32
    const double share_huge = (double)ctx->row_len_hist[8 + 7] / nnz;                                 // rows > 64
    // a few rows far longer than the rest (a cell coupled to thousands of others): every row-based kernel
    // serialises on them -> entry-balanced slices (spmv_merge.cu)
    if (ctx->max_row_len > 256 && (double)ctx->max_row_len > 16.0 * mean_len) return 8;
    if (share_long > 0.5) return 3;
    if (smem <= (size_t)kStreamSmemMax && share_huge < 0.05 && mean_len <= 48.0) return 6;   // pipelined stream
    return mean_len >= 16.0 ? 3 : 2;
}

int spmv_variant_in_use(const Context *ctx) { return pick_variant(ctx); }

int spmv_local(Context *ctx, const SpmvArgs &sa)
{
    if (!ctx->have_pattern || !ctx->have_values)
        return fail(ctx, OGL_ERR_INVALID, "SpMV without assembled matrix");
    // an empty rank of a decomposed case still launches: the kernel carries the halo
    // handshake and the all-reduce its peers are waiting in
    if (ctx->n == 0 && ctx->n_ranks == 1) return OGL_OK;
    SpmvK k;
    k.row_ptrs = ctx->d_row_ptrs;
    k.cols = ctx->d_cols;
    k.vals = ctx->d_vals;
    k.x = sa.x;
    k.y_in = sa.y_in ? sa.y_in : sa.y;
    k.y = sa.y;
    k.n = ctx->n;
    k.n_row_blocks = 0;
    k.blocked = ctx->tile_blocked ? 1 : 0;
    k.mat_policy = spmv_l2_policy(ctx);
    k.alpha = sa.alpha;
    k.beta = sa.beta;
    k.dot_with = sa.dot_with;
    k.partials = ctx->d_partials;
    k.ticket = ctx->d_ticket;
    k.state = ctx->d_state;
    k.epi = sa.epi;
    k.inline_epi = sa.inline_epi ? 1 : 0;
    k.guard_done = sa.guard_done ? 1 : 0;
    k.ea = make_epi_args(ctx, 0);
    k.ea.trace_tag = 20;
    const int nred = sa.nred;
    if (nred > 0 && !sa.dot_with) return fail(ctx, OGL_ERR_INVALID, "fused dot without vector");
    int variant = pick_variant(ctx);
    if (variant == 7 && sa.fused_halo) variant = 6;   // the flag-handshake kernel exists for CSR only
    if (variant == 8 && (sa.fused_halo || sa.ghost_x)) variant = 6;   // merge-path: local matrix only
    if (sa.vals_override) {
        // another operator over the same local pattern (ISAI): CSR kernels, local block only
        if (sa.fused_halo || sa.ghost_x) return fail(ctx, OGL_ERR_INVALID, "vals_override is local-only");
        k.vals = sa.vals_override;
        if (variant == 7 || variant == 4) variant = 6;
        if (variant == 6 && (size_t)ctx->max_block_nnz * sizeof(double) > (size_t)kStreamSmemMax) variant = 2;
        k.ea = make_epi_args(ctx, sa.ar_count), k.ea.trace_tag = 20;
        if (pick_variant(ctx) == 8) variant = 8;   // same pattern, same slices
    }
    // a rank without halo rows (n_halo == 0) runs the halo kernel on its local matrix
    const bool ghosted = (sa.fused_halo || sa.ghost_x) && ctx->have_ghosted;
    cudaStream_t st = ctx->stream;
#define DISPATCH(KERNEL, GRID, BLOCK, SMEM)                                              \
    do {                                                                                 \
        if (sa.advanced) {                                                               \
            if (nred == 0) KERNEL<true, 0><<<GRID, BLOCK, SMEM, st>>>(k);                \
            else if (nred == 1) KERNEL<true, 1><<<GRID, BLOCK, SMEM, st>>>(k);           \
            else KERNEL<true, 2><<<GRID, BLOCK, SMEM, st>>>(k);                          \
        } else {                                                                         \
            if (nred == 0) KERNEL<false, 0><<<GRID, BLOCK, SMEM, st>>>(k);               \
            else if (nred == 1) KERNEL<false, 1><<<GRID, BLOCK, SMEM, st>>>(k);          \
            else KERNEL<false, 2><<<GRID, BLOCK, SMEM, st>>>(k);                         \
        }                                                                                \
    } while (0)
    if (variant == 1) {
        const size_t smem = (size_t)(ghosted ? ctx->max_block_nnz_g : ctx->max_block_nnz) * sizeof(double);
        static bool attr_done[64] = {};   // cudaFuncSetAttribute is per device
        bool &attr_set = attr_done[ctx->device & 63];
        if (!attr_set) {
#define SET_ATTR(A, R)                                                                             \
    cudaFuncSetAttribute(k_spmv_stream<A, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                         kStreamSmemMax);                                                          \
    cudaFuncSetAttribute(k_spmv_stream<A, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                         kStreamSmemMax);
            SET_ATTR(false, 0) SET_ATTR(false, 1) SET_ATTR(false, 2)
            SET_ATTR(true, 0) SET_ATTR(true, 1) SET_ATTR(true, 2)
#undef SET_ATTR
            attr_set = true;
        }
        const int nblk = (ctx->n + kRowsPerBlock - 1) / kRowsPerBlock;
        k.n_row_blocks = nblk;
        const int64_t cap = ctx->stream_ctas > 0 ? ctx->stream_ctas : (int64_t)kNumSM * kStreamCtasPerSM;
        const int grid = nblk < 1 ? 1 : (nblk < cap ? nblk : (int)cap);
#define STREAM_LAUNCH(H)                                                                          \
    do {                                                                                          \
        if (sa.advanced) {                                                                        \
            if (nred == 0) k_spmv_stream<true, 0, H><<<grid, kStreamThreads, smem, st>>>(k);      \
            else if (nred == 1) k_spmv_stream<true, 1, H><<<grid, kStreamThreads, smem, st>>>(k); \
            else k_spmv_stream<true, 2, H><<<grid, kStreamThreads, smem, st>>>(k);                \
        } else {                                                                                  \
            if (nred == 0) k_spmv_stream<false, 0, H><<<grid, kStreamThreads, smem, st>>>(k);     \
            else if (nred == 1) k_spmv_stream<false, 1, H><<<grid, kStreamThreads, smem, st>>>(k);\
            else k_spmv_stream<false, 2, H><<<grid, kStreamThreads, smem, st>>>(k);               \
        }                                                                                         \
    } while (0)
        if (sa.fused_halo) {
            if (ghosted) {
                // the ghosted CSR: local entries + non-local ones behind them in every row
                k.row_ptrs = ctx->d_g_row_ptrs;
                k.cols = ctx->d_g_cols;
                k.vals = ctx->d_g_vals;
            }
            k.ea = make_epi_args(ctx, nred), k.ea.trace_tag = 20;
            STREAM_LAUNCH(true);
        } else {
            if (sa.ghost_x) {
                if (ghosted) {
                    k.row_ptrs = ctx->d_g_row_ptrs;
                    k.cols = ctx->d_g_cols;
                    k.vals = ctx->d_g_vals;
                }
                k.ea = make_epi_args(ctx, nred), k.ea.trace_tag = 20;
            }
            STREAM_LAUNCH(false);
        }
#undef STREAM_LAUNCH
    } else if (variant == 6) {
        const size_t smem = (size_t)(ghosted ? ctx->max_block_nnz_g : ctx->max_block_nnz) * sizeof(double);
        static bool attr6_done[64] = {};
        bool &attr6 = attr6_done[ctx->device & 63];
        if (!attr6) {
#define SET_ATTR6(A, R)                                                                          \
    cudaFuncSetAttribute(k_spmv_pipe<A, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                         kStreamSmemMax);                                                        \
    cudaFuncSetAttribute(k_spmv_pipe<A, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                         kStreamSmemMax);
            SET_ATTR6(false, 0) SET_ATTR6(false, 1) SET_ATTR6(false, 2)
            SET_ATTR6(true, 0) SET_ATTR6(true, 1) SET_ATTR6(true, 2)
#undef SET_ATTR6
            attr6 = true;
        }
        const int nblk = (ctx->n + kRowsPerBlock - 1) / kRowsPerBlock;
        k.n_row_blocks = nblk;
        const int64_t cap = ctx->stream_ctas > 0 ? ctx->stream_ctas : (int64_t)kNumSM * kStreamCtasPerSM;
        const int grid = nblk < 1 ? 1 : (nblk < cap ? nblk : (int)cap);
#define PIPE_ONE(A, R, H) \
    launch_pdl(k_spmv_pipe<A, R, H>, grid, kStreamThreads, smem, st, ctx->use_pdl != 0 && sa.guard_done, k)
#define PIPE_LAUNCH(H)                                                      \
    do {                                                                    \
        cudaError_t le;                                                     \
        if (sa.advanced) {                                                  \
            if (nred == 0) le = PIPE_ONE(true, 0, H);                       \
            else if (nred == 1) le = PIPE_ONE(true, 1, H);                  \
            else le = PIPE_ONE(true, 2, H);                                 \
        } else {                                                            \
            if (nred == 0) le = PIPE_ONE(false, 0, H);                      \
            else if (nred == 1) le = PIPE_ONE(false, 1, H);                 \
            else le = PIPE_ONE(false, 2, H);                                \
        }                                                                   \
        OGL_CUDA(ctx, le);                                                  \
    } while (0)
        if (sa.fused_halo) {
            if (ghosted) {
                // the ghosted CSR: local entries + non-local ones behind them in every row
                k.row_ptrs = ctx->d_g_row_ptrs;
                k.cols = ctx->d_g_cols;
                k.vals = ctx->d_g_vals;
            }
            k.ea = make_epi_args(ctx, nred), k.ea.trace_tag = 20;
            PIPE_LAUNCH(true);
        } else {
            if (sa.ghost_x) {
                if (ghosted) {
                    k.row_ptrs = ctx->d_g_row_ptrs;
                    k.cols = ctx->d_g_cols;
                    k.vals = ctx->d_g_vals;
                }
                k.ea = make_epi_args(ctx, nred), k.ea.trace_tag = 20;   // the fused sums are all-reduced in this launch
            }
            PIPE_LAUNCH(false);
        }
#undef PIPE_LAUNCH
#undef PIPE_ONE
    } else if (variant == 5) {
        const int warp_cap = (int)((ctx->max_warp_nnz + 1) & ~(int64_t)1);
        const size_t smem = (size_t)warp_cap * sizeof(double) * (kWarpCtaThreads / 32);
        if (smem > (size_t)kStreamSmemMax)
            return fail(ctx, OGL_ERR_UNSUPPORTED, "rows too long for the warp-tile SpMV");
        static bool attr5_done[64] = {};
        bool &attr5 = attr5_done[ctx->device & 63];
        if (!attr5) {
            cudaFuncSetAttribute(k_spmv_warp<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemMax);
            cudaFuncSetAttribute(k_spmv_warp<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemMax);
            cudaFuncSetAttribute(k_spmv_warp<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemMax);
            cudaFuncSetAttribute(k_spmv_warp<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemMax);
            cudaFuncSetAttribute(k_spmv_warp<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemMax);
            cudaFuncSetAttribute(k_spmv_warp<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemMax);
            attr5 = true;
        }
        const int64_t n_tiles = (ctx->n + kWarpRows - 1) / kWarpRows;
        const int64_t need = (n_tiles + (kWarpCtaThreads / 32) - 1) / (kWarpCtaThreads / 32);
        const int64_t cap = ctx->stream_ctas > 0 ? ctx->stream_ctas : (int64_t)kNumSM * kWarpCtasPerSM;
        const int grid = (int)(need < cap ? need : cap);
        if (sa.advanced) {
            if (nred == 0) k_spmv_warp<true, 0><<<grid, kWarpCtaThreads, smem, st>>>(k, warp_cap);
            else if (nred == 1) k_spmv_warp<true, 1><<<grid, kWarpCtaThreads, smem, st>>>(k, warp_cap);
            else k_spmv_warp<true, 2><<<grid, kWarpCtaThreads, smem, st>>>(k, warp_cap);
        } else {
            if (nred == 0) k_spmv_warp<false, 0><<<grid, kWarpCtaThreads, smem, st>>>(k, warp_cap);
            else if (nred == 1) k_spmv_warp<false, 1><<<grid, kWarpCtaThreads, smem, st>>>(k, warp_cap);
            else k_spmv_warp<false, 2><<<grid, kWarpCtaThreads, smem, st>>>(k, warp_cap);
        }
    } else if (variant == 4) {
        const int nblk = (ctx->n + kRowsPerBlock - 1) / kRowsPerBlock;
        k.n_row_blocks = nblk;
        const int stages = (int)ctx->tma_stages;
        const int cap = (int)(((ctx->max_block_nnz + 3) & ~(int64_t)3) + 4);
        const size_t stage_bytes = (size_t)cap * 12 + (size_t)(kRowsPerBlock + 4) * sizeof(label);
        const size_t smem = stage_bytes * stages + sizeof(TmaHdr) * kTmaMaxStages +
                            2 * sizeof(uint64_t) * kTmaMaxStages + 128;
        if (smem > 227 * 1024)
            return fail(ctx, OGL_ERR_UNSUPPORTED, "row blocks too long for the TMA SpMV pipeline");
        // resident CTAs per SM for this shared-memory footprint -> persistent grid
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 2048 / kTmaThreads) per_sm = 2048 / kTmaThreads;
        if (per_sm < 1) per_sm = 1;
        const int64_t want = ctx->stream_ctas > 0 ? ctx->stream_ctas : (int64_t)kNumSM * per_sm;
        const int grid = nblk < want ? nblk : (int)want;
#define TMA_LAUNCH(A, R)                                                                      \
    do {                                                                                      \
        cudaFuncSetAttribute(k_spmv_tma<A, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)smem);                                                      \
        k_spmv_tma<A, R><<<grid, kTmaThreads, smem, st>>>(k, cap, stages);                   \
    } while (0)
        if (sa.advanced) {
            if (nred == 0) TMA_LAUNCH(true, 0);
            else if (nred == 1) TMA_LAUNCH(true, 1);
            else TMA_LAUNCH(true, 2);
        } else {
            if (nred == 0) TMA_LAUNCH(false, 0);
            else if (nred == 1) TMA_LAUNCH(false, 1);
            else TMA_LAUNCH(false, 2);
        }
#undef TMA_LAUNCH
    } else if (variant == 7) {
        OGL_TRY(spmv_ell(ctx, k, sa, ghosted));
        return OGL_OK;
    } else if (variant == 8) {
        return spmv_merge(ctx, k, sa);
    } else if (variant == 2) {
        const int grid = (ctx->n + 255) / 256;
        DISPATCH(k_spmv_scalar, grid, 256, 0);
    } else {
        const int grid = (int)(((int64_t)ctx->n * 32 + 255) / 256);
        DISPATCH(k_spmv_vector, grid, 256, 0);
    }
#undef DISPATCH
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

int spmv_nonlocal(Context *ctx, const double *recv, double *y, double alpha,
                  const double *dot_with, int nred, bool guard_done, int epi,
                  bool inline_epi)
{
    const bool p2p = use_p2p(ctx);
    // peer-memory path: this launch also carries the all-reduce of the fused
    // sums (and the flag handshake), so it runs even without halo rows
    if (ctx->n_nl_rows == 0 && !(p2p && (nred > 0 || ctx->n_targets > 0))) return OGL_OK;
    NonLocalK k;
    k.n_rows = ctx->n_nl_rows;
    k.row_ids = ctx->d_nl_row_ids;
    k.row_ptrs = ctx->d_nl_row_ptrs;
    k.cols = ctx->d_nl_cols;
    k.vals = ctx->d_nl_vals;
    k.recv = recv;
    k.y = y;
    k.alpha = alpha;
    k.dot_with = dot_with;
    k.partials = ctx->d_partials;
    k.ticket = ctx->d_ticket;
    k.state = ctx->d_state;
    k.epi = epi;
    k.inline_epi = inline_epi ? 1 : 0;
    k.guard_done = guard_done ? 1 : 0;
    k.ea = make_epi_args(ctx, nred), k.ea.trace_tag = 20;
    int grid = (ctx->n_nl_rows + 255) / 256;
    if (grid < 1) grid = 1;
    if (nred == 0) k_spmv_nonlocal<0><<<grid, 256, 0, ctx->stream>>>(k);
    else if (nred == 1) k_spmv_nonlocal<1><<<grid, 256, 0, ctx->stream>>>(k);
    else k_spmv_nonlocal<2><<<grid, 256, 0, ctx->stream>>>(k);
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace ogl
