// ELL-family SpMV kernels for sm_100a (spmv_variant 7, `matrixFormat Ell`; SURVEY.md section 8
// a20 and f3, reference: DevicePersistent/CsrMatrixWrapper/CsrMatrixWrapper.H:140-160 builds the
// local and the non-local block in the chosen format).
//
// Layout: slot j of row r at [j * pitch + r], so a warp reads 32 consecutive values per slot and
// gathers x from (for a mesh-ordered matrix) 32 consecutive addresses; no row pointers, no shared
// memory staging, no barrier.  Slots are added left to right in the row's CSR order and padding is
// skipped, products are rounded before the add: bit-identical to the CSR kernels and to the
// sequential reference-executor row sum.
//
// The path is HBM-bound, so the kernels go after the bytes:
//   * pattern-coded columns.  On a mesh-ordered matrix almost every row has one of a handful of
//     (column - row) tuples (7-point stencil: interior, 6 faces, 12 edges, 8 corners = 27).  A
//     device-side hash table finds the distinct tuples once per sparsity pattern; rows whose tuple
//     is among the first 127 carry a ONE-BYTE code, the kernel rebuilds the columns from a table in
//     shared memory: 8 B values + 1/width B per entry instead of 12 B.  Bit 7 of the code says the
//     row also has ghost entries (columns >= n of a decomposed case, stored behind the local
//     ones): those few slots are read with their 4-byte columns, as are all slots of a row whose
//     tuple is not in the table (code 127).
//     Nothing about the arithmetic changes.  Unstructured patterns (> 25% escapes) stay plain ELL.
//   * CG step_1 fused into the SpMV (k_spmv_ell_cgp): the kernel gathers z and p, forms
//     p' = z + (rho/rho_prev) p per operand on the fly -- the same two operations on the same two
//     numbers as the p-update kernel, hence the same bits --, the row owner stores p'[row].  On
//     several ranks the ghost operands come straight from the stamped words the neighbours'
//     x/r-update kernels pushed into this rank's window, and the (single) row that references a
//     ghost column maintains the ghost entry of p.  A PCG iteration is two launches.
#include <climits>
#include <cstring>

#include "spmv.cuh"

namespace ogl {

int spmv_variant_in_use(const Context *ctx);

namespace {

constexpr label kPadDelta = INT_MIN;   // table entry of a padding slot
constexpr int kPatSlots = 1024;        // open-addressing hash table of tuples
constexpr int kPatProbes = 64;         // longest probe sequence (at most 127 tuples are ever coded)
constexpr int kPatMaxW = 8;            // widest row the tuple table holds
constexpr int kEscape = 127;           // low 7 bits of a code: row not in the table
constexpr int kGhostBit = 128;         // the row has ghost entries (columns >= n) behind its local ones
constexpr int kEllBatch = 8;
constexpr int kEllThreads = 256;
constexpr int kEllCtasPerSM = 4;

struct EllK {
    const label *cols;
    const double *vals;
    const unsigned char *code;
    const label *ptab;
    int64_t pitch;
    int width, n_patterns;
};

// ---- construction ------------------------------------------------------------------------

__global__ void k_ell_structure(label n, const label *__restrict__ row_ptrs, const label *__restrict__ cols,
                                int width, int64_t pitch, label *__restrict__ ell_cols)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    const label rs = row_ptrs[row], len = row_ptrs[row + 1] - rs;
    for (int j = 0; j < width; ++j) ell_cols[j * pitch + row] = j < len ? cols[rs + j] : -1;
}

__global__ void k_ell_values(label n, const label *__restrict__ row_ptrs, const double *__restrict__ vals,
                             int width, int64_t pitch, double *__restrict__ ell_vals)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    const label rs = row_ptrs[row], len = row_ptrs[row + 1] - rs;
    for (int j = 0; j < width; ++j) ell_vals[j * pitch + row] = j < len ? vals[rs + j] : 0.0;
}

// (column - row) tuple of the LOCAL entries of a row (a ghosted row keeps its ghost entries,
// columns >= n, behind them: they do not belong to the pattern); `ghost`: the row has some
__device__ __forceinline__ void row_tuple(label n, int64_t row, const label *__restrict__ ell_cols, int width,
                                          int64_t pitch, label (&d)[kPatMaxW], unsigned long long &h, bool &ghost)
{
    ghost = false;
    h = 0xcbf29ce484222325ull;
#pragma unroll
    for (int u = 0; u < kPatMaxW; ++u) {
        label c = -1;
        if (u < width) c = ell_cols[u * pitch + row];
        if (c >= n) {
            ghost = true;
            c = -1;
        }
        d[u] = c < 0 ? kPadDelta : c - (label)row;
        h = (h ^ (unsigned long long)(unsigned int)d[u]) * 0x100000001b3ull;
        h ^= h >> 29;
    }
    if (h == 0) h = 1;   // 0 marks an empty slot
}

// pass 1: every distinct tuple takes a slot (the CAS winner stores the tuple)
__global__ void k_pat_insert(label n, const label *__restrict__ ell_cols, int width, int64_t pitch,
                             unsigned long long *keys, label *tuples, int *overflow)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    label d[kPatMaxW];
    unsigned long long h;
    bool ghost;
    row_tuple(n, row, ell_cols, width, pitch, d, h, ghost);
    // an unstructured pattern has (almost) as many tuples as rows: once the table has overflowed
    // nobody probes any more -- those rows escape, and the format is dropped if they are many
    if (*reinterpret_cast<volatile int *>(overflow)) return;
    unsigned int slot = (unsigned int)(h % kPatSlots);
    for (int probe = 0; probe < kPatProbes; ++probe, slot = (slot + 1) % kPatSlots) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&keys[slot]);
        if (cur == 0) cur = atomicCAS(&keys[slot], 0ull, h);
        if (cur == 0) {   // mine
#pragma unroll
            for (int u = 0; u < kPatMaxW; ++u) tuples[slot * kPatMaxW + u] = d[u];
            return;
        }
        if (cur == h) return;
    }
    *overflow = 1;
}

// pass 2 (one thread): codes 0..126 for the occupied slots in slot order, compact table
__global__ void k_pat_ids(const unsigned long long *keys, const label *tuples, int width, int *ids,
                          label *ptab, int *n_patterns)
{
    int count = 0;
    for (int s = 0; s < kPatSlots; ++s) {
        if (keys[s] == 0) {
            ids[s] = kEscape;
            continue;
        }
        if (count < kEscape) {
            ids[s] = count;
            for (int u = 0; u < width; ++u) ptab[count * width + u] = tuples[s * kPatMaxW + u];
            ++count;
        } else {
            ids[s] = kEscape;
        }
    }
    *n_patterns = count;
}

// pass 3: the row's code; the tuple is compared in full (two tuples may share a hash)
__global__ void k_pat_assign(label n, const label *__restrict__ ell_cols, int width, int64_t pitch,
                             const unsigned long long *__restrict__ keys, const label *__restrict__ tuples,
                             const int *__restrict__ ids, unsigned char *__restrict__ code,
                             unsigned long long *n_escape)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int cd = kEscape;
    if (row < n) {
        label d[kPatMaxW];
        unsigned long long h;
        bool ghost;
        row_tuple(n, row, ell_cols, width, pitch, d, h, ghost);
        unsigned int slot = (unsigned int)(h % kPatSlots);
        for (int probe = 0; probe < kPatProbes; ++probe, slot = (slot + 1) % kPatSlots) {
            const unsigned long long cur = keys[slot];
            if (cur == 0) break;
            if (cur == h) {
                bool same = true;
#pragma unroll
                for (int u = 0; u < kPatMaxW; ++u) same = same && tuples[slot * kPatMaxW + u] == d[u];
                if (same) cd = ids[slot];
                break;
            }
        }
        code[row] = (unsigned char)(cd | (ghost ? kGhostBit : 0));
    }
    const unsigned int esc = __ballot_sync(0xffffffffu, row < n && cd == kEscape);   // (ghost rows with a coded local part do not count)
    if ((threadIdx.x & 31) == 0 && esc) atomicAdd(n_escape, (unsigned long long)__popc(esc));
}

// ---- kernels -------------------------------------------------------------------------------

__device__ __forceinline__ unsigned int ld_code(const unsigned char *p)
{
    unsigned int r;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// Row schedule of a persistent CTA: tiles of kEllThreads rows; `chunk` consecutive tiles per visit
// (consecutive tiles share most of the x lines they gather, so those come from L1 instead of L2),
// visits strided over the grid.  chunk == 1: plain grid-stride.
struct TileWalk {
    int64_t first, step, n_tiles;
    int chunk, k;
    __device__ __forceinline__ TileWalk(label n, int chunk_)
    {
        chunk = chunk_ < 1 ? 1 : chunk_;
        first = (int64_t)blockIdx.x * chunk;
        step = (int64_t)gridDim.x * chunk;
        n_tiles = ((int64_t)n + kEllThreads - 1) / kEllThreads;
        k = 0;
    }
    __device__ __forceinline__ int64_t tile() const { return first + k; }
    __device__ __forceinline__ bool done() const { return first + k >= n_tiles; }
    __device__ __forceinline__ void next()
    {
        if (++k == chunk) {
            k = 0;
            first += step;
        }
    }
};

// Local columns of a pattern-coded row from the shared-memory table.  Padding slots, ghost slots
// and every slot of an escape row get the row's own index (a valid address: the gather is
// unconditional, the slot is skipped in the sum) and a cleared bit in `live`.  The live slots are
// the leading ones of the row.
template <int W>
__device__ __forceinline__ unsigned int coded_columns(label (&c)[W], label row, unsigned int cd, const label *tab)
{
    unsigned int live = 0;
    const unsigned int id = cd & (unsigned int)kEscape;
    const label *t = tab + (id == (unsigned int)kEscape ? 0u : id) * W;
#pragma unroll
    for (int u = 0; u < W; ++u) {
        const label d = t[u];
        const bool on = d != kPadDelta && id != (unsigned int)kEscape;
        c[u] = on ? row + d : row;
        live |= on ? (1u << u) : 0u;
    }
    return live;
}

// does the row owe slots to the tail loop?  (ghost entries behind the coded local ones, or an
// escape row: everything)
__device__ __forceinline__ bool coded_tail(unsigned int cd)
{
    return (cd & (unsigned int)kGhostBit) != 0 || (cd & (unsigned int)kEscape) == (unsigned int)kEscape;
}

// W > 0: compile-time row width -- all loads of a row are issued before the first dependent
// instruction.  W == 0: any width (plain columns only).
//
// PC (pattern-coded columns), the hot configuration, is software-pipelined: the code and the W
// values of the thread's NEXT row are requested while this row's operands are gathered, so a
// trip costs one L2 round trip (the gather) with the HBM stream running underneath; the code of
// the current row arrived a trip ago, so the gather addresses need no HBM round trip either.
// Rows behind the escape code (ghost columns, irregular rows) are rare and take a plain loop.
template <bool ADV, int NRED, int W, bool PC, int MINB>
__global__ void __launch_bounds__(kEllThreads, MINB)
k_spmv_ell(const SpmvK a, const EllK m, const int chunk)
{
    __shared__ label tab[PC ? 128 * (W > 0 ? W : 1) : 1];
    if (a.guard_done && a.state->done) return;
    if (PC) {
        for (int i = threadIdx.x; i < m.n_patterns * W; i += kEllThreads) tab[i] = m.ptab[i];
        __syncthreads();
    }
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    if (W > 0 && PC) {
        constexpr int WW = W > 0 ? W : 1;
        TileWalk w(a.n, chunk);
        double vn[WW];
        unsigned int cdn = kEscape;
        int64_t row = w.tile() * kEllThreads + threadIdx.x;
        bool ok = !w.done() && row < a.n;
        if (ok) {
            cdn = ld_code(m.code + row);
#pragma unroll
            for (int u = 0; u < WW; ++u) vn[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
        }
        while (!w.done()) {
            const int64_t row_c = row;
            const bool ok_c = ok;
            const unsigned int cd = cdn;
            double v[WW], xv[WW];
            label c[WW];
            // gather for this row: the addresses come from the code that arrived a trip ago
            unsigned int live = 0;
            if (ok_c) {
                live = coded_columns<WW>(c, (label)row_c, cd, tab);
#pragma unroll
                for (int u = 0; u < WW; ++u) xv[u] = __ldg(&a.x[c[u]]);
            }
#pragma unroll
            for (int u = 0; u < WW; ++u) v[u] = vn[u];
            // next row: code + values on their way while the gather is in flight
            w.next();
            row = w.tile() * kEllThreads + threadIdx.x;
            ok = !w.done() && row < a.n;
            if (ok) {
                cdn = ld_code(m.code + row);
#pragma unroll
                for (int u = 0; u < WW; ++u) vn[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
            }
            if (!ok_c) continue;
            double dw = 0.0;
            if (NRED >= 1) asm volatile("ld.global.f64 %0, [%1];" : "=d"(dw) : "l"(a.dot_with + row_c));
            double sum = ADV ? __dmul_rn(a.beta, a.y_in[row_c]) : 0.0;
#pragma unroll
            for (int u = 0; u < WW; ++u) {
                const double s2 = __dadd_rn(sum, prod_of(v[u], xv[u], a.alpha, ADV));
                sum = (live >> u) & 1u ? s2 : sum;
            }
            if (coded_tail(cd)) {
                // the slots behind the coded ones (x carries its ghost part here), all requested
                // at once; rare (rows on a processor boundary, irregular rows)
                const int first = __popc(live);
                label ce[WW];
                double ve[WW], xe[WW];
#pragma unroll
                for (int u = 0; u < WW; ++u) ce[u] = u >= first ? m.cols[u * m.pitch + row_c] : -1;
#pragma unroll
                for (int u = 0; u < WW; ++u) ve[u] = u >= first ? m.vals[u * m.pitch + row_c] : 0.0;
#pragma unroll
                for (int u = 0; u < WW; ++u) xe[u] = ce[u] >= 0 ? a.x[ce[u]] : 0.0;
#pragma unroll
                for (int u = 0; u < WW; ++u)
                    if (ce[u] >= 0) sum = __dadd_rn(sum, prod_of(ve[u], xe[u], a.alpha, ADV));
            }
            a.y[row_c] = sum;
            if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
            if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
        }
    } else if (W > 0) {
        constexpr int WW = W > 0 ? W : 1;
        TileWalk w(a.n, chunk);
        for (; !w.done(); w.next()) {
            const int64_t row = w.tile() * kEllThreads + threadIdx.x;
            if (row >= a.n) continue;
            // everything the row needs is requested up front: a warp issues in order, so a load
            // placed behind the row sum would add its whole latency to every trip
            double dw = 0.0;
            if (NRED >= 1) asm volatile("ld.global.f64 %0, [%1];" : "=d"(dw) : "l"(a.dot_with + row));
            double sum = ADV ? __dmul_rn(a.beta, a.y_in[row]) : 0.0;
            label c[WW];
            double v[WW], xv[WW];
#pragma unroll
            for (int u = 0; u < WW; ++u) c[u] = ld_mat(&m.cols[u * m.pitch + row], a.mat_policy);
#pragma unroll
            for (int u = 0; u < WW; ++u) v[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
#pragma unroll
            for (int u = 0; u < WW; ++u) xv[u] = c[u] >= 0 ? __ldg(&a.x[c[u]]) : 0.0;
#pragma unroll
            for (int u = 0; u < WW; ++u)
                if (c[u] >= 0) sum = __dadd_rn(sum, prod_of(v[u], xv[u], a.alpha, ADV));
            a.y[row] = sum;
            if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
            if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
        }
    } else {
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < a.n; row += stride) {
            double dw = 0.0;
            if (NRED >= 1) asm volatile("ld.global.f64 %0, [%1];" : "=d"(dw) : "l"(a.dot_with + row));
            double sum = ADV ? __dmul_rn(a.beta, a.y_in[row]) : 0.0;
            for (int j0 = 0; j0 < m.width; j0 += kEllBatch) {
                label c[kEllBatch];
                double v[kEllBatch], xv[kEllBatch];
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u)
                    c[u] = j0 + u < m.width ? ld_mat(&m.cols[(j0 + u) * m.pitch + row], a.mat_policy) : -1;
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u)
                    v[u] = j0 + u < m.width ? ld_mat(&m.vals[(j0 + u) * m.pitch + row], a.mat_policy) : 0.0;
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u) xv[u] = c[u] >= 0 ? __ldg(&a.x[c[u]]) : 0.0;
#pragma unroll
                for (int u = 0; u < kEllBatch; ++u)
                    if (c[u] >= 0) sum = __dadd_rn(sum, prod_of(v[u], xv[u], a.alpha, ADV));
            }
            a.y[row] = sum;
            if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
            if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
        }
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// ---------------------------------------------------------------------------------------------
// TMA-fed variant of the pattern-coded kernel (option ell_tma).  The value stream never touches the
// load/store units: one producer lane per CTA walks the CTA's tiles ahead of the consumers and
// issues W + 1 bulk copies per 256-row tile (cp.async.bulk = the TMA engine: W slot slices of
// 2 KB and the 256 row codes) into a ring of shared-memory stages guarded by mbarriers; the 8
// consumer warps wait on a stage, read code and values from shared memory (conflict-free: thread t
// owns element t of every slice), gather x, add the row left to right, and hand the stage back.
// Bytes in flight per SM are set by the ring (3 CTAs x stages x 14.6 KB), not by how many loads a
// warp can have outstanding.  Same arithmetic and order as k_spmv_ell: bit-identical results.
// ---------------------------------------------------------------------------------------------
constexpr int kTmaEllThreads = kEllThreads + 32;   // 8 consumer warps + 1 producer warp
constexpr int kTmaEllCtasPerSM = 4;      // <= 56 registers; the ring depth decides how many fit (option ell_minb caps the grid)
constexpr int kTmaEllMaxStages = 4;

template <int NRED, int W>
__global__ void __launch_bounds__(kTmaEllThreads, kTmaEllCtasPerSM)
k_spmv_ell_tma(const SpmvK a, const EllK m, const int stages)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kStageBytes = W * kEllThreads * 8 + kEllThreads;   // W value slices + the codes
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * kStageBytes);
    uint64_t *empty = full + kTmaEllMaxStages;
    label *tab = reinterpret_cast<label *>(empty + kTmaEllMaxStages);
    if (a.guard_done && a.state->done) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int st = 0; st < stages; ++st) {
            tma::mbar_init(&full[st], 1);
            tma::mbar_init(&empty[st], kEllThreads / 32);
        }
        tma::fence_barrier_init();
        tma::fence_proxy_async();
    }
    for (int i = tid; i < m.n_patterns * W; i += kTmaEllThreads) tab[i] = m.ptab[i];
    __syncthreads();
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) red[j] = 0.0;
    const int64_t n_tiles = ((int64_t)a.n + kEllThreads - 1) / kEllThreads;
    if (warp == kEllThreads / 32) {
        // ===== producer: one lane feeds the ring =====
        if (lane == 0) {
            const uint64_t pol = tma::policy_evict_first();
            int it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int st = it % stages;
                const uint32_t round = (uint32_t)(it / stages);
                if (round > 0) tma::mbar_wait(&empty[st], (round - 1) & 1);
                const int64_t r0 = t * kEllThreads;
                const int64_t left = m.pitch - r0;                       // pitch is a multiple of 32 rows
                const uint32_t rows = (uint32_t)(left < kEllThreads ? left : kEllThreads);
                unsigned char *base = smem_raw + (size_t)st * kStageBytes;
                tma::mbar_expect_tx(&full[st], rows * 8u * W + rows);
#pragma unroll
                for (int u = 0; u < W; ++u)
                    tma::bulk_load(base + (size_t)u * kEllThreads * 8, m.vals + u * m.pitch + r0, rows * 8u, &full[st], pol);
                tma::bulk_load(base + (size_t)W * kEllThreads * 8, m.code + r0, rows, &full[st], pol);
            }
        }
        __syncwarp();   // reconverge the producer warp before the block-wide reduction
    } else {
        // ===== consumers: thread t owns row t of every tile =====
        int it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int st = it % stages;
            const uint32_t round = (uint32_t)(it / stages);
            const int64_t row = t * kEllThreads + tid;
            // the fused dot's operand does not depend on the stage: in flight during the wait
            double dw = 0.0;
            if (NRED >= 1 && row < a.n) asm volatile("ld.global.f64 %0, [%1];" : "=d"(dw) : "l"(a.dot_with + row));
            tma::mbar_wait(&full[st], round & 1);
            const unsigned char *base = smem_raw + (size_t)st * kStageBytes;
            const double *vs = reinterpret_cast<const double *>(base);
            if (row < a.n) {
                const unsigned int cd = base[(size_t)W * kEllThreads * 8 + tid];
                label c[W];
                double v[W], xv[W];
                const unsigned int live = coded_columns<W>(c, (label)row, cd, tab);
                // a row on a processor boundary owes one slot (rarely more) behind the coded ones: its
                // column is requested together with the gathers, its operand right behind -- one extra
                // memory round trip for the warp instead of two
                const int first = __popc(live);
                const bool tail = coded_tail(cd) && first < W;
                label ce0 = -1;
                if (tail) ce0 = m.cols[first * m.pitch + row];
#pragma unroll
                for (int u = 0; u < W; ++u) xv[u] = __ldg(&a.x[c[u]]);
#pragma unroll
                for (int u = 0; u < W; ++u) v[u] = vs[u * kEllThreads + tid];
                double xe0 = 0.0;
                if (ce0 >= 0) xe0 = a.x[ce0];
                double sum = 0.0;
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const double s2 = __dadd_rn(sum, __dmul_rn(v[u], xv[u]));
                    sum = (live >> u) & 1u ? s2 : sum;
                }
                if (ce0 >= 0) {
                    sum = __dadd_rn(sum, __dmul_rn(vs[first * kEllThreads + tid], xe0));
#pragma unroll 1
                    for (int u = first + 1; u < W; ++u) {
                        const label ce = m.cols[u * m.pitch + row];
                        if (ce < 0) break;
                        sum = __dadd_rn(sum, __dmul_rn(vs[u * kEllThreads + tid], a.x[ce]));
                    }
                }
                a.y[row] = sum;
                if (NRED >= 1) red[0] = __dadd_rn(red[0], __dmul_rn(dw, sum));
                if (NRED >= 2) red[1] = __dadd_rn(red[1], __dmul_rn(sum, sum));
            }
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&empty[st]);
        }
    }
    if (NRED > 0)
        grid_reduce<(NRED > 0 ? NRED : 1)>(red, a.partials, a.ticket, a.state, 0, a.epi,
                                           a.inline_epi != 0, a.ea);
}

// The fused CG kernel (k_spmv_ell_cgp below) on the same TMA ring: values and codes arrive through
// shared memory, the load/store units only see the two gathers (z, p) and the three row-local
// accesses.  x = z (or r), y_in = p (previous), y = q, p_new = the other p buffer.
template <int W, bool GHOST>
__global__ void __launch_bounds__(kTmaEllThreads, kTmaEllCtasPerSM)
k_spmv_ell_cgp_tma(const SpmvK a, const EllK m, double *__restrict__ p_new, const int stages)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kStageBytes = W * kEllThreads * 8 + kEllThreads;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * kStageBytes);
    uint64_t *empty = full + kTmaEllMaxStages;
    label *tab = reinterpret_cast<label *>(empty + kTmaEllMaxStages);
    if (a.guard_done && a.state->done) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int st = 0; st < stages; ++st) {
            tma::mbar_init(&full[st], 1);
            tma::mbar_init(&empty[st], kEllThreads / 32);
        }
        tma::fence_barrier_init();
        tma::fence_proxy_async();
    }
    for (int i = tid; i < m.n_patterns * W; i += kTmaEllThreads) tab[i] = m.ptab[i];
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) trace_event(a.ea, 0);
    const bool p_is_z = a.state->flag_p_is_z != 0;
    const double t = a.state->coef_p;
    const unsigned long long *zg = nullptr;
    unsigned long long stamp = 0;
    long long t0 = 0;
    if (GHOST) {
        const CommDev *cm = a.ea.comm;
        zg = reinterpret_cast<const unsigned long long *>(cm->my_recv + 2 * (size_t)cm->my_recv_stride);
        stamp = stamp_of(ld_ar_seq(cm));
        t0 = clock64();
    }
    double red[1] = {0.0};
    const int64_t n_tiles = ((int64_t)a.n + kEllThreads - 1) / kEllThreads;
    if (warp == kEllThreads / 32) {
        if (lane == 0) {
            const uint64_t pol = tma::policy_evict_first();
            int it = 0;
            for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++it) {
                const int st = it % stages;
                const uint32_t round = (uint32_t)(it / stages);
                if (round > 0) tma::mbar_wait(&empty[st], (round - 1) & 1);
                const int64_t r0 = tl * kEllThreads;
                const int64_t left = m.pitch - r0;
                const uint32_t rows = (uint32_t)(left < kEllThreads ? left : kEllThreads);
                unsigned char *base = smem_raw + (size_t)st * kStageBytes;
                tma::mbar_expect_tx(&full[st], rows * 8u * W + rows);
#pragma unroll
                for (int u = 0; u < W; ++u)
                    tma::bulk_load(base + (size_t)u * kEllThreads * 8, m.vals + u * m.pitch + r0, rows * 8u, &full[st], pol);
                tma::bulk_load(base + (size_t)W * kEllThreads * 8, m.code + r0, rows, &full[st], pol);
            }
        }
        __syncwarp();
    } else {
        int it = 0;
        for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++it) {
            const int st = it % stages;
            const uint32_t round = (uint32_t)(it / stages);
            const int64_t row = tl * kEllThreads + tid;
            tma::mbar_wait(&full[st], round & 1);
            const unsigned char *base = smem_raw + (size_t)st * kStageBytes;
            const double *vs = reinterpret_cast<const double *>(base);
            if (row < a.n) {
                const unsigned int cd = base[(size_t)W * kEllThreads * 8 + tid];
                label c[W];
                double zc[W], pc[W];
                const unsigned int live = coded_columns<W>(c, (label)row, cd, tab);
#pragma unroll
                for (int u = 0; u < W; ++u) zc[u] = __ldg(&a.x[c[u]]);
                if (!p_is_z) {
#pragma unroll
                    for (int u = 0; u < W; ++u) pc[u] = a.y_in[c[u]];
                }
                double sum = 0.0, mine = 0.0;
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const double pv = p_is_z ? zc[u] : __dadd_rn(zc[u], __dmul_rn(t, pc[u]));
                    const bool on = (live >> u) & 1u;
                    if (on && c[u] == (label)row) mine = pv;
                    const double s2 = __dadd_rn(sum, __dmul_rn(vs[u * kEllThreads + tid], pv));
                    sum = on ? s2 : sum;
                }
                if (coded_tail(cd)) {
#pragma unroll 1
                    for (int u = __popc(live); u < W; ++u) {
                        const label ce = m.cols[u * m.pitch + row];
                        if (ce < 0) break;
                        double zv = 0.0;
                        if (GHOST && ce >= a.n) {
                            if (!pull_stamped(zg + 2 * (size_t)(ce - a.n), stamp, t0, zv)) a.state->comm_error = 1;
                        } else {
                            zv = a.x[ce];
                        }
                        const double pv = p_is_z ? zv : __dadd_rn(zv, __dmul_rn(t, a.y_in[ce]));
                        if (ce == (label)row) mine = pv;
                        if (GHOST && ce >= a.n) p_new[ce] = pv;
                        sum = __dadd_rn(sum, __dmul_rn(vs[u * kEllThreads + tid], pv));
                    }
                }
                p_new[row] = mine;
                a.y[row] = sum;
                red[0] = __dadd_rn(red[0], __dmul_rn(mine, sum));
            }
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&empty[st]);
        }
    }
    grid_reduce<1>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
}

// CG step_1 fused into the SpMV:  x = z (or r), y_in = p (previous), y = q, p_new = the other p
// buffer.  GHOST (several ranks, peer-memory path): columns >= n are ghost operands -- z from the
// stamped words in slot 2 of this rank's window (pushed by the neighbours' k_cg_xr, stamped with
// the number of the all-reduce that kernel ended with), p from / to the ghost part of the p
// buffers, which the one row referencing the column keeps up to date.  Rows with ghost columns
// are escape rows of the pattern code; like all escape rows they take the plain loop at the end
// of a trip (ghost entries sit BEHIND the local ones in a row: the same left-to-right sum).
template <int W, bool PC, bool GHOST, int MINB>
__global__ void __launch_bounds__(kEllThreads, MINB)
k_spmv_ell_cgp(const SpmvK a, const EllK m, double *__restrict__ p_new, const int chunk)
{
    __shared__ label tab[PC ? 128 * W : 1];
    if (a.guard_done && a.state->done) return;
    if (PC) {
        for (int i = threadIdx.x; i < m.n_patterns * W; i += kEllThreads) tab[i] = m.ptab[i];
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(a.ea, 0);
    const bool p_is_z = a.state->flag_p_is_z != 0;
    const double t = a.state->coef_p;
    const unsigned long long *zg = nullptr;
    unsigned long long stamp = 0;
    long long t0 = 0;
    if (GHOST) {
        const CommDev *cm = a.ea.comm;
        zg = reinterpret_cast<const unsigned long long *>(cm->my_recv + 2 * (size_t)cm->my_recv_stride);
        stamp = stamp_of(ld_ar_seq(cm));
        t0 = clock64();
    }
    double red[1] = {0.0};
    TileWalk w(a.n, chunk);
    double vn[W];
    unsigned int cdn = kEscape;
    int64_t row = w.tile() * kEllThreads + threadIdx.x;
    bool ok = !w.done() && row < a.n;
    if (PC && ok) {
        cdn = ld_code(m.code + row);
#pragma unroll
        for (int u = 0; u < W; ++u) vn[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
    }
    while (!w.done()) {
        const int64_t row_c = row;
        const bool ok_c = ok;
        const unsigned int cd = cdn;
        double v[W], zc[W], pc[W];
        label c[W];
        unsigned int live = 0;
        if (PC) {
            if (ok_c) {
                live = coded_columns<W>(c, (label)row_c, cd, tab);
#pragma unroll
                for (int u = 0; u < W; ++u) zc[u] = __ldg(&a.x[c[u]]);
                if (!p_is_z) {
#pragma unroll
                    for (int u = 0; u < W; ++u) pc[u] = a.y_in[c[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < W; ++u) v[u] = vn[u];
        }
        w.next();
        row = w.tile() * kEllThreads + threadIdx.x;
        ok = !w.done() && row < a.n;
        if (PC && ok) {
            cdn = ld_code(m.code + row);
#pragma unroll
            for (int u = 0; u < W; ++u) vn[u] = ld_mat(&m.vals[u * m.pitch + row], a.mat_policy);
        }
        if (!ok_c) continue;
        bool tail = PC && coded_tail(cd);   // slots the tail below still owes
        if (!PC) {
#pragma unroll
            for (int u = 0; u < W; ++u) c[u] = ld_mat(&m.cols[u * m.pitch + row_c], a.mat_policy);
#pragma unroll
            for (int u = 0; u < W; ++u) v[u] = ld_mat(&m.vals[u * m.pitch + row_c], a.mat_policy);
#pragma unroll
            for (int u = 0; u < W; ++u) {
                bool on = c[u] >= 0;
                if (GHOST && c[u] >= a.n) {   // ghost entries follow the local ones: the tail takes them
                    on = false;
                    tail = true;
                }
                live |= on ? (1u << u) : 0u;
                if (!on) c[u] = (label)row_c;
                zc[u] = __ldg(&a.x[c[u]]);
            }
            if (!p_is_z) {
#pragma unroll
                for (int u = 0; u < W; ++u) pc[u] = a.y_in[c[u]];
            }
        }
        double sum = 0.0, mine = 0.0;
#pragma unroll
        for (int u = 0; u < W; ++u) {
            const double pv = p_is_z ? zc[u] : __dadd_rn(zc[u], __dmul_rn(t, pc[u]));
            const bool on = (live >> u) & 1u;
            if (on && c[u] == (label)row_c) mine = pv;               // the diagonal slot: my own p'
            const double s2 = __dadd_rn(sum, __dmul_rn(v[u], pv));
            sum = on ? s2 : sum;
        }
        if (tail) {
            // the slots behind the handled ones: ghost operands (z from the stamped words of the
            // window, p from / to the ghost part of the p buffers) and, for an escape row, its local
            // ones.  A face row of a decomposed stencil has ONE such slot, so a plain loop that
            // stops at the first padding slot is as fast as it gets and costs no registers.
#pragma unroll 1
            for (int u = __popc(live); u < W; ++u) {
                const label ce = m.cols[u * m.pitch + row_c];
                if (ce < 0) break;
                double zv = 0.0;
                if (GHOST && ce >= a.n) {
                    if (!pull_stamped(zg + 2 * (size_t)(ce - a.n), stamp, t0, zv)) a.state->comm_error = 1;
                } else {
                    zv = a.x[ce];
                }
                const double pv = p_is_z ? zv : __dadd_rn(zv, __dmul_rn(t, a.y_in[ce]));
                if (ce == (label)row_c) mine = pv;
                if (GHOST && ce >= a.n) p_new[ce] = pv;   // ghost entry of p: referenced by this row only
                sum = __dadd_rn(sum, __dmul_rn(m.vals[u * m.pitch + row_c], pv));
            }
        }
        p_new[row_c] = mine;
        a.y[row_c] = sum;
        red[0] = __dadd_rn(red[0], __dmul_rn(mine, sum));
    }
    grid_reduce<1>(red, a.partials, a.ticket, a.state, 0, a.epi, a.inline_epi != 0, a.ea);
}

}  // namespace

// --------------------------------------------------------------------------------------------

void ell_invalidate(Context *ctx, bool structure)
{
    for (Context::EllMatrix *e : {&ctx->ell, &ctx->gell}) {
        e->ready = false;
        if (structure) {
            e->structure_ready = false;
            e->coded = false;
            e->n_patterns = 0;
            e->n_escape = 0;
        }
    }
}

// find the distinct (column - row) tuples and give every row its code
static int ell_detect_patterns(Context *ctx, Context::EllMatrix &e)
{
    e.coded = false;
    e.n_patterns = 0;
    e.n_escape = 0;
    if (ctx->ell_coded == 0 || e.width > kPatMaxW || ctx->n == 0) return OGL_OK;
    cudaStream_t st = ctx->stream;
    unsigned long long *d_keys = nullptr, *d_esc = nullptr;
    label *d_tuples = nullptr;
    int *d_ids = nullptr, *d_misc = nullptr;
    auto cleanup = [&]() { cudaFree(d_keys), cudaFree(d_esc), cudaFree(d_tuples), cudaFree(d_ids), cudaFree(d_misc); };
    int rc;
    if ((rc = dev_alloc(ctx, &d_keys, kPatSlots)) || (rc = dev_alloc(ctx, &d_esc, 1)) ||
        (rc = dev_alloc(ctx, &d_tuples, kPatSlots * kPatMaxW)) || (rc = dev_alloc(ctx, &d_ids, kPatSlots)) ||
        (rc = dev_alloc(ctx, &d_misc, 2))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_keys, 0, sizeof(unsigned long long) * kPatSlots, st);
    cudaMemsetAsync(d_esc, 0, sizeof(unsigned long long), st);
    cudaMemsetAsync(d_misc, 0, 2 * sizeof(int), st);
    const int grid = (int)(((int64_t)ctx->n + 255) / 256);
    k_pat_insert<<<grid, 256, 0, st>>>(ctx->n, e.cols, e.width, e.pitch, d_keys, d_tuples, d_misc);
    k_pat_ids<<<1, 1, 0, st>>>(d_keys, d_tuples, e.width, d_ids, e.ptab, d_misc + 1);
    k_pat_assign<<<grid, 256, 0, st>>>(ctx->n, e.cols, e.width, e.pitch, d_keys, d_tuples, d_ids, e.code, d_esc);
    ctx->launches += 3;
    int misc[2] = {0, 0};
    unsigned long long esc = 0;
    cudaMemcpyAsync(misc, d_misc, sizeof(misc), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&esc, d_esc, sizeof(esc), cudaMemcpyDeviceToHost, st);
    cudaError_t err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = cudaGetLastError();
    cleanup();
    if (err != cudaSuccess) return fail(ctx, OGL_ERR_CUDA, std::string("ell_detect_patterns: ") + cudaGetErrorString(err));
    e.n_patterns = misc[1];
    e.n_escape = (int64_t)esc;
    // a full table (misc[0]) only means more escapes; the codes that were assigned are exact
    const bool worth = e.n_patterns > 0 && (ctx->ell_coded == 2 || 4 * e.n_escape <= (int64_t)ctx->n);
    e.coded = worth;
    return OGL_OK;
}

// (re)build the ELL copy: structure once per pattern, values once per coefficient update; never
// inside a graph capture
static int ell_prepare(Context *ctx, bool ghosted)
{
    Context::EllMatrix &e = ghosted ? ctx->gell : ctx->ell;
    if (e.structure_ready && e.ready) return OGL_OK;
    if (ctx->capturing) return fail(ctx, OGL_ERR_INVALID, "ELL matrix not built before the graph capture");
    const int width = (int)(ghosted ? ctx->max_row_len_g : ctx->max_row_len);
    const int64_t nnz = ghosted ? ctx->nnz + ctx->n_halo : ctx->nnz;
    if (width < 1 || width > 64 || (int64_t)width * ctx->n > 3 * nnz)
        return fail(ctx, OGL_ERR_UNSUPPORTED, "rows too long or too irregular for the ELL format");
    const int64_t pitch = ((int64_t)ctx->n + 31) / 32 * 32;
    const label *row_ptrs = ghosted ? ctx->d_g_row_ptrs : ctx->d_row_ptrs;
    const int grid = (int)(((int64_t)ctx->n + 255) / 256);
    if (!e.structure_ready) {
        if (e.width != width || e.pitch != pitch || !e.cols) {
            OGL_TRY(dev_alloc(ctx, &e.cols, (size_t)(width * pitch)));
            OGL_TRY(dev_alloc(ctx, &e.vals, (size_t)(width * pitch)));
            OGL_TRY(dev_alloc(ctx, &e.code, (size_t)pitch));
            OGL_TRY(dev_alloc(ctx, &e.ptab, (size_t)256 * (width < kPatMaxW ? kPatMaxW : width)));
            e.width = width;
            e.pitch = pitch;
            invalidate_graph(ctx);   // a captured chunk holds the old addresses
        }
        k_ell_structure<<<grid, 256, 0, ctx->stream>>>(ctx->n, row_ptrs, ghosted ? ctx->d_g_cols : ctx->d_cols,
                                                      width, pitch, e.cols);
        ctx->launches++;
        const bool was_coded = e.coded;
        OGL_TRY(ell_detect_patterns(ctx, e));
        if (was_coded != e.coded) invalidate_graph(ctx);
        e.structure_ready = true;
        e.ready = false;
    }
    if (!e.ready) {
        k_ell_values<<<grid, 256, 0, ctx->stream>>>(ctx->n, row_ptrs, ghosted ? ctx->d_g_vals : ctx->d_vals, width,
                                                   pitch, e.vals);
        ctx->launches++;
        OGL_CUDA(ctx, cudaGetLastError());
        e.ready = true;
    }
    return OGL_OK;
}

// Build whatever ELL copy the CG loop is going to use, outside the graph capture (on several
// ranks the prologue's SpMVs run over the CSR with the flag handshake and would not build it).
int ell_prepare_for_loop(Context *ctx, bool ghost)
{
    if (spmv_variant_in_use(ctx) != 7) return OGL_OK;
    const int rc = ell_prepare(ctx, ghost && ctx->have_ghosted);
    return rc == OGL_ERR_UNSUPPORTED ? OGL_OK : rc;   // the launch reports it if it really needs the copy
}

static EllK ell_args(const Context::EllMatrix &e)
{
    EllK m;
    m.cols = e.cols;
    m.vals = e.vals;
    m.code = e.code;
    m.ptab = e.ptab;
    m.pitch = e.pitch;
    m.width = e.width;
    m.n_patterns = e.n_patterns;
    return m;
}

static int ell_grid(const Context *ctx, int ctas_per_sm = kEllCtasPerSM)
{
    const int64_t need = ((int64_t)ctx->n + kEllThreads - 1) / kEllThreads;
    const int64_t cap = ctx->stream_ctas > 0 ? ctx->stream_ctas : (int64_t)kNumSM * ctas_per_sm;   // persistent
    const int64_t g = need < cap ? need : cap;
    return g < 1 ? 1 : (int)g;
}

int spmv_ell(Context *ctx, SpmvK &k, const SpmvArgs &sa, bool ghosted)
{
    OGL_TRY(ell_prepare(ctx, ghosted));
    const Context::EllMatrix &e = ghosted ? ctx->gell : ctx->ell;
    if (sa.ghost_x) k.ea = make_epi_args(ctx, sa.nred), k.ea.trace_tag = 20;   // all-reduce inside the launch
    const EllK m = ell_args(e);
    cudaStream_t st = ctx->stream;
    const int nred = sa.nred;
    const bool pc = e.coded && (e.width == 7 || e.width == 5);
    const int chunk = (int)ctx->ell_chunk;
    const int minb = (int)ctx->ell_minb;
    // auto: above 2 M rows; over a ghosted matrix already above 256 k (its boundary-row tails cost the
    // TMA-fed kernel less: 38.6 vs 39.2 us per iteration at 1 M rows per GPU on 2 GPUs)
    const bool tma_on = ctx->ell_tma == 1 ||
                        (ctx->ell_tma == 2 && (ctx->n > (1 << 21) || (ghosted && ctx->n > (1 << 18))));
    if (tma_on && pc && e.width == 7 && !sa.advanced) {
        // TMA-fed ring (k_spmv_ell_tma)
        int stages = (int)ctx->tma_stages;
        if (stages < 2) stages = 2;
        if (stages > kTmaEllMaxStages) stages = kTmaEllMaxStages;
        const size_t smem = (size_t)stages * (7 * kEllThreads * 8 + kEllThreads) +
                            2 * kTmaEllMaxStages * sizeof(uint64_t) + 128 * 7 * sizeof(label);
        int per_sm = (int)((227 * 1024) / (smem + 1024));      // resident CTAs for this ring depth
        if (per_sm > kTmaEllCtasPerSM) per_sm = kTmaEllCtasPerSM;
        if (per_sm > minb) per_sm = minb;
        if (per_sm < 1) per_sm = 1;
        const int grid_t = ell_grid(ctx, per_sm);
#define TMA_GO(R)                                                                                              \
    do {                                                                                                       \
        cudaFuncSetAttribute(k_spmv_ell_tma<R, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        k_spmv_ell_tma<R, 7><<<grid_t, kTmaEllThreads, smem, st>>>(k, m, stages);                              \
    } while (0)
        if (nred == 0) TMA_GO(0);
        else if (nred == 1) TMA_GO(1);
        else TMA_GO(2);
#undef TMA_GO
        ctx->launches++;
        OGL_CUDA(ctx, cudaGetLastError());
        return OGL_OK;
    }
    const int grid = ell_grid(ctx, pc ? minb : kEllCtasPerSM);
#define ELL_GO(A, R, W, P, B) k_spmv_ell<A, R, W, P, B><<<grid, kEllThreads, 0, st>>>(k, m, chunk)
#define ELL_WP(A, R, W)                                       \
    do {                                                      \
        if (pc) {                                             \
            if (minb == 3) ELL_GO(A, R, W, true, 3);          \
            else ELL_GO(A, R, W, true, 4);                    \
        } else {                                              \
            ELL_GO(A, R, W, false, 4);                        \
        }                                                     \
    } while (0)
#define ELL_W(A, R)                                           \
    do {                                                      \
        if (e.width == 7) ELL_WP(A, R, 7);                    \
        else if (e.width == 5) ELL_WP(A, R, 5);               \
        else if (e.width == 8) ELL_GO(A, R, 8, false, 4);     \
        else ELL_GO(A, R, 0, false, 4);                       \
    } while (0)
    if (sa.advanced) {
        if (nred == 0) ELL_W(true, 0);
        else if (nred == 1) ELL_W(true, 1);
        else ELL_W(true, 2);
    } else {
        if (nred == 0) ELL_W(false, 0);
        else if (nred == 1) ELL_W(false, 1);
        else ELL_W(false, 2);
    }
#undef ELL_W
#undef ELL_WP
#undef ELL_GO
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

// q = A p', p' = z + coef_p p, <p',q>, CG_BETA epilogue -- one launch (see k_spmv_ell_cgp).
// Returns OGL_ERR_UNSUPPORTED when the fused form does not apply (caller falls back).
int spmv_ell_cgp(Context *ctx, const double *z, const double *p_old, double *p_new, double *q, bool ghost)
{
    if (spmv_variant_in_use(ctx) != 7) return OGL_ERR_UNSUPPORTED;
    if (!ghost && ctx->n_ranks != 1) return OGL_ERR_UNSUPPORTED;
    // a rank without halo rows has no ghosted matrix: it runs its local one, all-reduce included
    const bool use_g = ghost && ctx->have_ghosted;
    {
        // the fused kernel exists for the stencil widths only: decide before building anything
        const int width = (int)(use_g ? ctx->max_row_len_g : ctx->max_row_len);
        if (width != 7 && width != 5) return OGL_ERR_UNSUPPORTED;
    }
    OGL_TRY(ell_prepare(ctx, use_g));
    const Context::EllMatrix &e = use_g ? ctx->gell : ctx->ell;
    SpmvK k;
    std::memset(&k, 0, sizeof(k));
    k.x = z;
    k.y_in = p_old;
    k.y = q;
    k.n = ctx->n;
    k.mat_policy = spmv_l2_policy(ctx);
    k.partials = ctx->d_partials;
    k.ticket = ctx->d_ticket;
    k.state = ctx->d_state;
    k.epi = EPI_CG_BETA;
    k.inline_epi = 1;
    k.guard_done = 1;
    k.ea = make_epi_args(ctx, ghost ? 1 : 0);
    k.ea.trace_tag = 20;
    const EllK m = ell_args(e);
    const bool tma_on = ctx->ell_tma == 1 || (ctx->ell_tma == 2 && ctx->n > (1 << 21));
    if (tma_on && e.coded && e.width == 7) {
        int stages = (int)ctx->tma_stages;
        if (stages < 2) stages = 2;
        if (stages > kTmaEllMaxStages) stages = kTmaEllMaxStages;
        const size_t smem = (size_t)stages * (7 * kEllThreads * 8 + kEllThreads) +
                            2 * kTmaEllMaxStages * sizeof(uint64_t) + 128 * 7 * sizeof(label);
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > kTmaEllCtasPerSM) per_sm = kTmaEllCtasPerSM;
        if (per_sm > (int)ctx->ell_minb) per_sm = (int)ctx->ell_minb;
        if (per_sm < 1) per_sm = 1;
        const int grid_t = ell_grid(ctx, per_sm);
        if (ghost) {
            cudaFuncSetAttribute(k_spmv_ell_cgp_tma<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_spmv_ell_cgp_tma<7, true><<<grid_t, kTmaEllThreads, smem, ctx->stream>>>(k, m, p_new, stages);
        } else {
            cudaFuncSetAttribute(k_spmv_ell_cgp_tma<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_spmv_ell_cgp_tma<7, false><<<grid_t, kTmaEllThreads, smem, ctx->stream>>>(k, m, p_new, stages);
        }
        ctx->launches++;
        OGL_CUDA(ctx, cudaGetLastError());
        return OGL_OK;
    }
    const int minb = (int)ctx->ell_minb_cgp;
    const int grid = ell_grid(ctx, minb);
    const int chunk = (int)ctx->ell_chunk;
    cudaStream_t st = ctx->stream;
#define CGP_GO(W, P, G, B) k_spmv_ell_cgp<W, P, G, B><<<grid, kEllThreads, 0, st>>>(k, m, p_new, chunk)
#define CGP_Q(W, P, G)                             \
    do {                                           \
        if (minb == 2) CGP_GO(W, P, G, 2);         \
        else if (minb == 3) CGP_GO(W, P, G, 3);    \
        else CGP_GO(W, P, G, 4);                   \
    } while (0)
#define CGP_W(W)                                   \
    do {                                           \
        if (e.coded) {                             \
            if (ghost) CGP_Q(W, true, true);       \
            else CGP_Q(W, true, false);            \
        } else {                                   \
            if (ghost) CGP_Q(W, false, true);      \
            else CGP_Q(W, false, false);           \
        }                                          \
    } while (0)
    if (e.width == 7) CGP_W(7);
    else CGP_W(5);
#undef CGP_W
#undef CGP_Q
#undef CGP_GO
    ctx->launches++;
    OGL_CUDA(ctx, cudaGetLastError());
    return OGL_OK;
}

}  // namespace ogl
