// TEST INFRASTRUCTURE -- parity oracle, see oracle.h.  Not product code.
//
// CPU restatement of the reference's LDU -> row-major COO assembly:
//   /root/reference/HostMatrix/HostMatrixFreeFunctions.C:21-201
//   /root/reference/HostMatrix/HostMatrix.C:180-207, 251-306, 412-466, 504-586,
//                                           634-704, 708-732
// Pinned by tests/test_oracle_golden.py against unitTests/test_HostMatrix.C and
// against oracle/_ref (the reference free functions compiled from source).
#include "oracle.h"

#include <algorithm>
#include <map>
#include <numeric>
#include <vector>

namespace {

struct Entry {
    orc_label row, col, face;
};

// Face triples ordered by (row, col) -- HostMatrixFreeFunctions.C:128-148 sorts
// with std::sort on that key; (row, col) pairs are unique for a valid mesh, so
// a stable sort yields the same sequence and fixes the order for degenerate
// input (duplicate faces).
std::vector<Entry> sorted_triangle(orc_label nf, const orc_label *row_of,
                                   const orc_label *col_of)
{
    std::vector<Entry> t(static_cast<size_t>(nf));
    for (orc_label f = 0; f < nf; ++f) t[f] = {row_of[f], col_of[f], f};
    std::stable_sort(t.begin(), t.end(), [](const Entry &a, const Entry &b) {
        return a.row != b.row ? a.row < b.row : a.col < b.col;
    });
    return t;
}

}  // namespace

extern "C" {

void orc_init_local_sparsity(orc_label nrows, orc_label upper_nnz,
                             int is_symmetric, const orc_label *upper,
                             const orc_label *lower, orc_label *rows,
                             orc_label *cols, orc_label *permute)
{
    // HostMatrixFreeFunctions.C:117: first staging slot after the off-diagonals
    const orc_label diag_base = is_symmetric ? upper_nnz : 2 * upper_nnz;
    // upper triangle: row = lowerAddr, col = upperAddr (:121-126)
    const auto up = sorted_triangle(upper_nnz, lower, upper);
    // lower triangle: row = upperAddr, col = lowerAddr (:137-142)
    const auto lo = sorted_triangle(upper_nnz, upper, lower);

    size_t iu = 0, il = 0, out = 0;
    for (orc_label r = 0; r < nrows; ++r) {
        // :162-178 lower entries of this row, slot f (sym) or F+f (asym)
        while (il < lo.size() && lo[il].row == r) {
            rows[out] = r;
            cols[out] = lo[il].col;
            permute[out] = is_symmetric ? lo[il].face : upper_nnz + lo[il].face;
            ++out, ++il;
        }
        // :180-184 diagonal
        rows[out] = r;
        cols[out] = r;
        permute[out] = diag_base + r;
        ++out;
        // :186-199 upper entries of this row, slot f
        while (iu < up.size() && up[iu].row == r) {
            rows[out] = r;
            cols[out] = up[iu].col;
            permute[out] = up[iu].face;
            ++out, ++iu;
        }
    }
}

void orc_merge_local_interfaces(orc_label nrows, orc_label upper_nnz,
                                int is_symmetric, orc_label n_iface,
                                const orc_label *iface_rows,
                                const orc_label *iface_cols, orc_label *rows,
                                orc_label *cols, orc_label *permute)
{
    if (n_iface <= 0) return;  // HostMatrix.C:506
    const orc_label local_nnz = nrows + 2 * upper_nnz;
    const orc_label diag_base = is_symmetric ? upper_nnz : 2 * upper_nnz;
    // :510-515 interface couplings in (row, col) order, original index kept
    std::vector<Entry> ifc(static_cast<size_t>(n_iface));
    for (orc_label i = 0; i < n_iface; ++i)
        ifc[i] = {iface_rows[i], iface_cols[i], i};
    std::stable_sort(ifc.begin(), ifc.end(), [](const Entry &a, const Entry &b) {
        return a.row != b.row ? a.row < b.row : a.col < b.col;
    });
    // :519-528 snapshot of the pattern without interfaces
    std::vector<orc_label> r0(rows, rows + local_nnz), c0(cols, cols + local_nnz),
        p0(permute, permute + local_nnz);
    orc_label src = 0, dst = 0;
    for (const Entry &e : ifc) {
        // :543-570 copy every existing entry that is not after (row, col);
        // an existing entry with the same (row, col) stays in front
        while (src < local_nnz &&
               (r0[src] < e.row || (r0[src] == e.row && c0[src] <= e.col))) {
            rows[dst] = r0[src], cols[dst] = c0[src], permute[dst] = p0[src];
            ++src, ++dst;
        }
        // :571-575
        rows[dst] = e.row;
        cols[dst] = e.col;
        permute[dst] = diag_base + nrows + e.face;
        ++dst;
    }
    // :580-585 tail
    for (; src < local_nnz; ++src, ++dst) {
        rows[dst] = r0[src], cols[dst] = c0[src], permute[dst] = p0[src];
    }
}

void orc_symmetric_update(orc_label total_nnz, orc_label upper_nnz,
                          const orc_label *permute, orc_scalar scale,
                          const orc_scalar *diag, const orc_scalar *upper,
                          orc_scalar *out)
{
    // documented intent of HostMatrixFreeFunctions.C:21-30 (README.md:81)
    for (orc_label i = 0; i < total_nnz; ++i) {
        const orc_label pos = permute[i];
        out[i] = scale * (pos >= upper_nnz ? diag[pos - upper_nnz] : upper[pos]);
    }
}

void orc_symmetric_update_as_written(orc_label total_nnz, orc_label upper_nnz,
                                     const orc_label *permute, orc_scalar scale,
                                     const orc_scalar *diag,
                                     const orc_scalar *upper, orc_scalar *out)
{
    // HostMatrixFreeFunctions.C:27-28 parses as (scale*(pos>=F)) ? diag : upper:
    // the value is never scaled and scale == 0 selects `upper` for every slot.
    for (orc_label i = 0; i < total_nnz; ++i) {
        const orc_label pos = permute[i];
        const bool pick_diag = (scale * (pos >= upper_nnz ? 1.0 : 0.0)) != 0.0;
        out[i] = pick_diag ? diag[pos - upper_nnz] : upper[pos];
    }
}

void orc_non_symmetric_update(orc_label total_nnz, orc_label upper_nnz,
                              const orc_label *permute, orc_scalar scale,
                              const orc_scalar *diag, const orc_scalar *upper,
                              const orc_scalar *lower, orc_scalar *out)
{
    // HostMatrixFreeFunctions.C:84-102
    for (orc_label i = 0; i < total_nnz; ++i) {
        const orc_label pos = permute[i];
        if (pos < upper_nnz)
            out[i] = scale * upper[pos];
        else if (pos < 2 * upper_nnz)
            out[i] = scale * lower[pos - upper_nnz];
        else
            out[i] = scale * diag[pos - 2 * upper_nnz];
    }
}

void orc_symmetric_update_w_interface(orc_label total_nnz, orc_label diag_nnz,
                                      orc_label upper_nnz,
                                      const orc_label *permute, orc_scalar scale,
                                      const orc_scalar *diag,
                                      const orc_scalar *upper,
                                      const orc_scalar *iface, orc_scalar *out)
{
    // HostMatrixFreeFunctions.C:32-55
    for (orc_label i = 0; i < total_nnz; ++i) {
        const orc_label pos = permute[i];
        orc_scalar v;
        if (pos < upper_nnz)
            v = upper[pos];
        else if (pos < upper_nnz + diag_nnz)
            v = diag[pos - upper_nnz];
        else
            v = iface[pos - upper_nnz - diag_nnz];
        out[i] = scale * v;
    }
}

void orc_non_symmetric_update_w_interface(
    orc_label total_nnz, orc_label diag_nnz, orc_label upper_nnz,
    const orc_label *permute, orc_scalar scale, const orc_scalar *diag,
    const orc_scalar *upper, const orc_scalar *lower, const orc_scalar *iface,
    orc_scalar *out)
{
    // HostMatrixFreeFunctions.C:57-81
    for (orc_label i = 0; i < total_nnz; ++i) {
        const orc_label pos = permute[i];
        orc_scalar v;
        if (pos < upper_nnz)
            v = upper[pos];
        else if (pos < 2 * upper_nnz)
            v = lower[pos - upper_nnz];
        else if (pos < 2 * upper_nnz + diag_nnz)
            v = diag[pos - 2 * upper_nnz];
        else
            v = iface[pos - 2 * upper_nnz - diag_nnz];
        out[i] = scale * v;
    }
}

void orc_gather_from_staging(orc_label total_nnz, const orc_label *permute,
                             const orc_scalar *staging, orc_scalar *out)
{
    // HostMatrix.C:685-703: Dense::row_gather(ldu_mapping) then copy_from
    for (orc_label i = 0; i < total_nnz; ++i) out[i] = staging[permute[i]];
}

void orc_negate(orc_label n, const orc_scalar *in, orc_scalar *out)
{
    // HostMatrix.C:204
    for (orc_label i = 0; i < n; ++i) out[i] = in[i] * -1.0;
}

void orc_comm_pattern(orc_label n_proc_ifaces, const orc_label *nbr,
                      const orc_label *sz, const orc_label *face_cells,
                      orc_label *n_targets, orc_label *target_ids,
                      orc_label *target_sizes, orc_label *send_idxs)
{
    // HostMatrix.C:257-278: std::map keyed by neighbour rank, faceCells appended
    // in interface order
    std::map<orc_label, std::vector<orc_label>> by_rank;
    orc_label off = 0;
    for (orc_label i = 0; i < n_proc_ifaces; ++i) {
        auto &v = by_rank[nbr[i]];
        v.insert(v.end(), face_cells + off, face_cells + off + sz[i]);
        off += sz[i];
    }
    // :284-303 ascending rank order
    orc_label t = 0, w = 0;
    for (const auto &kv : by_rank) {
        target_ids[t] = kv.first;
        target_sizes[t] = static_cast<orc_label>(kv.second.size());
        for (orc_label c : kv.second) send_idxs[w++] = c;
        ++t;
    }
    *n_targets = t;
}

void orc_non_local_pattern(orc_label n_halo, const orc_label *face_cells,
                           orc_label *rows, orc_label *cols, orc_label *permute)
{
    // HostMatrix.C:420-431: (running face counter, faceCell); :452-457 sort by
    // row only; :459-465 cols = permute = running counter
    std::vector<orc_label> order(static_cast<size_t>(n_halo));
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](orc_label a, orc_label b) {
                         return face_cells[a] < face_cells[b];
                     });
    for (orc_label k = 0; k < n_halo; ++k) {
        rows[k] = face_cells[order[k]];
        cols[k] = order[k];
        permute[k] = order[k];
    }
}

void orc_non_local_update(orc_label n_halo, const orc_label *permute,
                          const orc_scalar *neg_coeffs, orc_scalar *out)
{
    // HostMatrix.C:723-726
    for (orc_label k = 0; k < n_halo; ++k) out[k] = neg_coeffs[permute[k]];
}

}  // extern "C"
