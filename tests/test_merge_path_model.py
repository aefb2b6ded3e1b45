"""CPU: the slice decomposition of the merge-path SpMV (ogl_b200/csrc/spmv_merge.cu) restated in numpy with
the kernel's own index logic -- chunk_row by lower bound, rows owned by the slice they start in, the head
of a slice carried to the row it continues, carries added in slice order -- and checked against a plain
CSR product on matrices chosen to hit every boundary case: empty rows (leading, trailing, at slice
boundaries), rows spanning many slices, rows ending exactly on a boundary, a last ragged slice, a
number of entries that is a multiple of the slice length.  (The CUDA kernel itself is covered by
tests/test_gpu_spmv.py and tools/pending_gpu_tests.)"""
import numpy as np
import pytest


def chunk_rows(rp, n, nnz, tile):
    n_chunks = (nnz + tile - 1) // tile
    out = np.empty(n_chunks + 1, np.int64)
    for c in range(n_chunks):
        lo, hi = 0, n                       # k_mp_chunk_rows: the answer is in [0, n]
        target = c * tile
        while lo < hi:
            mid = lo + (hi - lo) // 2
            if rp[mid] >= target:
                hi = mid
            else:
                lo = mid + 1
        out[c] = lo
    out[n_chunks] = n
    return out


def merge_path_spmv(rp, cols, vals, x, tile, alpha=1.0, beta=0.0, y_in=None):
    n, nnz = len(rp) - 1, int(rp[-1])
    y = np.full(n, np.nan)                 # every row must be written by exactly one owner
    writes = np.zeros(n, np.int64)
    n_chunks = (nnz + tile - 1) // tile
    cr = chunk_rows(rp, n, nnz, tile)
    carry_row, carry_val = np.full(n_chunks, -1, np.int64), np.zeros(n_chunks)
    for c in range(n_chunks):
        c0, c1 = c * tile, min((c + 1) * tile, nnz)
        prod = alpha * vals[c0:c1] * x[cols[c0:c1]]
        rfo, reo = cr[c], cr[c + 1]
        first_start = rp[rfo]
        for r in range(rfo, reo):
            s, e = rp[r], min(rp[r + 1], c1)
            y[r] = (beta * y_in[r] if y_in is not None else 0.0) + prod[s - c0:max(e, s) - c0].sum()
            writes[r] += 1
        if first_start > c0:
            carry_row[c] = rfo - 1
            carry_val[c] = prod[:min(first_start, c1) - c0].sum()
    for c in range(n_chunks):               # k_spmv_merge_fixup: one thread per run of equal rows
        r = carry_row[c]
        if r < 0 or (c > 0 and carry_row[c - 1] == r):
            continue
        acc, cc = y[r], c
        while cc < n_chunks and carry_row[cc] == r:
            acc += carry_val[cc]
            cc += 1
        y[r] = acc
    return y, writes


def random_csr(rng, n, lengths):
    rp = np.zeros(n + 1, np.int64)
    rp[1:] = np.cumsum(lengths)
    nnz = int(rp[-1])
    cols = rng.integers(0, n, nnz)
    vals = rng.normal(size=nnz)
    return rp, cols, vals


CASES = {
    "short rows": lambda rng: rng.integers(1, 6, 200),
    "empty rows everywhere": lambda rng: rng.integers(0, 3, 300),
    "leading and trailing empty rows": lambda rng: np.concatenate([np.zeros(5, int), rng.integers(1, 9, 100), np.zeros(7, int)]),
    "one row over many slices": lambda rng: np.concatenate([rng.integers(1, 4, 20), [257], rng.integers(1, 4, 20)]),
    "several long rows back to back": lambda rng: np.array([1, 100, 64, 33, 2, 0, 0, 90, 1]),
    "rows ending on slice boundaries": lambda rng: np.full(40, 8),
    "multiple of the slice length": lambda rng: np.array([16, 16, 32, 8, 8, 16]),
    "a single row": lambda rng: np.array([100]),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tile", [8, 16])
def test_decomposition_equals_the_csr_product(name, tile):
    rng = np.random.default_rng(100 * list(CASES).index(name) + tile)
    lengths = CASES[name](rng)
    n = len(lengths)
    rp, cols, vals = random_csr(rng, n, lengths)
    if rp[-1] == 0:
        pytest.skip("no entries")
    x, y_in = rng.normal(size=n), rng.normal(size=n)
    ref = np.array([(vals[rp[i]:rp[i + 1]] * x[cols[rp[i]:rp[i + 1]]]).sum() for i in range(n)])
    y, writes = merge_path_spmv(rp, cols, vals, x, tile)
    assert np.all(writes == 1), "every row is owned by exactly one slice"
    assert np.allclose(y, ref, rtol=1e-13, atol=1e-13)
    ya, _ = merge_path_spmv(rp, cols, vals, x, tile, alpha=-1.0, beta=1.0, y_in=y_in)
    assert np.allclose(ya, y_in - ref, rtol=1e-13, atol=1e-13)
