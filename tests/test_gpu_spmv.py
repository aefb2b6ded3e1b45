"""-m gpu: FP64 CSR SpMV kernels against the oracle (bit-exact where the
summation order is the reference's) and size-independent properties at the
BASELINE size."""
import numpy as np
import pytest

from gpu_helpers import upload_system
from ogl_b200 import cases
from ogl_b200.backend import Context

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context()
    yield c
    c.close()


@pytest.mark.parametrize("builder", [lambda: cases.pressure_3d(20)[0], lambda: cases.momentum_3d(17)[0],
                                     lambda: cases.channel((16, 8, 8), (1, 1, 1))[0],
                                     lambda: cases.cavity_2d((1, 1, 1))[0]])
@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7])
def test_spmv_vs_oracle(ctx, oracle, builder, variant):
    s = builder()
    upload_system(ctx, s, partition=False)
    ctx.set_option("spmv_variant", variant)
    a = oracle.assemble(s)
    x = np.random.default_rng(3).normal(size=s.n)
    y = ctx.spmv(x)
    y_ref = oracle.dist_spmv([a], [x])[0]
    if variant in (1, 2, 4, 5, 6, 7):
        # same left-to-right row sums, products rounded before the add: bit-exact
        assert np.array_equal(y, y_ref)
    else:
        assert np.allclose(y, y_ref, rtol=1e-13, atol=1e-13 * np.abs(y_ref).max())
    ctx.set_option("spmv_variant", 0)


def test_spmv_one_long_row_takes_the_merge_path_kernel(ctx, oracle):
    # an "arrow" matrix: row 0 couples to everything (row length n) -> entry-balanced slices (variant 8)
    n = 3000
    lower = np.zeros(n - 1, np.int32)
    upper = np.arange(1, n, dtype=np.int32)
    rng = np.random.default_rng(0)
    diag, up = rng.uniform(1, 2, n), rng.normal(size=n - 1)
    ctx.pattern_from_ldu(n, lower, upper, True)
    ctx.values_update(diag, up)
    x = rng.normal(size=n)
    y = ctx.spmv(x)
    ref = diag * x
    ref[0] += up @ x[1:]
    ref[1:] += up * x[0]
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12)
    assert ctx.get_option("max_row_len") == n
    assert ctx.get_option("spmv_variant_in_use") == 8


def test_spmv_properties_at_baseline_size(ctx):
    # 100^3 pressure system (BASELINE configs[1]): linearity and the manufactured rhs
    s = cases.pressure_3d(100)[0]
    upload_system(ctx, s, partition=False)
    assert ctx.nnz == 100 ** 3 + 2 * 3 * 100 * 100 * 99
    rng = np.random.default_rng(11)
    u, v = rng.normal(size=s.n), rng.normal(size=s.n)
    yu, yv, yuv = ctx.spmv(u), ctx.spmv(v), ctx.spmv(u + 2.0 * v)
    scale = np.abs(yu).max() + np.abs(yv).max()
    assert np.abs(yuv - (yu + 2.0 * yv)).max() <= 1e-14 * scale
    # b was built as A x*: A x* must reproduce it to rounding
    assert np.abs(ctx.spmv(s.x_star) - s.source).max() <= 1e-14 * np.abs(s.source).max() + 1e-20
    # symmetric operator: <u, A v> == <v, A u>
    assert abs(u @ yv - v @ yu) <= 1e-12 * abs(u @ yv)
    # all three kernels agree
    ys = []
    for variant in (1, 2, 3, 4, 5, 6):
        ctx.set_option("spmv_variant", variant)
        ys.append(ctx.spmv(u))
    ctx.set_option("spmv_variant", 0)
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[3])
    assert np.array_equal(ys[0], ys[4]) and np.array_equal(ys[0], ys[5])
    assert np.abs(ys[2] - ys[0]).max() <= 1e-14 * scale


def test_l2_policy_of_the_matrix_stream_is_bit_exact(ctx, oracle):
    """l2_keep_mb only changes the cache hints of the (column, value) loads."""
    s = cases.pressure_3d(24)[0]
    upload_system(ctx, s, partition=False)
    x = np.random.default_rng(4).normal(size=s.n)
    ys = []
    for keep in (0, -1, 1):
        ctx.set_option("l2_keep_mb", keep)
        ys.append(ctx.spmv(x))
    ctx.set_option("l2_keep_mb", 0)
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])
    a = oracle.assemble(s)
    assert np.array_equal(ys[0], oracle.dist_spmv([a], [x])[0])


def test_spmv_bench_entry_point(ctx):
    s = cases.pressure_3d(32)[0]
    upload_system(ctx, s, partition=False)
    assert ctx.spmv_bench(5) > 0
    assert ctx.spmv_bench(5, fused_dot=True) > 0


@pytest.mark.parametrize("builder,patterns", [(lambda: cases.pressure_3d(20)[0], 27),
                                              (lambda: cases.momentum_3d(17)[0], 27),
                                              (lambda: cases.cavity_2d((1, 1, 1))[0], 9),
                                              (lambda: cases.channel((16, 8, 8), (1, 1, 1))[0], None)])
def test_ell_pattern_coded_columns(ctx, oracle, builder, patterns):
    """ELL with 1-byte row-pattern codes instead of 4-byte columns: same bits as plain ELL and as
    the oracle; a structured box has 27 (3-D) / 9 (2-D) distinct (column - row) tuples."""
    s = builder()
    x = np.random.default_rng(5).normal(size=s.n)
    ys = {}
    for coded in (0, 1):
        upload_system(ctx, s, partition=False)
        ctx.set_option("spmv_variant", 7)
        ctx.set_option("ell_coded", coded)
        ys[coded] = ctx.spmv(x)
        assert (ctx.get_option("ell_coded_active") & 1) == coded
        if coded:
            assert ctx.get_option("ell_escape_rows") == 0
            if patterns is not None:
                assert ctx.get_option("ell_patterns") == patterns
    assert np.array_equal(ys[0], ys[1])
    assert np.array_equal(ys[1], oracle.dist_spmv([oracle.assemble(s)], [x])[0])
    ctx.set_option("spmv_variant", 0)
    ctx.set_option("ell_coded", 1)


def test_ell_codes_fall_back_on_an_unstructured_pattern(ctx, oracle):
    """> 255 distinct tuples: the rows beyond the table escape to their 4-byte columns; with most
    rows escaping the format stays plain ELL.  Either way the result is the oracle's."""
    from conftest import random_ldu_mesh
    rng = np.random.default_rng(11)
    n = 2500
    lower, upper = random_ldu_mesh(rng, n, 1200)
    ctx.pattern_from_ldu(n, lower, upper, True)
    if ctx.get_option("max_row_len") > 8:
        pytest.skip("random mesh drew a row longer than the tuple table holds")
    diag, up = rng.uniform(4, 5, n), rng.normal(size=lower.size)
    x = rng.normal(size=n)
    ref = diag * x
    np.add.at(ref, lower, up * x[upper])
    np.add.at(ref, upper, up * x[lower])
    for coded in (1, 2):
        ctx.pattern_from_ldu(n, lower, upper, True)
        ctx.values_update(diag, up)
        ctx.set_option("spmv_variant", 7)
        ctx.set_option("ell_coded", coded)
        y = ctx.spmv(x)
        assert np.allclose(y, ref, rtol=1e-12, atol=1e-12)
        assert ctx.get_option("ell_escape_rows") > n // 4
        assert (ctx.get_option("ell_coded_active") & 1) == (1 if coded == 2 else 0)
    ctx.set_option("spmv_variant", 0)
    ctx.set_option("ell_coded", 1)


def test_row_length_histogram_drives_the_kernel_choice(ctx):
    """spmv_setup builds a power-of-two row-length histogram on the device; the automatic kernel
    choice reads it: stencil rows -> ELL / pipelined stream, an arrow matrix -> warp per row."""
    s = cases.pressure_3d(12)[0]
    upload_system(ctx, s, partition=False)
    ctx.set_option("spmv_variant", 0)
    hist = [ctx.get_option(f"row_len_hist_{b}") for b in range(8)]
    assert sum(hist) == s.n and hist[2] == 8 and hist[3] == s.n - 8   # corners: 4 entries, the rest 5..7
    assert ctx.get_option("spmv_variant_in_use") in (6, 7)
    n = 4000                                            # arrow: row 0 holds a third of the entries
    lower = np.zeros(n - 1, np.int32)
    upper = np.arange(1, n, dtype=np.int32)
    ctx.pattern_from_ldu(n, lower, upper, True)
    hist = [ctx.get_option(f"row_len_hist_{b}") for b in range(8)]
    assert hist[7] == 1 and hist[1] == n - 1
    assert ctx.get_option("spmv_variant_in_use") in (2, 3)


@pytest.mark.parametrize("cells", [4, 5, 7, 33])
def test_tma_fed_coded_ell_is_bit_identical(ctx, oracle, cells):
    """ell_tma: the value stream goes through cp.async.bulk + mbarrier stages instead of per-thread
    loads; same row sums.  64 / 125 / 343 rows: less than one or two tiles; 35937: a ragged last tile."""
    s = cases.pressure_3d(cells)[0]
    x = np.random.default_rng(8).normal(size=s.n)
    upload_system(ctx, s, partition=False)
    ctx.set_option("spmv_variant", 7)
    ref = oracle.dist_spmv([oracle.assemble(s)], [x])[0]
    for stages in (2, 3, 4):
        ctx.set_option("ell_tma", 1)
        ctx.set_option("tma_stages", stages)
        assert np.array_equal(ctx.spmv(x), ref)
    ms = ctx.spmv_bench(5, fused_dot=True)
    assert ms > 0
    ctx.set_option("ell_tma", 2)
    ctx.set_option("tma_stages", 3)
    ctx.set_option("spmv_variant", 0)
