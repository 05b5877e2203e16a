"""-m gpu: the repeat filter (miagpu_repeat_filter = sort_fsdb / sort_fsdb_qscore + set_uniq_in_fsdb) against the
oracle, which tests/test_oracle_vs_ref.py and tests/golden/repeat_cases.json pin to the reference's own functions.
Bit-exact: the permutation (ties keep FSDB order) and every unique_best flag."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(rng, n, span, ties):
    rc = rng.integers(0, 2, n).astype(np.uint8)
    as_ = rng.integers(0, span, n).astype(np.int32)
    ae = (as_ + rng.integers(30, 30 + ties, n)).astype(np.int32)
    key4 = rng.integers(2000, 2000 + ties, n).astype(np.int32)
    tr = rng.integers(0, 2, n).astype(np.uint8)
    return rc, as_, ae, key4, tr


@pytest.mark.parametrize("just_outer,tolerance", [(1, 0), (0, 0), (1, 2), (0, 3)])
def test_repeat_filter_matches_oracle(gpu, oracle, just_outer, tolerance):
    rng = np.random.default_rng(10 * just_outer + tolerance)
    for n, span, ties in ((1, 10, 2), (2, 3, 2), (37, 5, 3), (5000, 40, 4), (60000, 16600, 50), (200000, 300, 6)):
        rc, as_, ae, key4, tr = _case(rng, n, span, ties)
        o_order, o_uniq = oracle.repeat_filter(rc, as_, ae, key4, tr, just_outer, tolerance)
        g_order, g_uniq = gpu.repeat_filter(rc, as_, ae, key4, tr, just_outer, tolerance)
        assert (g_order == o_order).all(), (n, int((g_order != o_order).sum()))
        assert (g_uniq == o_uniq).all(), (n, int((g_uniq != o_uniq).sum()))
    # without the trimmed flags and without the permutation
    rc, as_, ae, key4, _ = _case(rng, 3000, 20, 3)
    _, o_uniq = oracle.repeat_filter(rc, as_, ae, key4, None, just_outer, tolerance)
    g_order, g_uniq = gpu.repeat_filter(rc, as_, ae, key4, None, just_outer, tolerance, want_order=False)
    assert g_order is None and (g_uniq == o_uniq).all()


def test_repeat_filter_at_scale_and_limits(gpu, oracle):
    from mia_b200 import api
    rng = np.random.default_rng(99)
    n = 3_000_000                                            # realistic: 16.5 kb circular genome, 180x coverage, real duplicates
    rc = rng.integers(0, 2, n).astype(np.uint8)
    as_ = rng.integers(0, 16569, n).astype(np.int32)
    ae = (as_ + rng.integers(34, 75, n)).astype(np.int32)
    score = (200 * (ae - as_ + 1) - rng.integers(0, 900, n)).astype(np.int32)
    g_order, g_uniq = gpu.repeat_filter(rc, as_, ae, score)
    o_order, o_uniq = oracle.repeat_filter(rc, as_, ae, score)
    assert (g_order == o_order).all() and (g_uniq == o_uniq).all()
    assert 0.2 < g_uniq.mean() < 0.9                         # the case really has duplicates
    # negative scores and the largest coordinates the 64-bit key holds
    rc, as_, ae = np.array([0, 0, 1, 1], np.uint8), np.array([0, 0, 2097000, 5], np.int32), np.array([2097151, 9, 2097151, 2097151], np.int32)
    sc = np.array([-1048576, 1048575, -5, 0], np.int32)
    assert (gpu.repeat_filter(rc, as_, ae, sc)[0] == oracle.repeat_filter(rc, as_, ae, sc)[0]).all()
    with pytest.raises(api.MiaGpuError, match="coordinates"):
        gpu.repeat_filter(rc, as_, np.array([2097152, 9, 3, 4], np.int32), sc)
