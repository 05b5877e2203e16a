"""-m gpu: column accumulation + base calling (C ABI) vs the CPU oracle, bit-exact."""
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cons_code,tiles", [(1, "1"), (2, "1"), (1, "0")])
def test_consensus_matches_oracle(gpu, oracle, cons_code, tiles, monkeypatch):
    # tiles: "1" = tile-private shared-memory accumulators (tile_kernel), "0" = global REDs (entry_kernel<1>)
    monkeypatch.setenv("MIAGPU_CONS_TILES", tiles)
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(3000, 4100, seed=51, divergence=0.03, indel_rate=0.01)
    problems, info = gpu_checks.check_consensus(gpu, oracle, ref, bases, off, rc, as_, ae, gpu_checks.load_pssm("onepass"),
                                                cons_code=cons_code)
    assert not problems, problems
    assert info["n_split"] > 0 and info["n_ins_cols"] > 0      # the case must exercise wrap splits and insert columns


def test_consensus_low_coverage_linear(gpu, oracle):
    # coverage ~1x: N calls, ties, uncovered columns; linear reference: no splits
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(60, 3000, seed=61, divergence=0.05, indel_rate=0.01)
    problems, _ = gpu_checks.check_consensus(gpu, oracle, ref, bases, off, rc, as_, ae, gpu_checks.load_pssm("ancient"),
                                             circular=0, drop_frac=0.3)
    assert not problems, problems


def test_consensus_that_does_not_fit_the_callers_buffer_fails(gpu):
    # the reference sizes its consensus string from the gaps it has counted (mia.c:527-533); a caller of the library
    # announces its buffer (miagpu_set_cons_capacity) and a longer consensus fails the call instead of overrunning it
    from mia_b200 import api
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(300, 1200, seed=71)
    gpu.set_pssm(gpu_checks.load_pssm("onepass"))
    gpu.set_reference(ref, circular=1, with_rc=0)
    gpu.realign_host(bases, off, rc, as_, ae)
    cons, _, _ = gpu.consensus_natural()
    assert len(cons) > 1000
    assert gpu.lib.miagpu_set_cons_capacity(gpu.h, len(cons))          # one byte short of the string + its terminator
    try:
        with pytest.raises(api.MiaGpuError, match="cons_out holds"):
            gpu.consensus_natural()
        assert gpu.lib.miagpu_set_cons_capacity(gpu.h, len(cons) + 1)
        assert gpu.consensus_natural()[0] == cons
    finally:
        assert gpu.lib.miagpu_set_cons_capacity(gpu.h, 0)
