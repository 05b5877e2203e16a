"""-m gpu: CUDA windowed realign (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("matrix", ["onepass", "ancient", "pe", "flat"])
def test_realign_matches_oracle(gpu, oracle, matrix):
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(3000, 4000, seed=11)
    bad, out = gpu_checks.check_realign(gpu, oracle, ref, bases, off, rc, as_, ae, gpu_checks.load_pssm(matrix))
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    assert (out["status"] == 0).all()


def test_realign_long_reads_and_indels(gpu, oracle):
    # merged-PE-like lengths (30-140) + long reads up to 256, heavy indels -> wider buckets
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(1500, 5000, seed=21, divergence=0.05, indel_rate=0.02,
                                                        min_len=30, max_len=256)
    bad, out = gpu_checks.check_realign(gpu, oracle, ref, bases, off, rc, as_, ae, gpu_checks.load_pssm("pe"))
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_realign_linear_reference_and_edges(gpu, oracle):
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(1000, 600, seed=31, min_len=20, max_len=60)
    bad, _ = gpu_checks.check_realign(gpu, oracle, ref, bases, off, rc, as_, ae, gpu_checks.load_pssm("ancient"), circular=0)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_realign_tiny_reads(gpu, oracle):
    # reads of 1..12 bases (tf13-really-short is 3 bp), windows ~100 columns
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(500, 800, seed=41, min_len=1, max_len=12)
    bad, _ = gpu_checks.check_realign(gpu, oracle, ref, bases, off, rc, as_, ae, gpu_checks.load_pssm("onepass"))
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_empty_batch(gpu):
    gpu.set_pssm(gpu_checks.load_pssm("flat"))
    gpu.set_reference("ACGT" * 50, circular=1)
    out = gpu.realign_host(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.uint8), np.zeros(0, np.int32),
                           np.zeros(0, np.int32))
    assert len(out["score"]) == 0


def test_packed_runs_equal_strided_runs(gpu, oracle):
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(2000, 3000, seed=91, divergence=0.04, indel_rate=0.02)
    gpu.set_pssm(gpu_checks.load_pssm("onepass"))
    gpu.set_reference(ref, circular=1)
    out = gpu.realign_host(bases, off, rc, as_, ae)
    n = len(off) - 1
    total, _, _ = gpu.get_runs_packed()
    assert total == int(np.maximum(out["n_runs"], 0).sum())
    run_off, packed = np.zeros(n + 1, np.int64), np.zeros(total, np.uint16)
    gpu.get_runs_packed(run_off, packed)
    assert (np.diff(run_off) == out["n_runs"]).all()
    for i in range(0, n, 7):
        assert (packed[run_off[i]:run_off[i + 1]] == out["runs"][i, :out["n_runs"][i]]).all()
