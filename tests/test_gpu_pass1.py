"""-m gpu: pass 1 (k-mer filter + whole-reference both-strand DP, chunked kernel) vs the oracle."""
import numpy as np
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


def _reads(n, ref_len, seed, **kw):
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(ref_len, seed=seed)
    g = synth.diverge(ref, kw.pop("divergence", 0.02), seed=seed + 1, indel_rate=kw.pop("indel_rate", 0.004))
    b, off, _ = synth.make_reads(g, n, kw.pop("min_len", 35), kw.pop("max_len", 75), seed=seed + 2, n_rate=0.002)
    reads = [synth.read_str(b, off, i) for i in range(n)]
    rng = np.random.default_rng(seed)
    for i in range(0, n, 17):       # unrelated reads: must be skipped by the filter / score below the cutoff
        reads[i] = "".join("ACGT"[x] for x in rng.integers(0, 4, 50))
    return ref, reads


@pytest.mark.parametrize("k", [0, 8, 12])
def test_pass1_circular(gpu, oracle, k):
    ref, reads = _reads(600, 3000, seed=101)
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("onepass"), 1, k)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    if k == 12:
        assert (out["status"] & 2).any()       # unrelated reads share no 12-mer with a 3 kb reference


def test_pass1_linear_low_complexity_softmask(gpu, oracle):
    # poly-AC stretch (k-mer position cap + saturation) and a lower-case region with -M
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(1200, seed=5) + "AC" * 300 + synth.random_reference(300, seed=6).lower() + synth.random_reference(600, seed=7)
    rng = np.random.default_rng(9)
    reads = []
    for _ in range(300):
        p = int(rng.integers(0, len(ref) - 80))
        reads.append(ref[p:p + int(rng.integers(25, 75))].upper())
    bad, _ = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("ancient"), 0, 9, soft_mask=1)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_pass1_long_reads(gpu, oracle):
    ref, reads = _reads(200, 2500, seed=111, min_len=100, max_len=256, divergence=0.04, indel_rate=0.01)
    bad, _ = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("pe"), 1, 10)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_pass1_fast_path_vs_oracle(gpu, oracle):
    # k = 12 on a 5 kb circular reference: nearly every read is one stretch on one strand -> windowed pair kernels;
    # reads near position 0 (real column 0) and in the wrap (stretch clipped at the last column) included
    ref, reads = _reads(3000, 5000, seed=211, divergence=0.01, indel_rate=0.002)
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("onepass"), 1, 12)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast > 2400 and fast + general + skipped == len(reads), (fast, general, skipped)


def test_pass1_fast_path_linear_short_k(gpu, oracle):
    # linear reference, k = 10: more spurious hits (second stretches, hits on both strands)
    ref, reads = _reads(1500, 4000, seed=223, divergence=0.03, indel_rate=0.006)
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("ancient"), 0, 10)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    fast, general, skipped = gpu.last_pass1_stats()
    route = gpu.last_pass1_route()
    # 3 % divergence + indels: gapped winners; their stretches are traced by the windowed 32-bit kernel (route 4), not the general one
    assert fast > 500 and int((route == 4).sum()) > 50, (fast, general, skipped, np.bincount(route).tolist())


def test_pass1_traced_winners_equal_general_kernel(gpu, monkeypatch):
    # size-independent property: gapped winners traced in their own stretch (realign_kernel<K, false, true>) come out exactly as the
    # general chunked kernel computes them over the whole masked strands
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(16569, seed=1)
    g = synth.diverge(ref, 0.08, seed=2, indel_rate=0.006)
    b, off, _ = synth.make_reads(g, 40000, 30, 140, seed=78)
    gpu.set_pssm(gpu_checks.load_pssm("ancient"))
    gpu.set_reference(ref, circular=1, with_rc=1)
    gpu.build_kmers(10)
    gpu.upload_reads(b, off)
    a = gpu.pass1()
    route = gpu.last_pass1_route()
    assert int((route == 4).sum()) > 4000, np.bincount(route).tolist()
    assert not (a["status"][route == 4] != 0).any()
    monkeypatch.setenv("MIAGPU_P1_NO_TRACE", "1")
    z = gpu.pass1()
    assert int((gpu.last_pass1_route() == 4).sum()) == 0
    for k in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[k] == z[k]).all(), (k, int((a[k] != z[k]).sum()), np.flatnonzero(a[k] != z[k])[:5].tolist())
    nr = np.maximum(a["n_runs"], 0)
    m = np.arange(a["runs"].shape[1])[None, :] < nr[:, None]
    assert (np.where(m, a["runs"], 0) == np.where(m, z["runs"], 0)).all()


def test_pass1_fast_equals_general_large(gpu, monkeypatch):
    # size-independent property at bench scale: the windowed fast path and the general chunked kernel agree read by read
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(16569, seed=1)
    g = synth.diverge(ref, 0.005, seed=2, indel_rate=0.001)
    b, off, _ = synth.make_reads(g, 60000, 35, 75, seed=77)
    gpu.set_pssm(gpu_checks.load_pssm("onepass"))
    gpu.set_reference(ref, circular=1, with_rc=1)
    gpu.build_kmers(12)
    gpu.upload_reads(b, off)
    a = gpu.pass1()
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast > 50000
    monkeypatch.setenv("MIAGPU_PASS1_FAST", "0")
    z = gpu.pass1()
    assert gpu.last_pass1_stats()[0] == 0
    for k in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[k] == z[k]).all(), (k, int((a[k] != z[k]).sum()))
    nr = np.maximum(a["n_runs"], 0)
    m = np.arange(a["runs"].shape[1])[None, :] < nr[:, None]
    assert (np.where(m, a["runs"], 0) == np.where(m, z["runs"], 0)).all()


def test_pass1_saturated_strands_take_the_16bit_sweep(gpu, monkeypatch):
    # reads with >= 128 k-mer hits on a strand (kmer.c:283-285: the whole strand is unmasked): whole-strand jobs of sweep16_kernel
    # in the re-based frame, two reads per half-warp, the other strand's stretches as windowed jobs -- equal to the general kernel
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(16569, seed=1)
    g = synth.diverge(ref, 0.004, seed=2, indel_rate=0.0005)
    b, off, _ = synth.make_reads(g, 30000, 120, 200, seed=79)
    gpu.set_pssm(gpu_checks.load_pssm("pe"))
    gpu.set_reference(ref, circular=1, with_rc=1)
    gpu.build_kmers(12)
    gpu.upload_reads(b, off)
    a = gpu.pass1()
    fast, general, skipped = gpu.last_pass1_stats()
    assert (a["hits"] >= 128).sum() > 10000 and general < 3000, (int((a["hits"] >= 128).sum()), fast, general, skipped)
    monkeypatch.setenv("MIAGPU_P1_NO_SWEEP", "1")
    z = gpu.pass1()
    assert gpu.last_pass1_stats()[1] > 10000
    for k in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[k] == z[k]).all(), (k, int((a[k] != z[k]).sum()), np.flatnonzero(a[k] != z[k])[:5].tolist())
    nr = np.maximum(a["n_runs"], 0)
    m = np.arange(a["runs"].shape[1])[None, :] < nr[:, None]
    assert (np.where(m, a["runs"], 0) == np.where(m, z["runs"], 0)).all()


@pytest.mark.parametrize("circular,k", [(0, 10), (1, 12)])
def test_pass1_multi_stretch_repeats_and_ties(gpu, oracle, circular, k):
    # exact and near-exact repeats far apart: several stretches per strand with equal / competing scores -> every
    # stretch is its own windowed job, the first maximum in column order wins (mia.c:1278-1302), ties between the
    # strands go to the reverse strand (palindromic repeat unit)
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    rng = np.random.default_rng(77)
    ref = list(synth.random_reference(9000, seed=70))
    unit = ref[500:700]
    ref[4000:4200] = unit                                   # exact copy
    near = list(unit)
    near[40], near[120] = ("A" if near[40] != "A" else "C"), ("G" if near[120] != "G" else "T")
    ref[6500:6700] = near                                   # two mismatches
    rcu = list(synth.revcomp_bytes(np.frombuffer("".join(unit).encode(), np.uint8)).tobytes().decode())
    ref[8000:8200] = rcu                                    # the unit's reverse complement: hits on both strands
    ref = "".join(ref)
    reads = []
    for _ in range(400):
        L = int(rng.integers(35, 76))
        p = int(rng.integers(500, 700 - L))
        r = ref[p:p + L]
        if rng.random() < 0.5:
            r = synth.revcomp_bytes(np.frombuffer(r.encode(), np.uint8)).tobytes().decode()
        if rng.random() < 0.3:                              # a substitution: breaks some ties
            q = int(rng.integers(0, L))
            r = r[:q] + "ACGT"[(("ACGT".index(r[q])) + 1) % 4] + r[q + 1:]
        reads.append(r)
    g = synth.diverge(ref, 0.01, seed=71, indel_rate=0.002)
    b, off, _ = synth.make_reads(g, 600, 35, 75, seed=72, circular=bool(circular))
    reads += [synth.read_str(b, off, i) for i in range(600)]
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("onepass"), circular, k)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast > 700, (fast, general, skipped)             # the repeat reads (3-4 stretches) stay on the fast path


@pytest.mark.parametrize("k", [0, 9])
def test_strip_team_kernel_equals_warp_kernel(gpu, monkeypatch, k):
    # the general kernel's two schedules (a warp per read / a team of 16 warps per read, chunks pipelined one row apart)
    # must agree field for field; low-complexity reads saturate the filter (whole strand unmasked)
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(5000, seed=31) + "AC" * 200 + synth.random_reference(3000, seed=32)
    g = synth.diverge(ref, 0.03, seed=33, indel_rate=0.006)
    b, off, _ = synth.make_reads(g, 500, 30, 140, seed=34)
    gpu.set_pssm(gpu_checks.load_pssm("ancient"))
    gpu.set_reference(ref, circular=1, with_rc=1)
    gpu.build_kmers(k)
    gpu.upload_reads(b, off)
    monkeypatch.setenv("MIAGPU_PASS1_FAST", "0")
    monkeypatch.setenv("MIAGPU_STRIP_TEAM", "0")
    a = gpu.pass1()
    monkeypatch.setenv("MIAGPU_STRIP_TEAM", "1")
    z = gpu.pass1()
    for key in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[key] == z[key]).all(), (key, int((a[key] != z[key]).sum()))
    nr = np.maximum(a["n_runs"], 0)
    m = np.arange(a["runs"].shape[1])[None, :] < nr[:, None]
    assert (np.where(m, a["runs"], 0) == np.where(m, z["runs"], 0)).all()


def test_unmasked_sweep16_equals_general_kernel(gpu, monkeypatch):
    # no k-mer filter: both whole strands in one 16-bit sweep (sweep16.cuh) vs the general chunked kernel, field for field;
    # circular wrap, indels (winner not a plain diagonal -> handed over), reads longer than the 16-bit frame (never swept)
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(16569, seed=1)
    g = synth.diverge(ref, 0.02, seed=41, indel_rate=0.004)
    b, off, _ = synth.make_reads(g, 6000, 30, 140, seed=42, n_rate=0.002)
    gpu.set_pssm(gpu_checks.load_pssm("onepass"))
    gpu.set_reference(ref, circular=1, with_rc=1)
    gpu.build_kmers(0)
    gpu.upload_reads(b, off)
    a = gpu.pass1()
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast > 3000 and general > 100 and fast + general == 6000, (fast, general, skipped)
    monkeypatch.setenv("MIAGPU_PASS1_FAST", "0")
    z = gpu.pass1()
    assert gpu.last_pass1_stats()[0] == 0
    for key in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[key] == z[key]).all(), (key, int((a[key] != z[key]).sum()), np.flatnonzero(a[key] != z[key])[:5])
    nr = np.maximum(a["n_runs"], 0)
    m = np.arange(a["runs"].shape[1])[None, :] < nr[:, None]
    assert (np.where(m, a["runs"], 0) == np.where(m, z["runs"], 0)).all()


@pytest.mark.parametrize("circular,matrix", [(0, "ancient"), (1, "flat")])
def test_unmasked_sweep16_edge_reads_vs_oracle(gpu, oracle, circular, matrix):
    # 1-bp to 12-bp reads, reads of N, reads hanging over either end of a linear reference, a reference whose width is
    # not a multiple of the 256-column chunk, ties between the strands (palindromes) -- all without the k-mer filter
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    rng = np.random.default_rng(5)
    ref = synth.random_reference(1337, seed=51)
    reads = []
    for L in list(range(1, 13)) + [20, 33, 64, 100, 121]:
        p = int(rng.integers(0, len(ref) - L))
        reads.append(ref[p:p + L])
        reads.append(synth.revcomp_bytes(np.frombuffer(ref[p:p + L].encode(), np.uint8)).tobytes().decode())
    reads += ["N" * 30, "ACGT" * 10, "AATT" * 8, ref[:40], ref[-40:], "GG" + ref[:30], ref[-30:] + "TTT", ref[600:640] + "NNN" + ref[643:670]]
    g = synth.diverge(ref, 0.05, seed=52, indel_rate=0.01)
    b, off, _ = synth.make_reads(g, 150, 25, 110, seed=53, circular=bool(circular), n_rate=0.01)
    reads += [synth.read_str(b, off, i) for i in range(150)]
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm(matrix), circular, 0)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast > 100, (fast, general, skipped)


def test_pass1_long_reads_take_the_windowed_fast_path(gpu, oracle, monkeypatch):
    # merged-pair lengths (BASELINE configs[2]: 30-140 bp, here up to 200): stretches of reads beyond the low 16-bit frame run in
    # the RB frame of the pair kernels (pair16.cuh 5.) instead of the general chunked kernel -- exact against the oracle, and
    # field for field against the general kernel
    ref, reads = _reads(2500, 6000, seed=411, min_len=90, max_len=200, divergence=0.02, indel_rate=0.002)
    sm = gpu_checks.load_pssm("pe")
    bad, a = gpu_checks.check_pass1(gpu, oracle, ref, reads, sm, 1, 12)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast > 1500, (fast, general, skipped)       # the rest: stretches wider than 256 columns (reads beyond ~230 bases), gapped winners
    monkeypatch.setenv("MIAGPU_PAIR_RB", "0")
    z = gpu.pass1()
    assert gpu.last_pass1_stats()[0] < fast
    for k in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[k] == z[k]).all(), (k, int((a[k] != z[k]).sum()))
