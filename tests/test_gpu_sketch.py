"""K1 (CUDA minimizer sketch) through the C-ABI against the reference's mm_sketch: same elements, same order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rand_seq(rng, n, n_frac=0.0, lower=False, repeat=None):
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
    if repeat:
        unit = s[:repeat]
        for st in range(0, n - repeat, repeat * 3):
            s[st:st + repeat] = unit
    if n_frac > 0:
        m = rng.random(n) < n_frac
        s[m] = ord("N")
        if n > 60:  # a long run of ambiguous bases too
            st = int(rng.integers(0, n - 50))
            s[st:st + int(rng.integers(1, 50))] = ord("N")
    b = s.tobytes()
    return b.lower() if lower else b


@pytest.mark.parametrize("w,k", [(19, 19), (10, 19), (19, 10), (10, 10), (5, 15), (1, 7), (50, 28), (255, 4), (3, 2)])
def test_sketch_matches_reference(ref, w, k):
    from oracle import refmm2
    from pangraph_b200 import abi
    rng = np.random.default_rng(100 * w + k)
    seqs = [rand_seq(rng, 5000), rand_seq(rng, 3000, n_frac=0.01), rand_seq(rng, 2000, lower=True),
            rand_seq(rng, 4000, repeat=37), rand_seq(rng, 800, n_frac=0.2), b"A" * 700, b"ACGT" * 200, b"AT" * 300,
            rand_seq(rng, w + k - 2), rand_seq(rng, w + k - 1), rand_seq(rng, w + k), rand_seq(rng, k), rand_seq(rng, max(1, k - 1)),
            b"N" * 100, b"NNNNACGTACGGTCAGTCAGCTAGCTAGGGATCGATCGACTCTAGCATCGNNNNNACGATCGATCGATCGACTGACTGACTAGCTAGCATCGATCAGCTAGCAT",
            b"G", rand_seq(rng, 20000, n_frac=0.001)]
    got = abi.sketch(seqs, w, k)
    for i, s in enumerate(seqs):
        want = refmm2.ref_sketch(ref, s, w, k, rid=i)
        g = list(zip((int(v) for v in got[i][0]), (int(v) for v in got[i][1])))
        assert g == want, (i, len(s), len(g), len(want), g[:3], want[:3])


def test_sketch_genome_scale(ref):
    """2 x 1 Mbp with a shared ancestor: every minimizer and its order, k=19 w=19 (asm5/asm10) and w=10 (asm20)."""
    from oracle import refmm2
    from pangraph_b200 import abi, synth
    gs = [g for _, g in synth.genomes(2, length=1_000_000)]
    for w in (19, 10):
        got = abi.sketch(gs, w, 19)
        for i, s in enumerate(gs):
            want = refmm2.ref_sketch(ref, s, w, 19, rid=i)
            assert len(got[i][0]) == len(want)
            assert np.array_equal(got[i][0], np.array([x for x, _ in want], dtype=np.uint64))
            assert np.array_equal(got[i][1], np.array([y for _, y in want], dtype=np.uint64))


@pytest.mark.parametrize("w,k", [(19, 19), (1, 3), (100, 27), (255, 5), (7, 21)])
def test_sketch_fused_path_across_tiles(ref, w, k):
    """Odd k takes the fused tile kernel (4096 positions per tile): lengths around multiples of the tile, ambiguous bases
    and repeats placed around the tile boundaries, several sequences per launch."""
    from oracle import refmm2
    from pangraph_b200 import abi
    rng = np.random.default_rng(7 * w + k)
    seqs = [rand_seq(rng, n) for n in (4095, 4096, 4097, 8191, 8192, 8193, 4096 + w, 4096 + w + k, 4096 - w, 3 * 4096 + 17)]
    s = bytearray(rand_seq(rng, 5 * 4096 + 100))
    for b in (4096, 8192, 12288, 16384):  # Ns just before / on / after the boundaries, at window and k-mer distance
        for d in (-w - k, -w, -k, -1, 0, 1, k - 1, w, w + k - 1):
            if 0 <= b + d < len(s):
                s[b + d] = ord("N")
    seqs.append(bytes(s))
    unit = rand_seq(rng, 31)
    seqs.append(unit * 300)                      # tandem repeat: equal keys inside every window, across the boundaries
    seqs.append(b"A" * 9000)                     # one key everywhere
    seqs.append(rand_seq(rng, 30000, n_frac=0.002))
    got = abi.sketch(seqs, w, k)
    for i, sq in enumerate(seqs):
        want = refmm2.ref_sketch(ref, sq, w, k, rid=i)
        g = list(zip((int(v) for v in got[i][0]), (int(v) for v in got[i][1])))
        assert g == want, (i, len(sq), len(g), len(want), [x for x in zip(g, want) if x[0] != x[1]][:2])
