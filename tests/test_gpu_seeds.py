"""K1+K3 (pgmm_collect_seeds) against the reference's own seeding stage: collect_minimizers + mm_seed_mz_flt +
collect_seed_hits (map.c:59-76,168-204; seed.c) called directly in oracle/_ref/libmm2ref_stage.so.  The reference sorts
the anchors before it returns them (map.c:202, an unstable sort), the CUDA stage returns them before that sort, so the
anchor lists are compared as sorted multisets; mini_pos and rep_len are compared as they are."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def both(seqs, names, preset="asm10", k=None):
    from oracle import refmm2
    from pangraph_b200 import abi
    if not os.path.exists(refmm2.STAGE_SO):
        pytest.skip("oracle/_ref/libmm2ref_stage.so not built")
    lib = refmm2.load_ref_stage()
    ridx = refmm2.Index(lib, seqs, names, preset, k, 90)
    idx = abi.Index(seqs, names, preset, k, 90)
    assert idx.mo.mid_occ == ridx.mo.mid_occ
    got = idx.collect_seeds()
    n_anchor = 0
    for i, (s, nm) in enumerate(zip(seqs, names)):
        wa, wm, wrep = refmm2.ref_collect_seed_hits(lib, ridx, s, nm)
        ga, gm, grep = got[i]
        assert grep == wrep, (i, grep, wrep)
        assert np.array_equal(gm, wm), (i, len(gm), len(wm))
        assert len(ga) == len(wa), (i, len(ga), len(wa))
        if len(ga):
            go = np.lexsort((ga[:, 1], ga[:, 0]))
            wo = np.lexsort((wa[:, 1], wa[:, 0]))
            assert np.array_equal(ga[go], wa[wo]), i
        n_anchor += len(wa)
    idx.close()
    ridx.close()
    return n_anchor


@pytest.mark.parametrize("preset", ["asm5", "asm10", "asm20"])
def test_collect_seeds_small_family(preset):
    from pangraph_b200 import synth
    gs = synth.genomes(5, length=60_000, n_rearr=6, len_lo=300, len_hi=8000)
    names = [str(v) for v in (3, 17, 5, 10442385907364519937, 100)]
    assert both([g for _, g in gs], names, preset) > 1000


def test_collect_seeds_repeats_and_high_occurrence():
    """Tandem arrays (minimizers above mid_occ: streak thinning, rep_len, MM_SEED_TANDEM), dispersed repeats, self hits
    (MM_SEED_SELF, NO_DIAG), Ns, a reverse-complemented genome."""
    from pangraph_b200 import synth
    anc = synth.ancestor(80_000, 7)
    unit = anc[1000:3500].copy()
    for st in (9000, 20000, 41000, 66000):
        anc[st:st + len(unit)] = unit
    short = anc[500:560].copy()
    for st in range(30000, 36000, 60):
        anc[st:st + 60] = short
    gs = [synth.mutate(anc, 900 + i, n_rearr=4, len_lo=300, len_hi=6000) for i in range(4)]
    gs[1][5000:5040] = ord("N")
    gs[2] = synth.revcomp(gs[2])
    assert both([g.tobytes() for g in gs], ["0", "1", "2", "3"]) > 1000


def test_collect_seeds_megabase_pair():
    from pangraph_b200 import synth
    gs = synth.genomes(2, length=1_000_000)
    assert both([g for _, g in gs], ["0", "1"]) > 50_000


def test_collect_seeds_real_pair():
    import realdata
    seqs, names = realdata.load_pair("ecoli")
    assert both(seqs, ["1", "2"]) > 400_000
