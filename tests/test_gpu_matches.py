"""The round as pangraph sees it, on the GPU: align_with_minimap2_lib and the alignment half of self_merge through the
C-ABI against the oracle (reference hits -> oracle/host_half.py)."""
import pytest

pytestmark = pytest.mark.gpu


def norm(ms):
    from oracle import host_half
    out = []
    for m in ms:
        d = dict(m)
        d["qry"], d["ref"] = tuple(int(v) for v in m["qry"]), tuple(int(v) for v in m["ref"])
        d["cigar"] = host_half.cigar_str(m["cigar"])
        d["matches"], d["length"], d["quality"] = int(m["matches"]), int(m["length"]), int(m["quality"])
        return_align = d.pop("align")
        d["align"] = float(return_align) if return_align is not None else None
        out.append(d)
    return out


@pytest.mark.parametrize("sens,thr", [(10, 100), (20, 100), (5, 50)])
def test_round_matches_oracle(ref, sens, thr):
    from oracle import host_half, refmm2
    from pangraph_b200 import abi, synth
    gs = synth.genomes(6, length=70_000, n_rearr=8, len_lo=300, len_hi=9000)
    # pangraph hands the blocks over in BTreeMap order, i.e. ascending BlockId (align_with_minimap2_lib.rs:19-22)
    ids = sorted([901, 17, 5, 10442385907364519937, 100, 9])
    seqs = [g.decode() for _, g in gs]
    names = [str(i) for i in ids]
    preset = {5: "asm5", 10: "asm10", 20: "asm20"}[sens]
    regs, _ = refmm2.ref_map_all(seqs, names, preset, None, max(thr - 10, 5), threads=6)
    args = abi.alignment_args(indel_len_threshold=thr, sensitivity=sens)
    blocks = dict(reversed(list(zip(ids, seqs))))  # caller order must not matter
    # align_with_minimap2_lib: every hit, queries in BlockId order
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    want_all = [host_half.from_reg(r, names[q], len(seqs[q]), names, [len(s) for s in seqs]) for q in order for r in regs[q]]
    got_all = abi.align_with_minimap2_lib(blocks, args)
    assert norm(got_all) == norm(want_all)
    # find_matches -> drop self hits -> split -> filter
    want = host_half.find_filtered_matches(regs, names, [len(s) for s in seqs], thr, args.alpha, args.beta)
    got = abi.find_filtered_matches(blocks, args)
    assert norm(got) == norm(want) and len(want) > 3


def test_error_behaviour():
    from pangraph_b200 import abi
    with pytest.raises(ValueError, match="Unknown sensitivity preset"):
        abi.align_with_minimap2_lib({1: "ACGT" * 100, 2: "ACGT" * 100}, abi.alignment_args(sensitivity=7))
    assert abi.align_with_minimap2_lib({}, abi.alignment_args()) == []
    assert abi.find_filtered_matches({5: "ACGTTGCA" * 40}, abi.alignment_args()) == []
