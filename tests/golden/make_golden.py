"""Regenerates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libmm2ref.so) -- run where
/root/reference is mounted:  python tests/golden/make_golden.py
Fixtures: every mm_reg1_t field + CIGAR the reference returns (oracle.refmm2.reg_to_tuple order) and the filtered match
list the host half derives from them, for (a) the reference's own boundary KAT and (b) a small synthetic family with
inversions, repeats and ambiguous bases."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def synth_case():
    from pangraph_b200 import synth
    anc = synth.ancestor(90_000, 21)
    unit = anc[3000:5500].copy()
    for st in (25000, 52000, 80000):
        anc[st:st + len(unit)] = unit
    gs = [synth.mutate(anc, 300 + i, n_rearr=8, len_lo=400, len_hi=9000) for i in range(4)]
    gs[2][11000:11025] = ord("N")
    gs[3] = synth.revcomp(gs[3])
    return [g.tobytes().decode() for g in gs], ["12", "7", "10442385907364519937", "3"]


def kat_case():
    recs = []
    for line in open(os.path.join(HERE, "kat_pair.fa")):
        line = line.strip()
        if line.startswith(">"):
            recs.append([line[1:], ""])
        elif line:
            recs[-1][1] += line
    return [s for _, s in recs], [n for n, _ in recs]


def jsonable(regs):
    return [[list(r[:18]) + [None if r[18] is None else [r[18][0], r[18][1], r[18][2], r[18][3], r[18][4], list(r[18][5]), r[18][6]]]
             for r in q] for q in regs]


def main():
    from oracle import host_half, refmm2
    out = {}
    for name, (seqs, names), preset, k in (("kat", kat_case(), "asm20", 10), ("synth", synth_case(), "asm10", None)):
        regs, mid_occ = refmm2.ref_map_all(seqs, names, preset, k, 90)
        matches = host_half.find_filtered_matches(regs, names, [len(s) for s in seqs])
        out[name] = dict(preset=preset, k=k, names=names, mid_occ=mid_occ, regs=jsonable(regs),
                         matches=[dict(m, cigar=host_half.cigar_str(m["cigar"])) for m in matches])
        n_inv = sum(1 for q in regs for r in q if (r[15] >> 11) & 1)
        n_split = sum(1 for q in regs for r in q if (r[15] >> 8) & 3)
        print(name, "hits", sum(len(q) for q in regs), "inversion hits", n_inv, "split hits", n_split, "filtered matches", len(matches))
    json.dump(out, open(os.path.join(HERE, "golden_regs.json"), "w"))


if __name__ == "__main__":
    main()
