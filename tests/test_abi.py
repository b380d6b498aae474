"""The C-ABI without a GPU: the library loads, exports every symbol include/pgmm_b200.h declares, and lays its structs
out exactly like the reference's minimap.h (offsets the Rust wrapper reads directly)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pgmm_b200.h")
REF_H = "/root/reference/packages/minimap2-sys/minimap2"

PROBE = r'''
#include <stdio.h>
#include <stddef.h>
#include "%s"
#define P(t, f) printf(#t "." #f " %%zu\n", offsetof(t, f))
int main(void) {
  printf("sizeof.mm_idx_seq_t %%zu\n", sizeof(mm_idx_seq_t)); printf("sizeof.mm_idx_t %%zu\n", sizeof(mm_idx_t));
  printf("sizeof.mm_extra_t %%zu\n", sizeof(mm_extra_t)); printf("sizeof.mm_reg1_t %%zu\n", sizeof(mm_reg1_t));
  printf("sizeof.mm_idxopt_t %%zu\n", sizeof(mm_idxopt_t)); printf("sizeof.mm_mapopt_t %%zu\n", sizeof(mm_mapopt_t));
  P(mm_idx_seq_t, name); P(mm_idx_seq_t, offset); P(mm_idx_seq_t, len); P(mm_idx_seq_t, is_alt);
  P(mm_idx_t, b); P(mm_idx_t, k); P(mm_idx_t, n_seq); P(mm_idx_t, seq); P(mm_idx_t, S); P(mm_idx_t, h);
  P(mm_extra_t, capacity); P(mm_extra_t, dp_score); P(mm_extra_t, dp_max2); P(mm_extra_t, n_cigar); P(mm_extra_t, cigar);
  P(mm_reg1_t, id); P(mm_reg1_t, qs); P(mm_reg1_t, parent); P(mm_reg1_t, as); P(mm_reg1_t, score0); P(mm_reg1_t, hash);
  P(mm_reg1_t, div); P(mm_reg1_t, p);
  P(mm_idxopt_t, bucket_bits); P(mm_idxopt_t, batch_size);
  P(mm_mapopt_t, flag); P(mm_mapopt_t, bw); P(mm_mapopt_t, chain_gap_scale); P(mm_mapopt_t, best_n); P(mm_mapopt_t, a);
  P(mm_mapopt_t, zdrop); P(mm_mapopt_t, min_dp_max); P(mm_mapopt_t, mid_occ_frac); P(mm_mapopt_t, mid_occ);
  P(mm_mapopt_t, occ_dist); P(mm_mapopt_t, max_sw_mat); P(mm_mapopt_t, split_prefix);
  mm_reg1_t r; unsigned *w = (unsigned*)((char*)&r + 60);
  *w = 0; r.mapq = 0xff; printf("bits.mapq %%x\n", *w); *w = 0; r.rev = 1; printf("bits.rev %%x\n", *w);
  *w = 0; r.inv = 1; printf("bits.inv %%x\n", *w); *w = 0; r.split = 3; printf("bits.split %%x\n", *w);
  *w = 0; r.split_inv = 1; printf("bits.split_inv %%x\n", *w);
  return 0;
}
'''


def probe(header, extra_inc=()):
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write(PROBE % header)
        subprocess.run(["gcc", "-std=gnu11", "-w", src, "-o", exe] + [f"-I{i}" for i in extra_inc], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return dict(line.split() for line in out.strip().splitlines())


def test_struct_layouts_measured():
    got = probe(HEADER)
    assert got["sizeof.mm_idx_seq_t"] == "24" and got["sizeof.mm_idx_t"] == "80" and got["sizeof.mm_extra_t"] == "24"
    assert got["sizeof.mm_reg1_t"] == "80" and got["sizeof.mm_idxopt_t"] == "24" and got["sizeof.mm_mapopt_t"] == "248"
    assert got["mm_idx_t.n_seq"] == "16" and got["mm_idx_t.seq"] == "32" and got["mm_reg1_t.hash"] == "64"
    assert got["mm_reg1_t.div"] == "68" and got["mm_reg1_t.p"] == "72" and got["mm_extra_t.cigar"] == "24"
    assert got["bits.mapq"] == "ff" and got["bits.rev"] == "400" and got["bits.inv"] == "800" and got["bits.split"] == "300"


@pytest.mark.skipif(not os.path.isdir(REF_H), reason="reference headers not mounted")
def test_struct_layouts_equal_reference_header():
    assert probe(HEADER) == probe(os.path.join(REF_H, "minimap.h"))


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)     # preprocessor lines
    return sorted(set(re.findall(r"PGMM_API\s+[^;(]*?\b(\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pangraph_b200 import abi
    lib = abi.lib()  # must load without a GPU
    names = declared_symbols()
    assert len(names) >= 25 and "mm_map" in names and "pgmm_map_batch" in names and "pgmm_find_filtered_matches" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    nm = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    assert set(names) <= exported
    leaked = [s for s in exported if not (s.startswith("mm_") or s.startswith("pgmm_"))]
    assert not leaked, leaked  # nothing but the boundary is visible


def test_no_device_count_call_crashes():
    from pangraph_b200 import abi
    assert abi.lib().pgmm_device_count() >= 0


def test_options_match_reference(ref):
    """mm_set_opt / mm_check_opt byte for byte against the reference's options.c, every preset it knows."""
    from oracle import refmm2
    from pangraph_b200 import abi
    L = abi.lib()
    for preset in [None, "map-ont", "ava-ont", "map10k", "map-pb", "ava-pb", "map-hifi", "map-ccs", "asm5", "asm10", "asm20",
                   "short", "sr", "splice", "splice:hq", "cdna", "asm7", "nonsense"]:
        a_io, a_mo, b_io, b_mo = abi.mm_idxopt_t(), abi.mm_mapopt_t(), refmm2.mm_idxopt_t(), refmm2.mm_mapopt_t()
        assert L.mm_set_opt(None, C.byref(a_io), C.byref(a_mo)) == ref.mm_set_opt(None, C.byref(b_io), C.byref(b_mo)) == 0
        if preset is not None:
            ra = L.mm_set_opt(preset.encode(), C.byref(a_io), C.byref(a_mo))
            rb = ref.mm_set_opt(preset.encode(), C.byref(b_io), C.byref(b_mo))
            assert ra == rb, preset
            if ra != 0:
                continue
        assert bytes(a_io) == bytes(b_io), preset
        assert bytes(a_mo) == bytes(b_mo), preset
        assert L.mm_check_opt(C.byref(a_io), C.byref(a_mo)) == ref.mm_check_opt(C.byref(b_io), C.byref(b_mo))
    # the validation codes, rule by rule
    def both(mut):
        a_io, a_mo, b_io, b_mo = abi.mm_idxopt_t(), abi.mm_mapopt_t(), refmm2.mm_idxopt_t(), refmm2.mm_mapopt_t()
        for lib_, io, mo in ((L, a_io, a_mo), (ref, b_io, b_mo)):
            lib_.mm_set_opt(None, C.byref(io), C.byref(mo))
            lib_.mm_set_opt(b"asm10", C.byref(io), C.byref(mo))
            mut(io, mo)
        return L.mm_check_opt(C.byref(a_io), C.byref(a_mo)), ref.mm_check_opt(C.byref(b_io), C.byref(b_mo))
    muts = [lambda io, mo: setattr(mo, "bw", mo.bw_long + 1), lambda io, mo: setattr(mo, "flag", mo.flag | 0x1000),
            lambda io, mo: setattr(io, "k", 0), lambda io, mo: setattr(mo, "best_n", -1), lambda io, mo: setattr(mo, "pri_ratio", 1.5),
            lambda io, mo: setattr(mo, "flag", mo.flag | 0x300000), lambda io, mo: setattr(mo, "e", 0),
            lambda io, mo: setattr(mo, "e2", 5), lambda io, mo: setattr(mo, "q2", 120), lambda io, mo: setattr(mo, "zdrop", 10),
            lambda io, mo: setattr(mo, "flag", mo.flag | 0x4000 | 0x800000), lambda io, mo: setattr(mo, "flag", mo.flag | 0x100000000 | 0x8)]
    for m in muts:
        a, b = both(m)
        assert a == b and a != 0
