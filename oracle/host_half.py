"""Pure-Python restatement of the HOST half of pangraph's alignment path -- what happens to minimap2's hits before the
graph is rewoven.  TEST INFRASTRUCTURE ONLY (see oracle/pgmm_oracle.c's header): the product's implementation is
pangraph_b200/csrc/matches.cpp; nothing under pangraph_b200/ imports this file.

Parity status: PINNED against the reference's own unit tests (tests/test_host_half.py carries their vectors):
  keep_groups / split_matches / side_patches   packages/pangraph/src/pangraph/split_matches.rs:249-594
  alignment_energy2                            packages/pangraph/src/align/energy.rs:90-112
  filter_matches / is_match_compatible         packages/pangraph/src/pangraph/graph_merging.rs:254-375

An alignment is a dict: qry/ref = (block_id, length, start, end), matches, length, quality, reverse (bool),
cigar = [(len, op)], divergence (float or None), align (float or None).
"""
import copy
import re

MATCH_OPS = "M=X"


def parse_cigar(s):
    return [(int(n), op) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", s.replace(" ", ""))]


def cigar_str(c):
    return "".join(f"{n}{op}" for n, op in c)


def keep_groups(cigar, threshold):
    """split_matches.rs:32-92"""
    groups = []
    g_start = last_match = None
    m_sum = i_sum = d_sum = 0
    for i, (n, op) in enumerate(cigar):
        if g_start is None:
            if op not in MATCH_OPS:
                continue
            g_start = i
        if op in MATCH_OPS:
            m_sum += n
            i_sum = d_sum = 0
            last_match = i
        elif op == "I":
            i_sum += n
        elif op == "D":
            d_sum += n
        else:
            raise ValueError(f"Unexpected CIGAR operation: '{op}'")
        if max(i_sum, d_sum) >= threshold:
            if g_start is not None and last_match is not None and m_sum >= threshold:
                groups.append((g_start, last_match))
            g_start = last_match = None
            m_sum = i_sum = d_sum = 0
    if g_start is not None and last_match is not None and m_sum >= threshold:
        groups.append((g_start, last_match))
    return groups


def _pos(cigar, idx, ops, inclusive):
    pos = 0
    for i, (n, op) in enumerate(cigar):
        if i == idx and not inclusive:
            return pos
        if op in ops:
            pos += n
        if i == idx:
            return pos
    raise AssertionError


def generate_subalignment(aln, group):
    """split_matches.rs:151-185"""
    g0, g1 = group
    c = aln["cigar"]
    qs, qe = _pos(c, g0, "MI=X", False), _pos(c, g1, "MI=X", True)
    rs, re_ = _pos(c, g0, "MD=X", False), _pos(c, g1, "MD=X", True)
    qn, ql, q0, q1 = aln["qry"]
    rn, rl, r0, _ = aln["ref"]
    if not aln["reverse"]:
        q = (qn, ql, q0 + qs, q0 + qe)
    else:
        q = (qn, ql, q1 - qe, q1 - qs)
    sub = c[g0:g1 + 1]
    out = copy.deepcopy(aln)
    out.update(qry=q, ref=(rn, rl, r0 + rs, r0 + re_), cigar=sub,
               matches=sum(n for n, op in sub if op in MATCH_OPS), length=sum(n for n, _ in sub))
    return out


def add_flanking_indel(cigar, kind, add_len, leading):
    """bam/cigar.rs:60-96"""
    order = range(len(cigar)) if leading else range(len(cigar) - 1, -1, -1)
    replace = None
    for i in order:
        n, op = cigar[i]
        if op in MATCH_OPS:
            break
        if op == kind:
            replace = i
    out = list(cigar)
    if replace is not None:
        out[replace] = (out[replace][0] + add_len, kind)
    else:
        out.insert(0 if leading else len(out), (add_len, kind))
    return out


def side_patches(aln, threshold):
    """split_matches.rs:189-237"""
    ops = list(aln["cigar"])
    rn, rl, rs, re_ = aln["ref"]
    if 0 < rs < threshold:
        aln["ref"] = (rn, rl, 0, aln["ref"][3])
        aln["length"] += rs
        ops = add_flanking_indel(ops, "D", rs, True)
    if re_ < rl and rl - re_ < threshold:
        aln["ref"] = (rn, rl, aln["ref"][2], rl)
        aln["length"] += rl - re_
        ops = add_flanking_indel(ops, "D", rl - re_, False)
    qn, ql, qs, qe = aln["qry"]
    if 0 < qs < threshold:
        aln["qry"] = (qn, ql, 0, aln["qry"][3])
        aln["length"] += qs
        ops = add_flanking_indel(ops, "I", qs, not aln["reverse"])
    if qe < ql and ql - qe < threshold:
        aln["qry"] = (qn, ql, aln["qry"][2], ql)
        aln["length"] += ql - qe
        ops = add_flanking_indel(ops, "I", ql - qe, bool(aln["reverse"]))
    aln["cigar"] = ops


def split_matches(aln, threshold):
    """split_matches.rs:13-24"""
    out = [generate_subalignment(aln, g) for g in keep_groups(aln["cigar"], threshold)]
    for a in out:
        side_patches(a, threshold)
    return out


def alignment_energy2(aln, alpha, beta):
    """energy.rs:37-54"""
    L = aln["matches"]
    M = (aln["divergence"] or 0.0) * float(L)
    C = 4
    _, ql, qs, qe = aln["qry"]
    _, rl, rs, re_ = aln["ref"]
    C -= (qs == 0) + (qe == ql) + (rs == 0) + (re_ == rl)
    return -float(L) + float(C) * alpha + M * beta


def filter_matches(alns, alpha, beta):
    """graph_merging.rs:187-242"""
    keyed = [(alignment_energy2(a, alpha, beta), i) for i, a in enumerate(alns)]
    keyed = [k for k in keyed if k[0] < 0.0]
    keyed.sort(key=lambda k: k[0])  # stable
    accepted, out = {}, []

    def free(name, s, e):
        return not any(ie > s and is_ < e for is_, ie in accepted.get(name, []))

    for _, i in keyed:
        a = alns[i]
        if free(a["ref"][0], a["ref"][2], a["ref"][3]) and free(a["qry"][0], a["qry"][2], a["qry"][3]):
            out.append(a)
            accepted.setdefault(a["ref"][0], []).append((a["ref"][2], a["ref"][3]))
            accepted.setdefault(a["qry"][0], []).append((a["qry"][2], a["qry"][3]))
    return out


def from_reg(reg, qname, qlen, tnames, tlens):
    """Alignment::from_minimap_paf_obj (align_with_minimap2_lib.rs:89-121) over oracle.refmm2.reg_to_tuple output."""
    rid, qs, qe, rs, re_ = reg[2], reg[4], reg[5], reg[6], reg[7]
    cap, dp_score, dp_max, dp_max2, n_ambi, cig, de = reg[18]
    return dict(qry=(int(qname), qlen, qs, qe), ref=(int(tnames[rid]), tlens[rid], rs, re_), matches=reg[11], length=reg[12],
                quality=reg[15] & 0xff, reverse=bool((reg[15] >> 10) & 1), cigar=[(c >> 4, "MIDNSHP=XB"[c & 15]) for c in cig],
                divergence=de, align=float(dp_score))


def find_filtered_matches(regs_per_query, names, lens, threshold=100, alpha=100.0, beta=10.0):
    """The alignment part of self_merge (graph_merging.rs:95-121) on the reference's hits, queries in BlockId order."""
    order = sorted(range(len(names)), key=lambda i: int(names[i]))
    alns = []
    for q in order:
        for reg in regs_per_query[q]:
            a = from_reg(reg, names[q], lens[q], names, lens)
            if a["qry"][0] != a["ref"][0]:
                alns.extend(split_matches(a, threshold))
    return filter_matches(alns, alpha, beta)
