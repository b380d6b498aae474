"""CPU model of the 16-bit-lane arithmetic of K5a / K5b (pangraph_b200/csrc/ksw_extd2.cu::pair_step), in numpy, against the
scalar recurrence of ksw_extd2 (ksw2_extd2_sse.c:228-275, left-aligned gaps, band not binding): offset-binary halves scaled by
8, the traceback direction riding in the three low bits through an unsigned max, continuation flags from carry bits, packed
subtraction as a plain 32-bit subtraction.  Also pins the identity the kernels use for the first-pass score: the reference's
data-dependent staircase sum (:367-379) equals the boundary value plus the sum of v along the last target column."""
import numpy as np
U32 = np.uint32
def halves(w): return (w & U32(0xffff)).astype(np.int64), (w >> U32(16)).astype(np.int64)
def pack(lo, hi):
    assert ((lo >= 0) & (lo < 65536) & (hi >= 0) & (hi < 65536)).all()
    return (lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(16))).astype(U32)
def viaddmax_s16x2(a, b, c):
    def sg(x): return np.where(x >= 32768, x - 65536, x)
    al, ah = halves(a); bl, bh = halves(b); cl, ch = halves(c)
    lo, hi = np.maximum(sg(al) + sg(bl), sg(cl)), np.maximum(sg(ah) + sg(bh), sg(ch))
    assert (np.abs(lo) < 32768).all() and (np.abs(hi) < 32768).all()
    return pack(lo & 0xffff, hi & 0xffff)
def viaddmax_u16x2(a, b, c):
    al, ah = halves(a); bl, bh = halves(b); cl, ch = halves(c)
    assert ((al + bl) < 65536).all() and ((ah + bh) < 65536).all()  # the kernel never relies on a wrapping unsigned add
    return pack(np.maximum(al + bl, cl), np.maximum(ah + bh, ch))
def vminu2(a, b):
    al, ah = halves(a); bl, bh = halves(b)
    return pack(np.minimum(al, bl), np.minimum(ah, bh))
def add32(*xs):
    s = np.zeros_like(xs[0], dtype=np.int64)
    for x in xs: s = s + np.asarray(x).astype(np.int64)
    return (s & 0xffffffff).astype(U32)
def sub32(a, b, c=0): return ((a.astype(np.int64) - b.astype(np.int64) + int(c)) & 0xffffffff).astype(U32)
def pk(v): v = int(v) & 0xffff; return U32(v | v << 16)

def fill(query, target, mch, mis, scn, q, e, q2, e2):
    qlen, tlen = len(query), len(target)
    if q2 + e2 < q + e: q, q2, e, e2 = q2, q, e2, e
    long_thres = (q2 - q) // (e - e2) - 1 if e != e2 else 0
    if q2 + e2 + long_thres * e2 > q + e + long_thres * e: long_thres += 1
    long_diff = long_thres * (e - e2) - (q2 - q) - e2
    def ufirst(r): return -q - e if r == 0 else -e if r < long_thres else long_diff if r == long_thres else -e2
    BIAS = 0x4000
    NP = (tlen + 1) // 2
    T = np.full(2 * NP + 2, 5, dtype=np.int64); T[:tlen] = target
    # state (one uint32 per pair)
    def mk(lo, hi): return pack(np.asarray(lo, dtype=np.int64), np.asarray(hi, dtype=np.int64))
    t_even, t_odd = np.arange(0, 2 * NP, 2), np.arange(1, 2 * NP + 1, 2)
    uf = np.array([ufirst(t) for t in range(2 * NP + 1)])
    Ub = mk(8 * uf[t_even] + BIAS, 8 * uf[t_odd] + BIAS)
    Vb = np.full(NP, pk(8 * (-q - e) + BIAS), dtype=U32)
    Xb = np.full(NP, pk(8 * (-q - e) + BIAS + 6), dtype=U32); Yb = np.full(NP, pk(8 * (-q - e) + BIAS + 5), dtype=U32)
    X2b = np.full(NP, pk(8 * (-q2 - e2) + BIAS + 4), dtype=U32); Y2b = np.full(NP, pk(8 * (-q2 - e2) + BIAS + 3), dtype=U32)
    MCH = pk(8 * mch + 0x8000 + 7); DM8 = 8 * (mch - mis); SCN = pk(8 * scn + 0x8000 + 7)
    CQ, CQ2 = pk(8 * q + BIAS), pk(8 * q2 + BIAS)
    NQE, NQE2 = pk(-8 * (q + e)), pk(-8 * (q2 + e2))
    FLX, FLY, FLX2, FLY2 = pk(BIAS + 6 - 8 * (q + e)), pk(BIAS + 5 - 8 * (q + e)), pk(BIAS + 4 - 8 * (q2 + e2)), pk(BIAS + 3 - 8 * (q2 + e2))
    C12, C13, C14, C15 = pk(-0x3008), pk(-0x2008), pk(-8), pk(0x3FF8)
    n_row = qlen + tlen - 1
    P = np.zeros((n_row, 2 * NP), dtype=np.uint8)
    H0, last, vsum = 0, 0, 0
    for r in range(n_row):
        st0, en0 = max(0, r - qlen + 1), min(r, tlen - 1)
        p_lo, p_hi = st0 >> 1, en0 >> 1
        sl = slice(p_lo, p_hi + 1)
        if en0 == r and (r & 1):  # first-row cell in the high half of its pair: the pair's previous step wrote garbage there
            p = r >> 1
            Ub[p] = U32((int(Ub[p]) & 0xffff) | ((8 * ufirst(r) + BIAS) << 16))
            Yb[p] = U32((int(Yb[p]) & 0xffff) | ((8 * (-q - e) + BIAS + 5) << 16))
            Y2b[p] = U32((int(Y2b[p]) & 0xffff) | ((8 * (-q2 - e2) + BIAS + 3) << 16))
        # left neighbours (previous anti-diagonal): high half of previous pair | own low half << 16
        def left(Wv, first_lo):
            prev = np.concatenate([[U32(0)], Wv[:-1]])
            w = ((prev >> U32(16)) | (Wv << U32(16))).astype(U32)
            w[0] = U32((int(w[0]) & 0xffff0000) | (first_lo & 0xffff))
            return w
        XT1 = left(Xb, 8 * (-q - e) + BIAS + 6)[sl]; X2T1 = left(X2b, 8 * (-q2 - e2) + BIAS + 4)[sl]; VT1 = left(Vb, 8 * ufirst(r) + BIAS)[sl]
        Uo, Yo, Y2o = Ub[sl], Yb[sl], Y2b[sl]
        # scores
        tl, th = T[2 * p_lo:2 * p_hi + 2:2], T[2 * p_lo + 1:2 * p_hi + 3:2]
        def qat(t):
            j = r - t
            out = np.full(len(t), 6, dtype=np.int64); ok = (j >= 0) & (j < qlen); out[ok] = query[j[ok]]; return out
        tt = np.arange(2 * p_lo, 2 * p_hi + 2, 2)
        ql, qh = qat(tt), qat(tt + 1)
        TQ, QQ = pack(tl, th), pack(ql, qh)
        ne = vminu2(TQ ^ QQ, pk(1))
        Z = sub32(np.full(len(ne), MCH, dtype=U32), (ne.astype(np.int64) * DM8))
        isn = ((TQ | QQ) & U32(0x00040004))
        il, ih = halves(isn); zl, zh = halves(Z); sl_, sh_ = halves(np.full(len(ne), SCN, dtype=U32))
        Z = pack(np.where(il != 0, sl_, zl), np.where(ih != 0, sh_, zh))
        M = viaddmax_u16x2(XT1, VT1, Z); M = viaddmax_u16x2(Yo, Uo, M); M = viaddmax_u16x2(X2T1, VT1, M); M = viaddmax_u16x2(Y2o, Uo, M)
        Z8 = vminu2(M, np.full(len(M), MCH, dtype=U32)) & U32(0xfff8fff8)
        Un = sub32(Z8, VT1); Vn = sub32(Z8, Uo)
        A = sub32(XT1, Un, CQ); B = sub32(Yo, Vn, CQ); A2 = sub32(X2T1, Un, CQ2); B2 = sub32(Y2o, Vn, CQ2)
        for w in (Un, Vn, A, B, A2, B2):  # packing exactness (no borrow between halves)
            lo, hi = halves(w); assert (lo > 0x3000).all() and (lo < 0x5000).all() and (hi > 0x3000).all() and (hi < 0x5000).all()
        # the three negative constants are added lane-wise (VIADD.16x2); + 0x3ff8 cannot carry out of a half (plain 32-bit add)
        def ladd(a, c):
            al, ah = halves(a); cl, ch = halves(np.full(len(a), c, dtype=U32)); return pack((al + cl) & 0xffff, (ah + ch) & 0xffff)
        acc = (ladd(A, C12) & U32(0x10001000)) | (ladd(B, C13) & U32(0x20002000)) | (ladd(A2, C14) & U32(0x40004000)) | (add32(B2, np.full(len(B2), C15, dtype=U32)) & U32(0x80008000))
        D = (acc >> U32(9)) | (~M & U32(0x00070007))
        dl, dh = halves(D)
        Xn = viaddmax_s16x2(A, np.full(len(A), NQE, dtype=U32), np.full(len(A), FLX, dtype=U32)); Yn = viaddmax_s16x2(B, np.full(len(A), NQE, dtype=U32), np.full(len(A), FLY, dtype=U32))
        X2n = viaddmax_s16x2(A2, np.full(len(A), NQE2, dtype=U32), np.full(len(A), FLX2, dtype=U32)); Y2n = viaddmax_s16x2(B2, np.full(len(A), NQE2, dtype=U32), np.full(len(A), FLY2, dtype=U32))
        Ub[sl], Vb[sl], Xb[sl], Yb[sl], X2b[sl], Y2b[sl] = Un, Vn, Xn, Yn, X2n, Y2n
        P[r, 2 * p_lo:2 * p_hi + 2:2] = dl & 0xff; P[r, 2 * p_lo + 1:2 * p_hi + 3:2] = dh & 0xff
        if en0 == tlen - 1:
            vsum += (int(Vb[(tlen - 1) >> 1]) >> (16 * ((tlen - 1) & 1))) & 0xffff
        def cell(W, t):
            lo, hi = halves(W[t >> 1:(t >> 1) + 1]); v = int(hi[0] if t & 1 else lo[0]); return (v - BIAS) >> 3
        if r > 0:
            in0, in1 = st0 <= last <= en0, st0 <= last + 1 <= en0
            d0 = cell(Vb, last) if in0 else 0; d1 = cell(Ub, last + 1) if in1 else 0
            if in0 and in1:
                if d0 > d1: H0 += d0
                else: H0 += d1; last += 1
            elif in0: H0 += d0
            else: last += 1; H0 += d1
        else: H0 = cell(Vb, 0) - (q + e); last = 0
    score_col = ((vsum - qlen * BIAS) >> 3) - min(q + e * tlen, q2 + e2 * tlen)
    return P[:, :tlen], H0, score_col

def scalar(query, target, mch, mis, scn, q, e, q2, e2):
    qlen, tlen = len(query), len(target)
    if q2 + e2 < q + e: q, q2, e, e2 = q2, q, e2, e
    long_thres = (q2 - q) // (e - e2) - 1 if e != e2 else 0
    if q2 + e2 + long_thres * e2 > q + e + long_thres * e: long_thres += 1
    long_diff = long_thres * (e - e2) - (q2 - q) - e2
    u = [-q - e] * tlen; v = list(u); x = list(u); y = list(u); x2 = [-q2 - e2] * tlen; y2 = list(x2)
    n_row = qlen + tlen - 1
    P = np.zeros((n_row, tlen), dtype=np.uint8); H0 = last = 0; qe, qe2 = q + e, q2 + e2
    for r in range(n_row):
        st0, en0 = max(0, r - qlen + 1), min(r, tlen - 1)
        ufirst = -q - e if r == 0 else -e if r < long_thres else long_diff if r == long_thres else -e2
        if en0 == r: y[r] = -q - e; y2[r] = -q2 - e2; u[r] = ufirst
        vn, xn, x2n = {}, {}, {}
        for t in range(st0, en0 + 1):
            if t == 0: xt1, x2t1, vt1 = -q - e, -q2 - e2, ufirst
            else: xt1, vt1, x2t1 = x[t - 1], v[t - 1], x2[t - 1]
            a, b = target[t], query[r - t]
            z = scn if (a == 4 or b == 4) else mch if a == b else mis
            ut = u[t]; A = xt1 + vt1; B = y[t] + ut; A2 = x2t1 + vt1; B2 = y2[t] + ut
            d = 1 if A > z else 0; z = max(z, A)
            d = 2 if B > z else d; z = max(z, B)
            d = 3 if A2 > z else d; z = max(z, A2)
            d = 4 if B2 > z else d; z = max(z, B2)
            z = min(z, mch)
            un, vnew = z - vt1, z - ut
            A -= z - q; B -= z - q; A2 -= z - q2; B2 -= z - q2
            pa, pb, pa2, pb2 = A > 0, B > 0, A2 > 0, B2 > 0
            u[t] = un; vn[t] = vnew
            xn[t] = (A if pa else 0) - qe; y[t] = (B if pb else 0) - qe; x2n[t] = (A2 if pa2 else 0) - qe2; y2[t] = (B2 if pb2 else 0) - qe2
            P[r, t] = d | pa << 3 | pb << 4 | pa2 << 5 | pb2 << 6
        for t in range(st0, en0 + 1): v[t], x[t], x2[t] = vn[t], xn[t], x2n[t]
        if r > 0:
            in0, in1 = st0 <= last <= en0, st0 <= last + 1 <= en0
            if in0 and in1:
                d0, d1 = v[last], u[last + 1]
                if d0 > d1: H0 += d0
                else: H0 += d1; last += 1
            elif in0: H0 += v[last]
            else: last += 1; H0 += u[last]
        else: H0 = v[0] - qe; last = 0
    return P, H0


def test_lane_model_matches_scalar_recurrence():
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kswref
    rng = np.random.default_rng(5)
    for it in range(48):
        ql, tl = int(rng.integers(1, 100)), int(rng.integers(1, 100))
        qq, tt = kswref.random_pair(rng, ql, tl, div=0.08, indel=0.03, n_frac=0.02 if it % 3 == 0 else 0.0,
                                    big_indel=int(rng.integers(0, 30)) if it % 2 else 0)
        if it % 5 == 4:
            qq = rng.integers(0, 4, size=ql).astype(np.uint8)  # unrelated sequences
        pars = [(1, -9, -1, 16, 2, 41, 1), (1, -19, -1, 39, 3, 81, 1), (1, -4, -1, 6, 2, 26, 1), (1, -9, -1, 41, 1, 16, 2)][it % 4]
        P1, h1, sc1 = fill(qq.astype(np.int64), tt.astype(np.int64), *pars)
        P2, h2 = scalar(qq.tolist(), tt.tolist(), *pars)
        for r in range(ql + tl - 1):  # real cells only
            st0, en0 = max(0, r - ql + 1), min(r, tl - 1)
            assert (P1[r, st0:en0 + 1] == P2[r, st0:en0 + 1]).all(), (it, r)
        assert h1 == h2 == sc1, (it, h1, h2, sc1)
