/* pgmm_oracle.c -- plain-C CPU restatement of the reference's alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it.  Nothing under pangraph_b200/ links or calls it, and it must never be
 * the thing that is measured or shipped.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_ksw.py, test_oracle_sketch.py, test_oracle_chain.py) against the unmodified
 * vendored minimap2 C of the reference, compiled from /root/reference by oracle/Makefile into
 * oracle/_ref/libmm2ref.so, which itself reproduces the reference's only boundary golden vector
 * (packages/pangraph/src/align/minimap2_lib/align_with_minimap2_lib.rs:135-204).
 *
 * Citations use C/ = packages/minimap2-sys/minimap2/ of the reference.
 * The code is scalar and written per cell / per element on purpose: it states WHAT the reference's SSE and
 * macro-generated code computes, one value at a time, so that a data-parallel kernel can be compared with it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NEG_INF (-0x40000000) /* C/ksw2.h:6 */
/* flag bits, C/ksw2.h:8-14 */
#define ORC_EZ_RIGHT 0x02
#define ORC_EZ_APPROX_MAX 0x08
#define ORC_EZ_EXTZ_ONLY 0x40
#define ORC_EZ_REV_CIGAR 0x80

typedef struct {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end, n_cigar;
} orc_ez_t;

/* ------------------------------------------------------------------------------------------------------------
 * ksw_extd2: dual-affine banded extension/global DP with traceback.  C/ksw2_extd2_sse.c:34-401.
 *
 * State per target position t (one signed byte each, wrap-around arithmetic): u,v,x,y,x2,y2 are the
 * Suzuki-Kasahara difference values of the previous anti-diagonal, s the substitution score last written for t.
 * The reference computes 16 positions at a time, so on every anti-diagonal r it evaluates the padded range
 * [st,en] (st rounded down, en rounded up to 16) although only [st0,en0] is inside the band; the padded cells
 * use whatever s[t] holds.  Those cells are observable (band edges, traceback), so they are restated too.
 * ---------------------------------------------------------------------------------------------------------- */

static inline int8_t w8(int v) { return (int8_t)(uint8_t)v; } /* wrap to int8 like _mm_add_epi8/_mm_sub_epi8 */

/* push one op of length 1 with run-length merging, C/ksw2.h:111-121 */
static void orc_push(uint32_t *cigar, int *n, uint32_t op, int len)
{
	if (*n == 0 || op != (cigar[*n - 1] & 0xf)) cigar[(*n)++] = (uint32_t)len << 4 | op;
	else cigar[*n - 1] += (uint32_t)len << 4;
}

int orc_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                  int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag,
                  orc_ez_t *ez, uint32_t *cigar /* capacity >= qlen+tlen+2 */)
{
	const int m = 5;
	int r, t, approx_max = !!(flag & ORC_EZ_APPROX_MAX);
	/* ksw_reset_extz, C/ksw2.h:161-166 */
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = ORC_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
	if (qlen <= 0 || tlen <= 0) return 0;
	if (q2 + e2 < q + e) { t = q, q = q2, q2 = t, t = e, e = e2, e2 = t; } /* :73 */
	int8_t sc_mch = mat[0], sc_mis = mat[1], sc_N = mat[m * m - 1] == 0 ? w8(-e2) : mat[m * m - 1]; /* :81-83 */
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
	int n_col = qlen < tlen ? qlen : tlen;
	n_col = ((n_col < w + 1 ? n_col : w + 1) + 15) / 16 + 1; /* :89-91 */
	{ /* :93-97 */
		int min_sc = mat[1];
		for (t = 1; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
		if (-min_sc > 2 * (q + e)) return 0;
	}
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0; /* :99-102 */
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	int long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	int T = tlen_ * 16, Q = (qlen_ + 1) * 16;
	/* one allocation laid out like the reference's (:104-107): ... s | sf | qr, so that reads past the end of the
	 * target copy land in the reversed query, and reads past the reversed query land in zeros */
	int8_t *u = malloc((size_t)T * 7 + T + Q + 16), *v = u + T, *x = v + T, *y = x + T, *x2 = y + T, *y2 = x2 + T, *s = y2 + T;
	uint8_t *sf = (uint8_t*)(s + T), *qr = sf + T;
	memset(u, w8(-q - e), (size_t)T * 4);
	memset(x2, w8(-q2 - e2), (size_t)T * 2);
	memset(s, 0, (size_t)T + T + Q + 16);
	int32_t *H = 0;
	if (!approx_max) {
		H = malloc(sizeof(int32_t) * T);
		for (t = 0; t < T; ++t) H[t] = ORC_NEG_INF;
	}
	int n_row = qlen + tlen - 1, stride = n_col * 16;
	uint8_t *p = malloc((size_t)n_row * stride + 16);
	int *off = malloc(sizeof(int) * 2 * n_row), *off_end = off + n_row;
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);

	int last_st = -1, last_en = -1, H0 = 0, last_H0_t = 0, qe = q + e, qe2 = q2 + e2;
	for (r = 0; r < n_row; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { ez->zdropped = 1; break; } /* :142-145 */
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		/* values standing in for position st-1 of the previous anti-diagonal, :149-159 */
		int8_t x1, x21, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = w8(-q - e), x21 = w8(-q2 - e2), v1 = w8(-q - e);
		} else {
			x1 = w8(-q - e), x21 = w8(-q2 - e2);
			v1 = r == 0 ? w8(-q - e) : r < long_thres ? w8(-e) : r == long_thres ? w8(long_diff) : w8(-e2);
		}
		if (en >= r) { /* first row of the matrix, :160-163 */
			y[r] = w8(-q - e), y2[r] = w8(-q2 - e2);
			u[r] = r == 0 ? w8(-q - e) : r < long_thres ? w8(-e) : r == long_thres ? w8(long_diff) : w8(-e2);
		}
		/* substitution scores: written in runs of 16 starting at st0, :165-180 (the run may pass en0 and, when
		 * tlen is a multiple of 16, the end of s[] -- it then lands in sf[0..], which is never read again) */
		const uint8_t *qrr = qr + (qlen - 1 - r);
		for (t = st0; t <= en0; t += 16) {
			int k;
			for (k = 0; k < 16; ++k) {
				uint8_t a = sf[t + k], b = qrr[t + k];
				s[t + k] = (a == m - 1 || b == m - 1) ? sc_N : a == b ? sc_mch : sc_mis;
			}
		}
		/* the recurrence, one cell at a time; xp/vp/x2p carry the previous anti-diagonal's value at t-1 */
		int8_t xp = x1, vp = v1, x2p = x21;
		uint8_t *pr = p + (size_t)r * stride - st;
		off[r] = st, off_end[r] = en;
		for (t = st; t <= en; ++t) {
			int8_t z = s[t], xt1 = xp, vt1 = vp, x2t1 = x2p, ut = u[t], d;
			xp = x[t], vp = v[t], x2p = x2[t];
			int8_t a = w8(xt1 + vt1), b = w8(y[t] + ut), a2 = w8(x2t1 + vt1), b2 = w8(y2[t] + ut);
			if (!(flag & ORC_EZ_RIGHT)) { /* gap left-alignment, :228-275 */
				d = a > z ? 1 : 0;  z = z > a ? z : a;
				d = b > z ? 2 : d;  z = z > b ? z : b;
				d = a2 > z ? 3 : d; z = z > a2 ? z : a2;
				d = b2 > z ? 4 : d; z = z > b2 ? z : b2;
			} else { /* gap right-alignment, :276-322 */
				d = z > a ? 0 : 1;  z = z > a ? z : a;
				d = z > b ? d : 2;  z = z > b ? z : b;
				d = z > a2 ? d : 3; z = z > a2 ? z : a2;
				d = z > b2 ? d : 4; z = z > b2 ? z : b2;
			}
			z = z < sc_mch ? z : sc_mch;
			u[t] = w8(z - vt1), v[t] = w8(z - ut);
			int8_t tmp = w8(z - q), tmp2 = w8(z - q2);
			a = w8(a - tmp), b = w8(b - tmp), a2 = w8(a2 - tmp2), b2 = w8(b2 - tmp2);
			if (!(flag & ORC_EZ_RIGHT)) {
				x[t]  = w8((a  > 0 ? a  : 0) - qe);  if (a  > 0) d |= 0x08;
				y[t]  = w8((b  > 0 ? b  : 0) - qe);  if (b  > 0) d |= 0x10;
				x2[t] = w8((a2 > 0 ? a2 : 0) - qe2); if (a2 > 0) d |= 0x20;
				y2[t] = w8((b2 > 0 ? b2 : 0) - qe2); if (b2 > 0) d |= 0x40;
			} else {
				x[t]  = w8((a  >= 0 ? a  : 0) - qe);  if (a  >= 0) d |= 0x08;
				y[t]  = w8((b  >= 0 ? b  : 0) - qe);  if (b  >= 0) d |= 0x10;
				x2[t] = w8((a2 >= 0 ? a2 : 0) - qe2); if (a2 >= 0) d |= 0x20;
				y2[t] = w8((b2 >= 0 ? b2 : 0) - qe2); if (b2 >= 0) d |= 0x40;
			}
			pr[t] = (uint8_t)d;
		}
		if (!approx_max) { /* exact max with a 32-bit score per position, :323-366 */
			int32_t max_H, max_t;
			if (r > 0) {
				int en1 = st0 + (en0 - st0) / 4 * 4, i;
				int32_t HH[4], tt[4];
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4) /* four interleaved running maxima, first strict maximum wins per lane */
					for (i = 0; i < 4; ++i) {
						H[t + i] += v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en0;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			/* ksw_apply_zdrop, C/ksw2.h:168-184, with e2 */
			if (max_H > ez->max) ez->max = max_H, ez->max_t = max_t, ez->max_q = r - max_t;
			else if (max_t >= ez->max_t && r - max_t >= ez->max_q) {
				int tl = max_t - ez->max_t, ql = (r - max_t) - ez->max_q, l = tl > ql ? tl - ql : ql - tl;
				if (zdrop >= 0 && ez->max - max_H > zdrop + l * e2) { ez->zdropped = 1; break; }
			}
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* one tracked cell, :367-383 (KSW_EZ_APPROX_DROP is never set on pangraph's path) */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += v[last_H0_t];
				else ++last_H0_t, H0 += u[last_H0_t];
			} else H0 = v[0] - qe, last_H0_t = 0;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	/* traceback, C/ksw2.h:127-159 and C/ksw2_extd2_sse.c:388-399 */
	{
		int i = -1, j = -1, n = 0, state = 0, go = 1;
		if (!ez->zdropped && !(flag & ORC_EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
		else if (!ez->zdropped && (flag & ORC_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) ez->reach_end = 1, i = ez->mqe_t, j = qlen - 1;
		else if (ez->max_t >= 0 && ez->max_q >= 0) i = ez->max_t, j = ez->max_q;
		else go = 0;
		if (go) {
			while (i >= 0 && j >= 0) {
				int force = -1, tmp;
				r = i + j;
				if (i < off[r]) force = 2;
				if (i > off_end[r]) force = 1;
				tmp = force < 0 ? p[(size_t)r * stride + i - off[r]] : 0;
				if (state == 0) state = tmp & 7;
				else if (!(tmp >> (state + 2) & 1)) state = 0;
				if (state == 0) state = tmp & 7;
				if (force >= 0) state = force;
				if (state == 0) orc_push(cigar, &n, 0, 1), --i, --j;
				else if (state == 1 || state == 3) orc_push(cigar, &n, 2, 1), --i;
				else orc_push(cigar, &n, 1, 1), --j;
			}
			if (i >= 0) orc_push(cigar, &n, 2, i + 1);
			if (j >= 0) orc_push(cigar, &n, 1, j + 1);
			if (!(flag & ORC_EZ_REV_CIGAR))
				for (i = 0; i < n >> 1; ++i) { uint32_t c = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = c; }
			ez->n_cigar = n;
		}
	}
	free(u); free(H); free(p); free(off);
	return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * The same recurrence for problems whose band never binds (w >= max(qlen, tlen): every gap fill on pangraph's
 * path, C/align.c:733-755), stated WITHOUT the 16-lane artefacts: only cells inside the matrix are evaluated, in
 * plain int arithmetic.  This is what a kernel may compute when padded cells cannot reach any real cell: a real
 * cell (r,t) reads (r-1,t-1) and (r-1,t), which are real cells or the first-row / first-column boundary values,
 * and the traceback never leaves the matrix.  Real cells never overflow a signed byte -- the reference's own
 * precondition (q+e)+(q2+e2) <= 127, C/options.c:205 -- so int arithmetic equals the wrap-around byte arithmetic.
 * tests/test_oracle_ksw.py checks this function against ksw_extd2_sse on unbanded problems.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_ksw_extd2_unbanded(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                           int q, int e, int q2, int e2, int zdrop, int end_bonus, int flag,
                           orc_ez_t *ez, uint32_t *cigar, int *overflow /* set to 1 if any value left [-128,127] */)
{
	const int m = 5;
	int r, t, approx_max = !!(flag & ORC_EZ_APPROX_MAX);
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = ORC_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
	*overflow = 0;
	if (qlen <= 0 || tlen <= 0) return 0;
	if (q2 + e2 < q + e) { t = q, q = q2, q2 = t, t = e, e = e2, e2 = t; }
	int sc_mch = mat[0], sc_mis = mat[1], sc_N = mat[m * m - 1] == 0 ? -e2 : mat[m * m - 1];
	{
		int min_sc = mat[1];
		for (t = 1; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
		if (-min_sc > 2 * (q + e)) return 0;
	}
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	int *u = malloc(sizeof(int) * tlen * 8), *v = u + tlen, *x = v + tlen, *y = x + tlen, *x2 = y + tlen, *y2 = x2 + tlen;
	int *vn = y2 + tlen, *xn = vn + tlen; /* v and x of the anti-diagonal being written (x2 reuses a third copy below) */
	int *x2n = malloc(sizeof(int) * tlen);
	int32_t *H = malloc(sizeof(int32_t) * tlen);
	for (t = 0; t < tlen; ++t) u[t] = v[t] = x[t] = y[t] = -q - e, x2[t] = y2[t] = -q2 - e2, H[t] = ORC_NEG_INF;
	int n_row = qlen + tlen - 1;
	uint8_t *p = malloc((size_t)n_row * tlen);
	int H0 = 0, last_H0_t = 0, qe = q + e, qe2 = q2 + e2;
#define ORC_CHK(val) do { int v_ = (val); if (v_ < -128 || v_ > 127) *overflow = 1; } while (0)
	for (r = 0; r < n_row; ++r) {
		int st0 = r - qlen + 1 > 0 ? r - qlen + 1 : 0, en0 = r < tlen - 1 ? r : tlen - 1;
		int ufirst = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		if (en0 == r) y[r] = -q - e, y2[r] = -q2 - e2, u[r] = ufirst; /* first row */
		for (t = st0; t <= en0; ++t) {
			int xt1, vt1, x2t1;
			if (t == 0) xt1 = -q - e, x2t1 = -q2 - e2, vt1 = ufirst; /* first column */
			else xt1 = x[t - 1], vt1 = v[t - 1], x2t1 = x2[t - 1];
			int a = target[t], b = query[r - t];
			int z = (a == m - 1 || b == m - 1) ? sc_N : a == b ? sc_mch : sc_mis, d;
			int ut = u[t], A = xt1 + vt1, B = y[t] + ut, A2 = x2t1 + vt1, B2 = y2[t] + ut;
			ORC_CHK(A); ORC_CHK(B); ORC_CHK(A2); ORC_CHK(B2);
			if (!(flag & ORC_EZ_RIGHT)) {
				d = A > z ? 1 : 0;  z = z > A ? z : A;
				d = B > z ? 2 : d;  z = z > B ? z : B;
				d = A2 > z ? 3 : d; z = z > A2 ? z : A2;
				d = B2 > z ? 4 : d; z = z > B2 ? z : B2;
			} else {
				d = z > A ? 0 : 1;  z = z > A ? z : A;
				d = z > B ? d : 2;  z = z > B ? z : B;
				d = z > A2 ? d : 3; z = z > A2 ? z : A2;
				d = z > B2 ? d : 4; z = z > B2 ? z : B2;
			}
			z = z < sc_mch ? z : sc_mch;
			int un = z - vt1, vnew = z - ut;
			A -= z - q, B -= z - q, A2 -= z - q2, B2 -= z - q2;
			ORC_CHK(un); ORC_CHK(vnew); ORC_CHK(A); ORC_CHK(B); ORC_CHK(A2); ORC_CHK(B2);
			int pa, pb, pa2, pb2;
			if (!(flag & ORC_EZ_RIGHT)) pa = A > 0, pb = B > 0, pa2 = A2 > 0, pb2 = B2 > 0;
			else pa = A >= 0, pb = B >= 0, pa2 = A2 >= 0, pb2 = B2 >= 0;
			u[t] = un, vn[t] = vnew;
			xn[t] = (pa ? A : 0) - qe, y[t] = (pb ? B : 0) - qe, x2n[t] = (pa2 ? A2 : 0) - qe2, y2[t] = (pb2 ? B2 : 0) - qe2;
			ORC_CHK(xn[t]); ORC_CHK(y[t]); ORC_CHK(x2n[t]); ORC_CHK(y2[t]);
			p[(size_t)r * tlen + t] = (uint8_t)(d | pa << 3 | pb << 4 | pa2 << 5 | pb2 << 6);
		}
		for (t = st0; t <= en0; ++t) v[t] = vn[t], x[t] = xn[t], x2[t] = x2n[t];
		if (!approx_max) {
			int32_t max_H, max_t;
			if (r > 0) {
				int en1 = st0 + (en0 - st0) / 4 * 4, i;
				int32_t HH[4], tt[4];
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en0;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (max_H > ez->max) ez->max = max_H, ez->max_t = max_t, ez->max_q = r - max_t;
			else if (max_t >= ez->max_t && r - max_t >= ez->max_q) {
				int tl = max_t - ez->max_t, ql = (r - max_t) - ez->max_q, l = tl > ql ? tl - ql : ql - tl;
				if (zdrop >= 0 && ez->max - max_H > zdrop + l * e2) { ez->zdropped = 1; break; }
			}
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else {
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += v[last_H0_t];
				else ++last_H0_t, H0 += u[last_H0_t];
			} else H0 = v[0] - qe, last_H0_t = 0;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
	}
	{
		int i = -1, j = -1, n = 0, state = 0, go = 1;
		if (!ez->zdropped && !(flag & ORC_EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
		else if (!ez->zdropped && (flag & ORC_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) ez->reach_end = 1, i = ez->mqe_t, j = qlen - 1;
		else if (ez->max_t >= 0 && ez->max_q >= 0) i = ez->max_t, j = ez->max_q;
		else go = 0;
		if (go) {
			while (i >= 0 && j >= 0) {
				int tmp = p[(size_t)(i + j) * tlen + i];
				if (state == 0) state = tmp & 7;
				else if (!(tmp >> (state + 2) & 1)) state = 0;
				if (state == 0) state = tmp & 7;
				if (state == 0) orc_push(cigar, &n, 0, 1), --i, --j;
				else if (state == 1 || state == 3) orc_push(cigar, &n, 2, 1), --i;
				else orc_push(cigar, &n, 1, 1), --j;
			}
			if (i >= 0) orc_push(cigar, &n, 2, i + 1);
			if (j >= 0) orc_push(cigar, &n, 1, j + 1);
			if (!(flag & ORC_EZ_REV_CIGAR))
				for (i = 0; i < n >> 1; ++i) { uint32_t c = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = c; }
			ez->n_cigar = n;
		}
	}
	free(u); free(x2n); free(H); free(p);
	return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * mm_sketch: symmetric (w,k)-minimizers.  C/sketch.c:77-143 (hash: :28-38, base codes: :9-26), non-HPC.
 * Restated with an explicit ring of the last w "events" (an event is any position except a skipped palindromic
 * k-mer) so that the emission rules can be read one by one; the CUDA kernel decides the same rules per event in
 * parallel.  Output: x = hash<<8 | span, y = rid<<32 | lastPos<<1 | strand, in emission order.
 * ---------------------------------------------------------------------------------------------------------- */
static uint64_t orc_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}
static int orc_code(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	default: return c < 4 ? c : 4; /* the table also maps bytes 0..3 to themselves */
	}
}
/* returns the number of minimizers; out_x/out_y need room for len entries */
long orc_sketch(const char *str, int len, int w, int k, uint32_t rid, uint64_t *out_x, uint64_t *out_y)
{
	const uint64_t MAXV = ~0ULL, shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1;
	uint64_t kmer[2] = {0, 0}, ring_x[256], ring_y[256], min_x = MAXV, min_y = MAXV;
	int i, j, l = 0, pos = 0, min_pos = 0;
	long n = 0;
	for (j = 0; j < w; ++j) ring_x[j] = ring_y[j] = MAXV;
#define ORC_EMIT(X, Y) (out_x[n] = (X), out_y[n] = (Y), ++n)
	for (i = 0; i < len; ++i) {
		int c = orc_code((unsigned char)str[i]);
		uint64_t ix = MAXV, iy = MAXV;
		if (c < 4) {
			int z, span = l + 1 < k ? l + 1 : k;
			kmer[0] = (kmer[0] << 2 | c) & mask;
			kmer[1] = (kmer[1] >> 2) | (3ULL ^ c) << shift1;
			if (kmer[0] == kmer[1]) continue; /* strand unknown: no event at all (:108) */
			z = kmer[0] < kmer[1] ? 0 : 1;
			if (++l >= k) ix = orc_hash64(kmer[z], mask) << 8 | span, iy = (uint64_t)rid << 32 | (uint32_t)i << 1 | z;
		} else l = 0; /* the k-mer registers are NOT cleared */
		ring_x[pos] = ix, ring_y[pos] = iy;
		if (l == w + k - 1 && min_x != MAXV) { /* first full window: duplicates of the minimum (:116-121) */
			for (j = pos + 1; j < w; ++j) if (min_x == ring_x[j] && ring_y[j] != min_y) ORC_EMIT(ring_x[j], ring_y[j]);
			for (j = 0; j < pos; ++j) if (min_x == ring_x[j] && ring_y[j] != min_y) ORC_EMIT(ring_x[j], ring_y[j]);
		}
		if (ix <= min_x) { /* new minimum displaces the old one (:122-124) */
			if (l >= w + k && min_x != MAXV) ORC_EMIT(min_x, min_y);
			min_x = ix, min_y = iy, min_pos = pos;
		} else if (pos == min_pos) { /* the minimum leaves the window (:125-138) */
			if (l >= w + k - 1 && min_x != MAXV) ORC_EMIT(min_x, min_y);
			for (j = pos + 1, min_x = MAXV; j < w; ++j) if (min_x >= ring_x[j]) min_x = ring_x[j], min_y = ring_y[j], min_pos = j;
			for (j = 0; j <= pos; ++j) if (min_x >= ring_x[j]) min_x = ring_x[j], min_y = ring_y[j], min_pos = j;
			if (l >= w + k - 1 && min_x != MAXV) {
				for (j = pos + 1; j < w; ++j) if (min_x == ring_x[j] && min_y != ring_y[j]) ORC_EMIT(ring_x[j], ring_y[j]);
				for (j = 0; j <= pos; ++j) if (min_x == ring_x[j] && min_y != ring_y[j]) ORC_EMIT(ring_x[j], ring_y[j]);
			}
		}
		if (++pos == w) pos = 0;
	}
	if (min_x != MAXV) ORC_EMIT(min_x, min_y);
	return n;
}

/* ------------------------------------------------------------------------------------------------------------------
 * orc_chain_fill -- the score fill of mg_lchain_rmq (C/lchain.c:276-358) WITHOUT its two balanced trees.
 *
 * What the trees hold at anchor i is a contiguous window of the sorted anchor array: [st, i0) for the range-minimum
 * tree and [st_in, i0) for the near-neighbourhood tree (anchors enter group by group when the target position changes,
 * :280-293, and leave from the front, :295-312).  So the fill can be stated with plain scans over those windows:
 *   - the RMQ (:313-316) is "smallest priority among window members whose query position lies in (y - max_dist, y], at y
 *     itself only the very first anchor (closed upper key (y, 0))";
 *   - the walk (:319-348) visits the members of the inner window with query position in [y - max_dist_inner, y - 1] in
 *     descending (query position, index) order with the reference's skip counter and t[] marks.
 * The one thing a scan cannot know is which of several EQUAL priorities the reference's tree would return (that depends
 * on its rotation history): such an anchor is reported in *n_tied and the caller must not compare beyond it.
 * f, p (-1 = none), v as in the reference; t is scratch of n int32.  undet[i] (optional) is set for every anchor from a
 * non-unique RMQ minimum to the end of its independent segment (the trees are empty wherever the target / strand changes
 * or two neighbours are more than max_dist apart, so nothing carries over).  Returns the index of the first such anchor,
 * or n when there was none.
 * ------------------------------------------------------------------------------------------------------------------ */
static float orc_log2(float x) /* C/mmpriv.h:118-126 */
{
	union { float f; uint32_t i; } z = { x };
	float log_2 = ((z.i >> 23) & 255) - 128;
	z.i &= ~(255 << 23);
	z.i += 127 << 23;
	log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return log_2;
}

static int32_t orc_link(uint64_t xi, uint64_t yi, uint64_t xj, uint64_t yj, float pen_gap, float pen_skip, int *exact, int32_t *width)
{ /* C/lchain.c:232-248 */
	int32_t dq = (int32_t)yi - (int32_t)yj, dr = (int32_t)(xi - xj), dd, dg, q_span, sc;
	*width = dd = dr > dq ? dr - dq : dq - dr;
	dg = dr < dq ? dr : dq;
	q_span = yj >> 32 & 0xff;
	sc = q_span < dg ? q_span : dg;
	if (exact) *exact = (dd == 0 && dg <= q_span);
	if (dd || dq > q_span) {
		float lin_pen = pen_gap * (float)dd + pen_skip * (float)dg;
		float log_pen = dd >= 1 ? orc_log2(dd + 1) : 0.0f;
		sc -= (int)(lin_pen + .5f * log_pen);
	}
	return sc;
}

typedef struct { int32_t y; int64_t i; } orc_key_t;
static int orc_key_desc(const void *a, const void *b)
{
	const orc_key_t *p = (const orc_key_t *)a, *q = (const orc_key_t *)b;
	if (p->y != q->y) return p->y > q->y ? -1 : 1;
	return p->i > q->i ? -1 : p->i < q->i ? 1 : 0;
}

int64_t orc_chain_fill(int64_t n, const uint64_t *xy, int max_dist, int max_dist_inner, int bw, int max_skip, int cap,
                       float pen_gap, float pen_skip, int32_t *f, int32_t *p, int32_t *v, int32_t *t, uint8_t *undet)
{
	int64_t i, j, i0 = 0, st = 0, st_in = 0, first_tie = n;
	int tainted = 0;
	orc_key_t *cand = (orc_key_t *)malloc(sizeof(orc_key_t) * (size_t)(n > 0 ? n : 1));
	if (max_dist < bw) max_dist = bw; /* :262-263 */
	if (max_dist_inner <= 0 || max_dist_inner >= max_dist) max_dist_inner = 0;
	for (i = 0; i < n; ++i) t[i] = 0;
	for (i = 0; i < n; ++i) {
		const uint64_t xi = xy[2 * i], yi = xy[2 * i + 1];
		const int32_t y = (int32_t)yi, q_span = yi >> 32 & 0xff;
		int32_t max_f = q_span;
		int64_t max_j = -1, best = -1, n_best = 0;
		double best_pri = 0;
		if (i0 < i && xy[2 * i0] != xi) i0 = i; /* :280-293: the previous group becomes visible */
		/* :295-312, both windows: too far behind, other target / strand, or more members than the cap */
		while (st < i && (xi >> 32 != xy[2 * st] >> 32 || xi > xy[2 * st] + (uint64_t)max_dist || (i0 > st ? i0 - st : 0) > cap)) ++st;
		if (max_dist_inner > 0)
			while (st_in < i && (xi >> 32 != xy[2 * st_in] >> 32 || xi > xy[2 * st_in] + (uint64_t)max_dist_inner || (i0 > st_in ? i0 - st_in : 0) > cap)) ++st_in;
		/* :313-316 */
		for (j = st; j < i0; ++j) {
			const int32_t yj = (int32_t)xy[2 * j + 1];
			double pri;
			if (!(yj > y - max_dist && (yj < y || (yj == y && j == 0)))) continue;
			pri = -(f[j] + 0.5 * pen_gap * ((int32_t)xy[2 * j] + (int32_t)xy[2 * j + 1])); /* :285 */
			if (best < 0 || pri < best_pri) best = j, best_pri = pri, n_best = 1;
			else if (pri == best_pri) ++n_best;
		}
		if (i > 0 && (xi >> 32 != xy[2 * (i - 1)] >> 32 || xi > xy[2 * (i - 1)] + (uint64_t)max_dist)) tainted = 0; /* a new segment */
		if (n_best > 1) {
			tainted = 1;
			if (first_tie == n) first_tie = i;
		}
		if (undet) undet[i] = (uint8_t)tainted;
		if (best >= 0) {
			int exact;
			int32_t width, n_skip = 0;
			int32_t sc = f[best] + orc_link(xi, yi, xy[2 * best], xy[2 * best + 1], pen_gap, pen_skip, &exact, &width);
			if (width <= bw && sc > max_f) max_f = sc, max_j = best;
			if (!exact && max_dist_inner > 0 && i0 > st_in && y > 0) { /* :319-348 */
				int64_t m = 0, k;
				for (j = st_in; j < i0; ++j) {
					const int32_t yj = (int32_t)xy[2 * j + 1];
					if (yj <= y - 1 && yj >= y - max_dist_inner) cand[m].y = yj, cand[m].i = j, ++m;
				}
				qsort(cand, (size_t)m, sizeof(orc_key_t), orc_key_desc);
				for (k = 0; k < m; ++k) {
					j = cand[k].i;
					sc = f[j] + orc_link(xi, yi, xy[2 * j], xy[2 * j + 1], pen_gap, pen_skip, 0, &width);
					if (width <= bw) {
						if (sc > max_f) {
							max_f = sc, max_j = j;
							if (n_skip > 0) --n_skip;
						} else if (t[j] == (int32_t)i) {
							if (++n_skip > max_skip) break;
						}
						if (p[j] >= 0) t[p[j]] = (int32_t)i;
					}
				}
			}
		}
		f[i] = max_f, p[i] = (int32_t)max_j;
		v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f; /* :354 */
	}
	free(cand);
	return first_tie;
}
