/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the banded affine alignment behind pangraph's `map_variations`
 * (SURVEY 8f-1).  Nothing under pangraph_b200/ may include, link or call this file; tests/ and __graft_entry__.smoke()
 * use it as the checker for the CUDA path (pangraph_b200/csrc/nextalign.cu).
 *
 * Follows, statement by statement (PG = packages/pangraph/src):
 *   simple_stripes            PG/align/nextclade/align/band_2d.rs:36-54
 *   score_matrix              PG/align/nextclade/align/score_matrix.rs:23-200   (i32 scores, i8 paths, one stripe per ref row)
 *   backtrace                 PG/align/nextclade/align/backtrace.rs:17-98
 *   align_nuc_simplestripe    PG/align/nextclade/align/align.rs:33-75          (band doubling while the boundary is hit)
 *   insertions_strip          PG/align/nextclade/align/insertions_strip.rs:48-98
 *   find_nuc_changes          PG/align/nextclade/analyze/nuc_changes.rs:19-70
 *   align_with_nextclade      PG/align/nextclade/align_with_nextclade.rs:24-76 (leading / trailing deletions appended)
 *   map_variations            PG/align/map_variations.rs:39-80                 (Edit: subs, dels, inss; ins position + 1)
 *   Nuc codes and the IUPAC match table: PG/align/nextclade/alphabet/nuc.rs:10-31, align/score_matrix_nuc.rs:7-30
 *     (code + 1 is the set of bases T=1 A=2 C=4 G=8; two non-gap letters match iff the sets intersect)
 * Pinned on the reference's own unit-test vectors (tests/test_oracle_nextalign.py): the four map_variations tests, the four
 * align_with_nextclade tests and the two align_nuc_simplestripe tests.  Parity status: pinned. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NA_MATCH 1
#define NA_REF_GAP_MATRIX 2
#define NA_QRY_GAP_MATRIX 4
#define NA_REF_GAP_EXTEND 8
#define NA_QRY_GAP_EXTEND 16
#define NA_BOUNDARY 32
#define NA_NO_ALIGN (-1000000000)

typedef struct {
	int32_t penalty_gap_extend, penalty_gap_open, penalty_mismatch, score_match;
	int32_t left_terminal_gaps_free, right_terminal_gaps_free, left_align; /* gap_alignment_side: Left = 1 */
	int32_t min_length, max_alignment_attempts;
} orc_na_params_t;

/* NextalignParams::default() with map_variations' overrides (min_length 1, max_alignment_attempts from the build args) */
void orc_na_default_params(orc_na_params_t *p, int max_alignment_attempts)
{
	p->penalty_gap_extend = 0, p->penalty_gap_open = 6, p->penalty_mismatch = 1, p->score_match = 3;
	p->left_terminal_gaps_free = 1, p->right_terminal_gaps_free = 1, p->left_align = 1;
	p->min_length = 1, p->max_alignment_attempts = max_alignment_attempts;
}

/* to_nuc: -1 for a character the reference rejects */
int orc_na_code(char c)
{
	static const char *abc = "TAWCYMHGKRDSBVN-";
	const char *p = c ? strchr(abc, c) : 0;
	return p ? (int)(p - abc) : -1;
}
static int na_match(int x, int y) /* lookup_nuc_scoring_matrix(x, y) > 0 */
{
	if (x == 15 || y == 15) return (x == 15 && y == 15) || x == 14 || y == 14;
	return ((x + 1) & (y + 1)) != 0;
}

typedef struct { int begin, end; } na_stripe_t;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

static void simple_stripes(int mean_shift, int band_width, int ref_len, int qry_len, na_stripe_t *s)
{
	for (int i = 0; i <= ref_len; ++i) {
		s[i].begin = clampi(-mean_shift - band_width + i, 0, qry_len);
		s[i].end = clampi(-mean_shift + band_width + i + 1, 1, qry_len + 1);
	}
	s[0].begin = 0;
	s[ref_len].end = qry_len + 1;
}

/* one attempt: score_matrix + backtrace.  aln_* receive the alignment back to front.  Returns the alignment length,
 * -2 when the backtrace meets a cell with no origin (the reference's unreachable!()), -3 when it leaves the band. */
static int64_t align_pairwise(const uint8_t *q, int qlen, const uint8_t *r, int rlen, const orc_na_params_t *pa, const na_stripe_t *st,
                              char *aln_q, char *aln_r, int *score_out, int *hit_boundary)
{
	static const char *abc = "TAWCYMHGKRDSBVN-";
	const int n_cols = qlen + 1;
	int64_t *row0 = malloc(sizeof(int64_t) * (size_t)(rlen + 2));
	row0[0] = 0;
	for (int i = 0; i <= rlen; ++i) row0[i + 1] = row0[i] + (st[i].end - st[i].begin);
	int32_t *scores = calloc((size_t)row0[rlen + 1] + 1, sizeof(int32_t));
	int8_t *paths = calloc((size_t)row0[rlen + 1] + 1, 1);
	int32_t *qry_gaps = malloc(sizeof(int32_t) * (size_t)n_cols);
#define SC(i, j) scores[row0[i] + ((j) - st[i].begin)]
#define PA(i, j) paths[row0[i] + ((j) - st[i].begin)]
	const int ext = pa->penalty_gap_extend, gopen = pa->penalty_gap_open; /* gap_open_close is flat */
	PA(0, 0) = 0, SC(0, 0) = 0;
	for (int qpos = st[0].begin + 1; qpos < st[0].end; ++qpos) {
		PA(0, qpos) = NA_REF_GAP_EXTEND + NA_REF_GAP_MATRIX;
		if (pa->left_terminal_gaps_free) SC(0, qpos) = 0;
		else if (qpos == 1) SC(0, 1) = -gopen;
		else SC(0, qpos) = SC(0, qpos - 1) - ext;
	}
	for (int j = 0; j < n_cols; ++j) qry_gaps[j] = NA_NO_ALIGN;
	for (int ri = 1; ri <= rlen; ++ri) {
		int32_t ref_gaps = NA_NO_ALIGN;
		for (int qpos = st[ri].begin; qpos < st[ri].end; ++qpos) {
			int tmp_path = 0, origin = 0;
			int32_t score = NA_NO_ALIGN, tmp_score;
			if (qpos == 0) {
				tmp_path = NA_QRY_GAP_EXTEND, origin = NA_QRY_GAP_MATRIX;
				if (pa->left_terminal_gaps_free) score = 0;
				else if (ri == 1) score = -gopen;
				else score = SC(ri - 1, 0) - ext;
			} else {
				if (qpos > st[ri - 1].begin && qpos - 1 < st[ri - 1].end) {
					const int a = q[qpos - 1], b = r[ri - 1];
					if (a == 14 || b == 14) score = SC(ri - 1, qpos - 1) + pa->score_match - 1;
					else if (na_match(a, b)) score = SC(ri - 1, qpos - 1) + pa->score_match;
					else score = SC(ri - 1, qpos - 1) - pa->penalty_mismatch;
					origin = NA_MATCH;
				} else if (ri < rlen && qpos < qlen) tmp_path |= NA_BOUNDARY;
				if (qpos > st[ri].begin) {
					int32_t r_gap_extend, r_gap_open;
					if (ri != rlen || !pa->right_terminal_gaps_free) r_gap_extend = ref_gaps - ext, r_gap_open = SC(ri, qpos - 1) - gopen;
					else r_gap_extend = ref_gaps, r_gap_open = SC(ri, qpos - 1);
					if (r_gap_extend >= r_gap_open && qpos > st[ri].begin + 1) tmp_score = r_gap_extend, tmp_path += NA_REF_GAP_EXTEND;
					else tmp_score = r_gap_open;
					ref_gaps = tmp_score;
					if (score - pa->left_align < tmp_score) score = tmp_score, origin = NA_REF_GAP_MATRIX;
				} else if (ri < rlen && qpos < qlen) tmp_path |= NA_BOUNDARY; /* n_rows - 1 == ref_len */
				if (qpos < st[ri - 1].end) {
					int32_t q_gap_extend, q_gap_open;
					if (qpos != qlen || !pa->right_terminal_gaps_free) q_gap_extend = qry_gaps[qpos] - ext, q_gap_open = SC(ri - 1, qpos) - gopen;
					else q_gap_extend = qry_gaps[qpos], q_gap_open = SC(ri - 1, qpos);
					if (q_gap_extend >= q_gap_open && qpos < st[ri - 2].end) tmp_score = q_gap_extend, tmp_path += NA_QRY_GAP_EXTEND;
					else tmp_score = q_gap_open;
					qry_gaps[qpos] = tmp_score;
					if (score - pa->left_align < tmp_score) score = tmp_score, origin = NA_QRY_GAP_MATRIX;
				} else if (qpos < n_cols - 1 && ri < rlen) {
					qry_gaps[qpos] = NA_NO_ALIGN;
					tmp_path |= NA_BOUNDARY;
				}
			}
			tmp_path += origin;
			PA(ri, qpos) = (int8_t)tmp_path;
			SC(ri, qpos) = score;
		}
	}
	/* backtrace */
	int r_pos = rlen, q_pos = n_cols - 1, current = 0, hb = 0;
	int64_t n = 0;
	while (r_pos > 0 || q_pos > 0) {
		if (q_pos < st[r_pos].begin || q_pos >= st[r_pos].end) { n = -3; break; }
		const int origin = PA(r_pos, q_pos);
		if (origin & NA_BOUNDARY) hb = 1;
		if ((origin & NA_MATCH) && current == 0) {
			--q_pos, --r_pos;
			aln_q[n] = abc[q[q_pos]], aln_r[n] = abc[r[r_pos]], ++n;
		} else if (((origin & NA_REF_GAP_MATRIX) && current == 0) || current == NA_REF_GAP_MATRIX) {
			--q_pos;
			aln_q[n] = abc[q[q_pos]], aln_r[n] = '-', ++n;
			current = (origin & NA_REF_GAP_EXTEND) ? NA_REF_GAP_MATRIX : 0;
		} else if (((origin & NA_QRY_GAP_MATRIX) && current == 0) || current == NA_QRY_GAP_MATRIX) {
			--r_pos;
			aln_q[n] = '-', aln_r[n] = abc[r[r_pos]], ++n;
			current = (origin & NA_QRY_GAP_EXTEND) ? NA_QRY_GAP_MATRIX : 0;
		} else { n = -2; break; }
	}
	*score_out = SC(rlen, n_cols - 1);
	*hit_boundary = hb;
#undef SC
#undef PA
	free(row0), free(scores), free(paths), free(qry_gaps);
	return n;
}

/* align_nuc_simplestripe: aln_qry / aln_ref (capacity qlen + rlen + 1 each) receive the alignment front to back.
 * Returns its length; -1: bad character or query shorter than min_length (the reference returns Err), -2 / -3: see above.
 * *band_width_used, *attempts report the last attempt. */
int64_t orc_align_nuc_simplestripe(const char *qry, int qlen, const char *ref, int rlen, int mean_shift, int band_width,
                                   const orc_na_params_t *pa, char *aln_qry, char *aln_ref, int *score, int *hit_boundary,
                                   int *band_width_used, int *attempts)
{
	if (qlen < pa->min_length) return -1;
	uint8_t *q = malloc((size_t)qlen + 1), *r = malloc((size_t)rlen + 1);
	int64_t n = -1;
	for (int i = 0; i < qlen; ++i) { int c = orc_na_code(qry[i]); if (c < 0) goto done; q[i] = (uint8_t)c; }
	for (int i = 0; i < rlen; ++i) { int c = orc_na_code(ref[i]); if (c < 0) goto done; r[i] = (uint8_t)c; }
	{
		na_stripe_t *st = malloc(sizeof(na_stripe_t) * (size_t)(rlen + 1));
		int bw = band_width, attempt = 1;
		simple_stripes(mean_shift, bw, rlen, qlen, st);
		n = align_pairwise(q, qlen, r, rlen, pa, st, aln_qry, aln_ref, score, hit_boundary);
		while (n >= 0 && *hit_boundary && attempt < pa->max_alignment_attempts) {
			const int ams = mean_shift < 0 ? -mean_shift : mean_shift, alt = ams > 1 ? ams : 1;
			bw = 2 * bw > alt ? 2 * bw : alt;
			simple_stripes(mean_shift, bw, rlen, qlen, st);
			++attempt;
			n = align_pairwise(q, qlen, r, rlen, pa, st, aln_qry, aln_ref, score, hit_boundary);
		}
		*band_width_used = bw, *attempts = attempt;
		free(st);
		for (int64_t i = 0; i < n / 2; ++i) {
			char t = aln_qry[i]; aln_qry[i] = aln_qry[n - 1 - i]; aln_qry[n - 1 - i] = t;
			t = aln_ref[i]; aln_ref[i] = aln_ref[n - 1 - i]; aln_ref[n - 1 - i] = t;
		}
		if (n >= 0) aln_qry[n] = aln_ref[n] = 0;
	}
done:
	free(q), free(r);
	return n;
}

/* map_variations: the Edit of qry against ref.  band_width is the caller's (extra_band_width is added here, like
 * map_variations.rs:51).  Outputs (capacities: subs rlen, dels rlen + 2, inss qlen + 1 records, ins_seq qlen + 1 chars):
 *   sub_pos[], sub_chr[]; del_pos[], del_len[] in the reference's order (inner deletions ascending, then the leading,
 *   then the trailing one); ins_pos[] (already + 1), ins_off[], ins_len[] into ins_seq.
 * Returns 0, or the negative code of orc_align_nuc_simplestripe. */
int orc_map_variations(const char *ref, int rlen, const char *qry, int qlen, int mean_shift, int band_width, int extra_band_width,
                       int max_alignment_attempts, int32_t *n_sub, int32_t *sub_pos, char *sub_chr, int32_t *n_del, int32_t *del_pos,
                       int32_t *del_len, int32_t *n_ins, int32_t *ins_pos, int32_t *ins_off, int32_t *ins_len, char *ins_seq,
                       int *hit_boundary, int *attempts)
{
	orc_na_params_t pa;
	orc_na_default_params(&pa, max_alignment_attempts);
	char *aq = malloc((size_t)qlen + rlen + 2), *ar = malloc((size_t)qlen + rlen + 2);
	int score, bw_used;
	const int64_t n = orc_align_nuc_simplestripe(qry, qlen, ref, rlen, mean_shift, band_width + extra_band_width, &pa, aq, ar, &score,
	                                             hit_boundary, &bw_used, attempts);
	*n_sub = *n_del = *n_ins = 0;
	if (n < 0) { free(aq), free(ar); return (int)n; }
	/* insertions_strip: drop the columns where the reference has a gap, remember what the query had there */
	char *stripped = malloc((size_t)rlen + 1);
	int ref_pos = -1, cur_len = 0, cur_start = -1, ns = 0, ioff = 0;
	for (int64_t i = 0; i < n; ++i) {
		if (ar[i] == '-') {
			if (cur_len == 0) cur_start = ref_pos, ins_off[*n_ins] = ioff;
			ins_seq[ioff++] = aq[i], ++cur_len;
		} else {
			stripped[ns++] = aq[i], ++ref_pos;
			if (cur_len) ins_pos[*n_ins] = cur_start + 1, ins_len[*n_ins] = cur_len, ++*n_ins, cur_len = 0; /* map_variations.rs:73: pos + 1 */
		}
	}
	if (cur_len) ins_pos[*n_ins] = cur_start + 1, ins_len[*n_ins] = cur_len, ++*n_ins;
	/* (insertions.sort() by (pos, len): positions are strictly increasing here, the order is already sorted) */
	/* find_nuc_changes on (stripped query, reference) */
	int64_t n_d = 0, d_pos = -1, a_start = -1, a_end = -1;
	int before = 1;
	for (int i = 0; i < ns; ++i) {
		const char d = stripped[i];
		if (d != '-') {
			if (before) a_start = i, before = 0;
			else if (n_d > 0) del_pos[*n_del] = (int32_t)d_pos, del_len[*n_del] = (int32_t)n_d, ++*n_del, n_d = 0;
			a_end = i + 1;
		}
		if (d != '-' && d != ref[i]) sub_pos[*n_sub] = i, sub_chr[*n_sub] = d, ++*n_sub;
		else if (d == '-' && !before) {
			if (n_d == 0) d_pos = i;
			++n_d;
		}
	}
	/* align_with_nextclade.rs:50-67: terminal gaps become deletions, appended after the sorted inner ones */
	if (a_start >= 0 && a_end >= 0) {
		if (a_start > 0) del_pos[*n_del] = 0, del_len[*n_del] = (int32_t)a_start, ++*n_del;
		if (a_end < rlen) del_pos[*n_del] = (int32_t)a_end, del_len[*n_del] = (int32_t)(rlen - a_end), ++*n_del;
	} else del_pos[*n_del] = 0, del_len[*n_del] = rlen, ++*n_del;
	free(aq), free(ar), free(stripped);
	return 0;
}
