// K5: ksw_extd2 on the GPU -- dual-affine banded DP (Suzuki-Kasahara difference recurrence, int8) with traceback.
//
// Behavioural contract: bit-identical ksw_extz_t and CIGAR to the reference's ksw_extd2_sse
// (packages/minimap2-sys/minimap2/ksw2_extd2_sse.c:34-401, ksw2.h:127-184) for the four flag combinations that occur
// on pangraph's path (align.c:714,755,759,798).  That includes the reference's 16-lane artefacts: on every
// anti-diagonal the padded range [st,en] is evaluated, padded cells read whatever substitution score was last written
// for their target position, and band edges / traceback see those cells.
//
// Mapping: one CTA per DP problem.  All per-target-position state (u,y,y2,s in place; v,x,x2 ping-ponged between
// anti-diagonals so that a single barrier separates them) lives in shared memory -- or in a global scratch slab when a
// problem is too long -- as packed int8x4 words; a thread advances one word (4 cells) per step with the byte-wise SIMD
// integer instructions.  The traceback matrix streams to HBM with 32-bit stores (one row per anti-diagonal); the
// traceback walk runs on thread 0 after the wavefront.  No tensor cores: this is integer add/compare/select work.
#include "ksw_extd2.h"
#include "pgmm_cuda.h"

#include <algorithm>
#include <type_traits>
#include <numeric>

namespace pgmm {

namespace {

constexpr int kMiscWords = 96;  // per-CTA scratch: warp partial keys (64 words) + scalars

__host__ __device__ inline int round16_up(int x) { return (x + 15) / 16 * 16; }

struct Geom {
  int T, Qp, n_col16, n_row;
  size_t state_bytes, p_bytes;
};

__host__ __device__ inline Geom make_geom(int qlen, int tlen, int w, int flag) {
  Geom g;
  if (w < 0) w = tlen > qlen ? tlen : qlen;
  g.T = round16_up(tlen);
  g.Qp = (qlen + 3) / 4 * 4 + 24;  // 4 leading + >=16 trailing zero bytes around the reversed query
  int n_col = qlen < tlen ? qlen : tlen;
  n_col = ((n_col < w + 1 ? n_col : w + 1) + 15) / 16 + 1;  // ksw2_extd2_sse.c:89-91
  g.n_col16 = n_col * 16;
  g.n_row = qlen + tlen - 1;
  g.p_bytes = (size_t)g.n_row * g.n_col16;
  g.state_bytes = (size_t)g.T * 11 + ((flag & KSW_APPROX_MAX) ? 0 : (size_t)g.T * 4) + g.Qp;
  return g;
}

__device__ __forceinline__ uint32_t rep4(int v) { return (uint32_t)(uint8_t)v * 0x01010101u; }
__device__ __forceinline__ uint32_t sel4(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }
// 0xff in byte k iff lo <= t0+k <= hi
__device__ __forceinline__ uint32_t range_mask(int t0, int lo, int hi) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t0 + k >= lo && t0 + k <= hi) m |= 0xffu << (8 * k);
  return m;
}
__device__ __forceinline__ int8_t byte_of(uint32_t w, int k) { return (int8_t)(w >> (8 * k)); }
__device__ __forceinline__ uint32_t set_byte(uint32_t w, int k, int v) {
  return (w & ~(0xffu << (8 * k))) | ((uint32_t)(uint8_t)v << (8 * k));
}

// band of anti-diagonal r (ksw2_extd2_sse.c:137-147); returns false when the band is empty
__device__ __forceinline__ bool band(int r, int qlen, int tlen, int w, int &st0, int &en0) {
  int st = 0, en = tlen - 1;
  if (st < r - qlen + 1) st = r - qlen + 1;
  if (en > r) en = r;
  if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
  if (en > (r + w) >> 1) en = (r + w) >> 1;
  st0 = st, en0 = en;
  return st <= en;
}

// GPU-side time span of a wave: slot 1 of the wave's counter block = earliest CTA start, slot 2 = latest CTA end
// (%globaltimer, ns); slots 3 / 4 are stamped by marker kernels on the wave's main stream (PGMM_TRACE only)
__device__ __forceinline__ void span_begin(unsigned long long *counter) {
  if (threadIdx.x == 0) atomicMin(counter + 1, trace::now());
}
__device__ __forceinline__ void span_end(unsigned long long *counter) { atomicMax(counter + 2, trace::now()); }
__global__ void stamp_kernel(unsigned long long *slot) { *slot = trace::now(); }

struct EzState {
  int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end;
};

template <int NT>
__device__ __forceinline__ void cta_bar() {
  if (NT == 32) __syncwarp();
  else __syncthreads();
}

// Traceback (ksw2.h:127-159, ksw2_extd2_sse.c:388-399), CIGAR packing and -- for first-pass fills -- the mm_test_zdrop
// scan, run by one thread once the wavefront of a problem is done.  TQ8 = target, QR8[i] = query[qlen-1-i].
__device__ __noinline__ void finish_job(const KswJob &job, int jid, EzState ez, int n_diag, int w, bool abs_layout, int stride,
                                        const uint8_t *__restrict__ P, const uint8_t *TQ8, const uint8_t *QR8, const KswScoring &sc,
                                        uint32_t *__restrict__ cig_arena, uint32_t *__restrict__ cig_packed,
                                        unsigned long long *__restrict__ cig_counter, KswOut *__restrict__ outs, int n_walked = -1) {
  const int qlen = job.qlen, tlen = job.tlen, flag = job.flag;
  {
    int i = -1, j = -1, n = 0, state = 0;
    bool go = n_walked < 0;  // a warp-cooperative walk (walk_warp) has already left its n_walked runs in the arena
    if (!go) n = n_walked;
    if (!ez.zdropped && !(flag & KSW_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
    else if (!ez.zdropped && (flag & KSW_EXTZ_ONLY) && ez.mqe + job.end_bonus > ez.max) ez.reach_end = 1, i = ez.mqe_t, j = qlen - 1;
    else if (ez.max_t >= 0 && ez.max_q >= 0) i = ez.max_t, j = ez.max_q;
    else go = false;
    uint32_t *cig = cig_arena + job.cig_off;
    if (go && n_walked < 0) {
      uint32_t last = 0;  // run being built: len<<4|op, flushed when the op changes
      while (i >= 0 && j >= 0) {
        const int r = i + j;
        int st0, en0;
        band(r, qlen, tlen, w, st0, en0);
        // rows hold the padded range [off, off_end] of their anti-diagonal (the reference's layout) or, for the
        // kernels that only evaluate real cells, all target positions from 0
        const int off = abs_layout ? 0 : st0 / 16 * 16, off_end = abs_layout ? tlen - 1 : (en0 + 16) / 16 * 16 - 1;
        int force = -1;
        if (i < off) force = 2;
        if (i > off_end) force = 1;
        const int tmp = force < 0 ? __ldcg(P + (size_t)r * stride + (i - off)) : 0;
        if (state == 0) state = tmp & 7;
        else if (!((tmp >> (state + 2)) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force >= 0) state = force;
        uint32_t op;
        if (state == 0) op = 0, --i, --j;
        else if (state == 1 || state == 3) op = 2, --i;
        else op = 1, --j;
        if (last != 0 && (last & 0xf) == op) last += 1u << 4;
        else {
          if (last != 0) cig[n++] = last;
          last = 1u << 4 | op;
        }
      }
      if (i >= 0) {  // leading deletion
        if (last != 0 && (last & 0xf) == 2) last += (uint32_t)(i + 1) << 4;
        else {
          if (last != 0) cig[n++] = last;
          last = (uint32_t)(i + 1) << 4 | 2;
        }
      }
      if (j >= 0) {  // leading insertion
        if (last != 0 && (last & 0xf) == 1) last += (uint32_t)(j + 1) << 4;
        else {
          if (last != 0) cig[n++] = last;
          last = (uint32_t)(j + 1) << 4 | 1;
        }
      }
      if (last != 0) cig[n++] = last;
    }
    // reserve a slot in the packed output and move the run there (the walk produced it back to front)
    const unsigned long long pos = n > 0 ? atomicAdd(cig_counter, (unsigned long long)n) : 0ull;
    uint32_t *dst = cig_packed + pos;
    if (flag & KSW_REV_CIGAR)
      for (int k = 0; k < n; ++k) dst[k] = cig[k];
    else
      for (int k = 0; k < n; ++k) dst[k] = cig[n - 1 - k];
    KswOut o;
    o.cig_pos = (uint32_t)pos;
    o.zd_max = 0, o.zd_t0 = o.zd_t1 = o.zd_q0 = o.zd_q1 = -1;
    if ((flag & KSW_APPROX_MAX) && n > 0) {
      // first-pass fill: re-score the path like mm_test_zdrop (align.c:33-68) while the sequences are still in shared
      // memory, so that the host only has to look at five numbers per fill
      const int amb = sc.sc_ambi < 0 ? sc.sc_ambi : -sc.sc_ambi;
      int32_t score = 0, mx = INT32_MIN, mx_i = -1, mx_j = -1, ti = 0, qj = 0, zd = 0;
      int p00 = -1, p01 = -1, p10 = -1, p11 = -1;
      for (int k = 0; k < n; ++k) {
        const uint32_t c = dst[k], op = c & 0xf, len = c >> 4;
        if (op == 0) {
          for (uint32_t l = 0; l < len; ++l) {
            const int ct = TQ8[ti + l], cq = QR8[qlen - 1 - (qj + (int)l)];
            score += (ct > 3 || cq > 3) ? amb : ct == cq ? sc.sc_mch : sc.sc_mis;
            if (score < mx) {
              const int li = ti + (int)l - mx_i, lj = qj + (int)l - mx_j, diff = li > lj ? li - lj : lj - li;
              const int z = mx - score - diff * sc.e;
              if (z > zd) zd = z, p00 = mx_i, p01 = ti + (int)l, p10 = mx_j, p11 = qj + (int)l;
            } else mx = score, mx_i = ti + (int)l, mx_j = qj + (int)l;
          }
          ti += len, qj += len;
        } else {
          score -= sc.q + sc.e * (int)len;
          if (op == 1) qj += len;
          else ti += len;
          if (score < mx) {
            const int li = ti - mx_i, lj = qj - mx_j, diff = li > lj ? li - lj : lj - li;
            const int z = mx - score - diff * sc.e;
            if (z > zd) zd = z, p00 = mx_i, p01 = ti, p10 = mx_j, p11 = qj;
          } else mx = score, mx_i = ti, mx_j = qj;
        }
      }
      o.zd_max = zd, o.zd_t0 = p00, o.zd_t1 = p01, o.zd_q0 = p10, o.zd_q1 = p11;
    }
    o.max = ez.max, o.zdropped = ez.zdropped, o.max_q = ez.max_q, o.max_t = ez.max_t, o.mqe = ez.mqe, o.mqe_t = ez.mqe_t;
    o.mte = ez.mte, o.mte_q = ez.mte_q, o.score = ez.score, o.reach_end = ez.reach_end, o.n_cigar = n;
    o.n_diag = n_diag;
    outs[jid] = o;
  }
}

// The traceback walk of finish_job for a problem whose rows hold all target positions (K5a / K5b layout), run by a whole
// warp: from a cell reached in state 0 the 32 lanes read the next 32 cells of the DIAGONAL at once and the walk takes all
// leading ones whose direction is 0 in one step (at 1 % divergence that is almost the whole path: ~n/32 + #gaps dependent
// loads instead of n).  Cells inside a gap are walked one at a time (every lane reads the same byte).  Runs are written by
// lane 0, back to front like the serial walk; returns their number.  Same rules as ksw2.h:127-159.
__device__ __forceinline__ int walk_warp(int i, int j, int stride, const uint8_t *__restrict__ P, uint32_t *__restrict__ cig, int lane) {
  int n = 0, state = 0;
  uint32_t last = 0;
  const auto push = [&](uint32_t op, uint32_t len) {
    if (last != 0 && (last & 0xf) == op) last += len << 4;
    else {
      if (last != 0 && lane == 0) cig[n] = last;
      n += last != 0;
      last = len << 4 | op;
    }
  };
  while (i >= 0 && j >= 0) {
    if (state == 0) {
      const bool valid = i - lane >= 0 && j - lane >= 0;
      const int tmp = valid ? __ldcg(P + (size_t)(i + j - 2 * lane) * stride + (i - lane)) : 7;
      const unsigned stop = __ballot_sync(0xffffffffu, (tmp & 7) != 0);
      const int run = stop ? __ffs(stop) - 1 : 32;  // leading diagonal cells
      if (run > 0) push(0, (uint32_t)run), i -= run, j -= run;
      if (run == 32 || i < 0 || j < 0) continue;
      state = __shfl_sync(0xffffffffu, tmp, run) & 7;  // the cell that leaves the diagonal: state was 0, so it takes its direction
    } else {
      const int tmp = __ldcg(P + (size_t)(i + j) * stride + i);
      if (!((tmp >> (state + 2)) & 1)) state = 0;
      if (state == 0) state = tmp & 7;
      if (state == 0) {  // back on the diagonal: this cell is a match, continue with the wide reads
        push(0, 1), --i, --j;
        continue;
      }
    }
    if (state == 1 || state == 3) push(2, 1), --i;
    else push(1, 1), --j;
  }
  if (i >= 0) push(2, (uint32_t)(i + 1));  // leading deletion
  if (j >= 0) push(1, (uint32_t)(j + 1));  // leading insertion
  if (last != 0) {
    if (lane == 0) cig[n] = last;
    ++n;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------------------------
// The cell-pair step shared by K5a and K5b: two neighbouring target positions per 32-bit register, 16 bits each.
//
// Representation (chosen so that one anti-diagonal step of a pair costs ~30 integer instructions, all of them either
// DPX three-operand instructions -- VIADDMNMX.U16x2 = max(a + b, c) per lane -- or plain 32-bit IADD3 / LOP3 / IMAD):
//   * every value is scaled by 8 and stored in offset binary: half = 8*value + 0x4000.  All halves are positive and
//     differences of two stay positive after the constant is restored, so a packed SUBTRACT is one 32-bit IADD3 with no
//     borrow between the halves (there is no 16x2 subtract instruction; the negation costs two more);
//   * the three spare low bits carry the traceback direction through the maximum: the score carries tag 7, x+v (dir 1)
//     6, y+u (2) 5, x2+v (3) 4, y2+u (4) 3, so among equal scores the unsigned max keeps the earlier candidate exactly
//     like the reference's strict compares (ksw2_extd2_sse.c:240-249), and dir = ~max & 7.  x, y, x2, y2 keep their tag
//     in the state, u, v and z are tag-free;
//   * the continuation flags a > 0 ... (:262-269) come out of carry bits: for a half h = 8a + 0x4000 + tag,
//     bit 15 of h + 0x3ff8, bit 14 of h - 8, bit 13 of h - 0x2008, bit 12 of h - 0x3008 are set iff a >= 1
//     (|a| <= 128 under the reference's (q+e)+(q2+e2) <= 127 precondition), one add and one mask-or per flag.
// The same arithmetic as a numpy model, checked against the scalar recurrence: tests/test_lane_model.py.
// Cells outside the matrix: below / right of it they are cells of the matrix extended with base 'A' (legitimate DP
// values, same bounds); above the first row they hold unbounded garbage, but such a cell is never the LOW half of a pair
// whose high half is real, and carries only travel upwards.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kLB = 0x4000u;  // offset of a half
__host__ __device__ __forceinline__ uint32_t enc1(int v) { return (uint32_t)(8 * v + (int)kLB) & 0xffffu; }
__host__ __device__ __forceinline__ uint32_t rep2(int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; }
__device__ __forceinline__ int dec_lo(uint32_t w) { return ((int)(w & 0xffffu) - (int)kLB) >> 3; }
__device__ __forceinline__ int dec_hi(uint32_t w) { return ((int)(w >> 16) - (int)kLB) >> 3; }

struct LaneConsts {  // computed on the host, passed as a kernel parameter: the loop reads them as constant-bank operands
  uint32_t MCH7, SCN7, DM8, CQ, CQ2, NQE, NQE2, FLX, FLY, FLX2, FLY2;
  uint32_t X0, X20;          // x[-1], x2[-1] of the first column (low half, with tag)
  uint32_t Y0, Y20, V0;      // first-row y, y2 (with tag) and the initial v as single halves
  uint32_t UF0, UF1, UF2, UF3;  // u of the first row / v of the first column: r == 0, r < long_thres, r == long_thres, beyond
  int32_t q, e, q2, e2, long_thres;  // gap costs ordered like the reference orders them (:50-52)
  __host__ __device__ void init(const KswScoring &sc) {
    q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
    if (q2 + e2 < q + e) {
      int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t;
    }
    const int long_thres0 = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
    long_thres = (q2 + e2 + long_thres0 * e2 > q + e + long_thres0 * e) ? long_thres0 + 1 : long_thres0;
    const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
    UF0 = enc1(-q - e), UF1 = enc1(-e), UF2 = enc1(long_diff), UF3 = enc1(-e2);
    const int scN = sc.sc_ambi == 0 ? -e2 : -(sc.sc_ambi < 0 ? -sc.sc_ambi : sc.sc_ambi);
    MCH7 = rep2(8 * sc.sc_mch + 0x8000 + 7), SCN7 = rep2(8 * scN + 0x8000 + 7), DM8 = (uint32_t)(8 * (sc.sc_mch - sc.sc_mis));
    CQ = rep2(8 * q + (int)kLB), CQ2 = rep2(8 * q2 + (int)kLB);
    NQE = rep2(-8 * (q + e)), NQE2 = rep2(-8 * (q2 + e2));
    FLX = rep2((int)kLB + 6 - 8 * (q + e)), FLY = rep2((int)kLB + 5 - 8 * (q + e));
    FLX2 = rep2((int)kLB + 4 - 8 * (q2 + e2)), FLY2 = rep2((int)kLB + 3 - 8 * (q2 + e2));
    X0 = enc1(-q - e) + 6, X20 = enc1(-q2 - e2) + 4, Y0 = enc1(-q - e) + 5, Y20 = enc1(-q2 - e2) + 3, V0 = enc1(-q - e);
  }
  __device__ __forceinline__ uint32_t ufirst(int r) const { return r == 0 ? UF0 : r < long_thres ? UF1 : r == long_thres ? UF2 : UF3; }
};

// One anti-diagonal step of a pair.  In: x, v, x2 of the cells to the left (previous anti-diagonal), u, y, y2 of the
// pair itself, the two target and the two query codes.  Out: the new state and the two traceback bytes (low byte =
// low cell).  HAS_N: whether the two windows hold any ambiguous base (uniform per problem; the loops exist in both versions).
template <bool HAS_N>
__device__ __forceinline__ uint32_t pair_step(const LaneConsts &c, uint32_t XT1, uint32_t VT1, uint32_t X2T1, uint32_t Uo,
                                              uint32_t Yo, uint32_t Y2o, uint32_t tq, uint32_t qq, uint32_t &Un, uint32_t &Vn,
                                              uint32_t &Xn, uint32_t &Yn, uint32_t &X2n, uint32_t &Y2n) {
  const uint32_t ne = __vminu2(tq ^ qq, 0x00010001u);
  uint32_t Z = c.MCH7 - ne * c.DM8;  // halves stay positive: exact per half
  if (HAS_N) {
    const uint32_t isn = __vminu2((tq | qq) & 0x00040004u, 0x00010001u) * 0xffffu;
    Z = (Z & ~isn) | (c.SCN7 & isn);
  }
  uint32_t M = __viaddmax_u16x2(XT1, VT1, Z);
  M = __viaddmax_u16x2(Yo, Uo, M);
  M = __viaddmax_u16x2(X2T1, VT1, M);
  M = __viaddmax_u16x2(Y2o, Uo, M);
  const uint32_t Z8 = __vminu2(M, c.MCH7) & 0xfff8fff8u;  // z, offset 0x8000
  Un = Z8 - VT1, Vn = Z8 - Uo;                            // offset back to 0x4000
  const uint32_t A = XT1 - Un + c.CQ, B = Yo - Vn + c.CQ, A2 = X2T1 - Un + c.CQ2, B2 = Y2o - Vn + c.CQ2;
  // plain 32-bit adds (they can go to the FMA pipe; the kernel is bound by the ALU pipe): the three negative constants make
  // the low half carry into the high one, always, which lifts the high half by one -- below the slack of the thresholds
  // (the tags are at most 6, so 8a + tag + 1 >= 8 still means a >= 1)
  uint32_t acc = (A + 0xcff8cff8u) & 0x10001000u;   // - 0x3008
  acc |= (B + 0xdff8dff8u) & 0x20002000u;           // - 0x2008
  acc |= (A2 + 0xfff8fff8u) & 0x40004000u;          // - 8
  acc |= (B2 + 0x3ff83ff8u) & 0x80008000u;          // + 0x3ff8: no carry out of a half
  // max(a, 0) - (q + e): signed lanes here (the addend is negative, all three operands are far from the 16-bit limits)
  Xn = __viaddmax_s16x2(A, c.NQE, c.FLX), Yn = __viaddmax_s16x2(B, c.NQE, c.FLY);
  X2n = __viaddmax_s16x2(A2, c.NQE2, c.FLX2), Y2n = __viaddmax_s16x2(B2, c.NQE2, c.FLY2);
  const uint32_t D = (acc >> 9) | (~M & 0x00070007u);
  return __byte_perm(D, 0u, 0x4420);
}

// ---------------------------------------------------------------------------------------------------------------
// K5a: first-pass gap fills (flag == KSW_APPROX_MAX, band never binding, target window <= 256 bases) -- about 85 % of
// all DP cells of a round (~28 000 windows of ~205 x 205 per 5-Mbp pair).  Specialised because, when the band cannot
// bind, padded cells can never reach a real cell (oracle/pgmm_oracle.c::orc_ksw_extd2_unbanded states and tests this):
// only real cells are evaluated, and since real cells never leave the signed-byte range (the reference's
// (q+e)+(q2+e2) <= 127 precondition) the recurrence runs in 16-bit lanes, two cells per 32-bit register, on the native
// VIADD.16x2 / VIMNMX.S16x2 instructions (the byte-wide __v*4 intrinsics are emulated with 5-6 instructions each).
//
// One warp per problem, all state in registers: pair p = 32*slot + lane holds cells t = 2p, 2p+1 (4 slots = 256 cells).
// Slots that do not intersect the active range of an anti-diagonal are skipped; the left neighbour of a pair comes
// from the previous lane by shuffle.  Traceback bytes go to HBM (row = anti-diagonal, column = target position).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFillWarps = 4;     // problems per CTA
constexpr int kFillMaxT = 256;    // target window limit (4 slots x 32 lanes x 2 cells)
constexpr int kFillMaxQ = 1024;   // query window limit (two 16-bit copies of the reversed query in shared memory)


__global__ void __launch_bounds__(kFillWarps * 32, 7) ksw_fill_small_kernel(const KswJob *__restrict__ jobs, const int *__restrict__ job_ids,
                                                                         int n_jobs, const uint8_t *__restrict__ qcodes,
                                                                         const uint8_t *__restrict__ tcodes, KswScoring sc,
                                                                         const __grid_constant__ LaneConsts lc, int q_cap,
                                                                         uint8_t *__restrict__ p_arena, uint32_t *__restrict__ cig_arena,
                                                                         KswOut *__restrict__ outs, uint32_t *__restrict__ cig_packed,
                                                                         unsigned long long *__restrict__ cig_counter) {
  extern __shared__ uint32_t dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot_id = blockIdx.x * kFillWarps + warp;
  span_begin(cig_counter);
  if (slot_id >= n_jobs) return;  // whole warp leaves together
  const int jid = job_ids[slot_id];
  const KswJob job = jobs[jid];
  const unsigned long long tr0 = trace::begin();
  const int qlen = job.qlen, tlen = job.tlen;
  // per-warp shared memory: target bytes [kFillMaxT+16], reversed query bytes [4 + q_cap + 28], and the reversed query
  // as 16-bit codes twice (second copy shifted by one element) so that the two query bases of a pair are always one
  // aligned 32-bit load; both copies have 2*kFillMaxT zero elements in front and behind
  const int qh_len = q_cap + 4 * kFillMaxT;  // halfwords per copy
  const int per_warp_bytes = (kFillMaxT + 16) + (q_cap + 32) + 2 * qh_len * 2;
  uint8_t *wbase = (uint8_t *)dyn_smem + (size_t)warp * ((per_warp_bytes + 15) / 16 * 16);
  uint8_t *TQ8 = wbase;
  uint8_t *QRraw = TQ8 + kFillMaxT + 16;  // QR8 = QRraw + 4
  uint16_t *QH0 = (uint16_t *)(QRraw + q_cap + 32), *QH1 = QH0 + qh_len;
  uint8_t *QR8 = QRraw + 4;
  for (int i = lane; i < (kFillMaxT + 16) / 4; i += 32) ((uint32_t *)TQ8)[i] = 0;
  for (int i = lane; i < (q_cap + 32) / 4; i += 32) ((uint32_t *)QRraw)[i] = 0;
  for (int i = lane; i < qh_len; i += 32) ((uint32_t *)QH0)[i] = 0;  // both copies (2*qh_len halfwords)
  __syncwarp();
  uint32_t any_n = 0;
  {
    const uint8_t *tb = tcodes + job.t_off, *qb = qcodes + job.q_off;
    for (int i = lane; i < tlen; i += 32) {
      const uint8_t c = tb[i];
      TQ8[i] = c, any_n |= c;
    }
    for (int i = lane; i < qlen; i += 32) {
      const uint8_t c = qb[qlen - 1 - i];  // reversed query: element i pairs with target t on anti-diagonal r when i = qlen-1-r+t
      QR8[i] = c, any_n |= c;
      QH0[2 * kFillMaxT + i] = c;
      QH1[2 * kFillMaxT + i - 1] = c;  // QH1[k] = QH0[k+1]
    }
  }
  const bool has_n = __any_sync(0xffffffffu, (any_n & 4u) != 0);  // ambiguous bases anywhere in the two windows
  __syncwarp();

  const int Tp = (tlen + 15) / 16 * 16, n_row = qlen + tlen - 1;
  uint8_t *P = p_arena + job.p_off;

  uint32_t U[4], Y[4], Y2[4], V[4], X[4], X2[4], TQ[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    U[k] = V[k] = lc.V0 * 0x00010001u, X[k] = lc.X0 * 0x00010001u, Y[k] = lc.Y0 * 0x00010001u;
    X2[k] = lc.X20 * 0x00010001u, Y2[k] = lc.Y20 * 0x00010001u;
    const int t = 2 * (32 * k + lane);
    TQ[k] = (uint32_t)TQ8[t] | (uint32_t)TQ8[t + 1] << 16;
  }
  int32_t H0 = 0;
  uint32_t vsum = 0;
  const int k_end = (tlen - 1) >> 6, sh_end = ((tlen - 1) & 1) * 16;  // slot and half of the last target position
  // per-lane addresses, advanced incrementally: the query word of (slot 0, this lane) in either copy (shared-space
  // addresses: no generic-address arithmetic in the loop) and this lane's two traceback bytes of the current row
  const uint32_t qa0 = (uint32_t)__cvta_generic_to_shared(QH0) + 4u * lane, qa1 = (uint32_t)__cvta_generic_to_shared(QH1) + 4u * lane;
  uint8_t *prow_l = P + 2 * lane;
  const bool last_lane = lane == 31;

  // the loop exists twice: with and without the ambiguous-base score (decided once per problem)
  const auto main_loop = [&](auto has_n_tag) {
  constexpr bool HAS_N = decltype(has_n_tag)::value;
  for (int r = 0; r < n_row; ++r, prow_l += Tp) {
    const int st0 = r - qlen + 1 > 0 ? r - qlen + 1 : 0, en0 = r < tlen - 1 ? r : tlen - 1;
    const uint32_t UF = lc.ufirst(r);
    const int p_lo = st0 >> 1, p_hi = en0 >> 1;  // active pairs
    const int qbase = 2 * kFillMaxT + (qlen - 1 - r);  // halfword index of the query base of target position 0
    // the copy in which (t=0, t=1) is an aligned 32-bit word: copy 1 is copy 0 shifted by one element
    uint32_t qaddr = ((qbase & 1) ? qa1 : qa0) + 4u * (uint32_t)(qbase >> 1);
    int fr_k = en0 == r ? r >> 6 : -1, fr_lane = (r >> 1) & 31;  // slot and lane of the first-row cell, if any
    // opaque to the optimiser: left alone it re-derives all three in every slot (the shared-window base from %cluster_ctaid,
    // the first-row test from r) -- 20 of ~80 instructions per slot
    asm volatile("" : "+r"(qaddr), "+r"(fr_k), "+r"(fr_lane));
#pragma unroll
    for (int k = 3; k >= 0; --k) {
      if (32 * k > p_hi || 32 * k + 31 < p_lo) continue;  // warp-uniform
      // previous anti-diagonal's x, v, x2 of the cell to the left: high half of the previous pair.  Lane 31 offers the
      // last pair of the previous slot (read by lane 0); in slot 0, lane 0 reads the first-column values (:156-159).
      uint32_t sx = X[k], sv = V[k], sx2 = X2[k];
      if (k > 0) {
        if (last_lane) sx = X[k - 1], sv = V[k - 1], sx2 = X2[k - 1];
      } else if (last_lane) sx = lc.X0 << 16, sv = UF << 16, sx2 = lc.X20 << 16;
      const uint32_t px = __shfl_sync(0xffffffffu, sx, (lane + 31) & 31), pv = __shfl_sync(0xffffffffu, sv, (lane + 31) & 31),
                     px2 = __shfl_sync(0xffffffffu, sx2, (lane + 31) & 31);
      const uint32_t XT1 = __funnelshift_l(px, X[k], 16), VT1 = __funnelshift_l(pv, V[k], 16), X2T1 = __funnelshift_l(px2, X2[k], 16);
      uint32_t Uo = U[k], Yo = Y[k], Y2o = Y2[k];
      if (fr_k == k) {  // first row (:160-163): warp-uniform test, one lane acts
        if (fr_lane == lane) {
          if (r & 1) Uo = (Uo & 0x0000ffffu) | UF << 16, Yo = (Yo & 0x0000ffffu) | lc.Y0 << 16, Y2o = (Y2o & 0x0000ffffu) | lc.Y20 << 16;
          else Uo = (Uo & 0xffff0000u) | UF, Yo = (Yo & 0xffff0000u) | lc.Y0, Y2o = (Y2o & 0xffff0000u) | lc.Y20;
        }
      }
      uint32_t qq;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(qq) : "r"(qaddr + 128u * k));
      const uint32_t d16 = pair_step<HAS_N>(lc, XT1, VT1, X2T1, Uo, Yo, Y2o, TQ[k], qq, U[k], V[k], X[k], Y[k], X2[k], Y2[k]);
      if (32 * k + lane <= p_hi) *(uint16_t *)(prow_l + 64 * k) = (uint16_t)d16;  // cells past the target end share the padded row
    }
    // The score (:367-379).  The reference walks a data-dependent staircase from (0,0) and adds v[last] or u[last+1] on
    // every anti-diagonal; u and v are exact differences of one score matrix, so the sum does not depend on the path
    // (tests/test_lane_model.py checks the identity): here the lane that owns the last target position adds up its v
    // on the anti-diagonals that reach it, which needs no communication; the boundary H(tlen-1, -1) is added at the end.
    if (en0 == tlen - 1) {
      const uint32_t w = k_end == 0 ? V[0] : k_end == 1 ? V[1] : k_end == 2 ? V[2] : V[3];
      vsum += (w >> sh_end) & 0xffffu;
    }
  }
  };
  if (has_n) main_loop(std::true_type{});
  else main_loop(std::false_type{});
  {
    const int gap1 = lc.q + lc.e * tlen, gap2 = lc.q2 + lc.e2 * tlen;
    const uint32_t tot = __shfl_sync(0xffffffffu, vsum, ((tlen - 1) >> 1) & 31);  // qlen halves of 8*v + 0x4000 each
    H0 = ((int32_t)(tot - (uint32_t)qlen * kLB) >> 3) - (gap1 < gap2 ? gap1 : gap2);
  }

  __threadfence_block();  // every lane's traceback bytes must be visible to the lanes that walk them
  __syncwarp();
  const unsigned long long tr1 = trace::begin();
  const int n_runs = walk_warp(tlen - 1, qlen - 1, Tp, P, cig_arena + job.cig_off, lane);  // first-pass fill: from the last cell
  __syncwarp();
  if (lane == 0) {
    EzState ez;
    ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
    ez.max = 0, ez.mqe = ez.mte = KSW_NEG_INF;
    ez.zdropped = 0, ez.reach_end = 0;
    ez.score = H0;  // the last anti-diagonal is the single cell (tlen-1, qlen-1)
    finish_job(job, jid, ez, n_row, /*w=*/tlen > qlen ? tlen : qlen, /*abs_layout=*/true, Tp, P, TQ8, QR8, sc, cig_arena, cig_packed, cig_counter, outs, n_runs);
    if (warp == 0) trace::emit(1, tr0, tr1, (unsigned)n_row);
    span_end(cig_counter);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K5b: wide problems whose band cannot bind -- the long gap fills a chain makes across an inversion or a big indel
// (up to ~10k x 10k), first pass (approximate max) and second pass (exact max with z-drop).  Same recurrence and lane
// format as K5a, one CTA per problem: pair p = slot*NT + thread, KP slots per thread in registers.  A pair's left
// neighbour comes by shuffle; across warp boundaries it comes from a small shared-memory mailbox that the last lane
// of every warp fills at the end of an anti-diagonal (double-buffered by parity), so ONE block barrier per
// anti-diagonal is enough.  All threads keep the scalar bookkeeping (tracked score, running maximum, z-drop state)
// redundantly from values published through that mailbox, so no thread ever waits for a "leader".
// ---------------------------------------------------------------------------------------------------------------
struct WideMail {      // per parity
  uint32_t x[32][8], v[32][8], x2[32][8];  // [warp][slot]: new x, v, x2 of the warp's last pair
  int32_t hhi[32][8];                      // and the exact-mode H of its high cell
  long long wbest[32];                     // per-warp best (H, rank) key
  int32_t d0, d1;                          // v[last], u[last+1] for the tracked score
  int32_t hen, hst0;                       // H[en0], H[st0] of this anti-diagonal
};

template <int NW, int KP, bool EXACT>
__global__ void __launch_bounds__(NW * 32) ksw_fill_wide_kernel(const KswJob *__restrict__ jobs, const int *__restrict__ job_ids,
                                                                const uint8_t *__restrict__ qcodes, const uint8_t *__restrict__ tcodes,
                                                                KswScoring sc, const __grid_constant__ LaneConsts lc,
                                                                uint8_t *__restrict__ p_arena,
                                                                uint32_t *__restrict__ cig_arena, KswOut *__restrict__ outs,
                                                                uint32_t *__restrict__ cig_packed, unsigned long long *__restrict__ cig_counter) {
  constexpr int NT = NW * 32;
  extern __shared__ uint32_t dyn_smem[];
  __shared__ WideMail mail[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jid = job_ids[blockIdx.x];
  const KswJob job = jobs[jid];
  const unsigned long long tr0 = trace::begin();
  span_begin(cig_counter);
  const int qlen = job.qlen, tlen = job.tlen;
  const int Tp = (tlen + 15) / 16 * 16, n_row = qlen + tlen - 1, qe = lc.q + lc.e, e2 = lc.e2;
  // shared memory: target bytes [Tp + 16], reversed query bytes [4 + qlen + pad], reversed query as 16-bit codes with
  // Tp + 16 zero elements in front and behind
  uint8_t *TQ8 = (uint8_t *)dyn_smem;
  const int qr_bytes = (qlen + 4 + 28 + 15) / 16 * 16;
  uint8_t *QRraw = TQ8 + Tp + 16, *QR8 = QRraw + 4;
  uint16_t *QH = (uint16_t *)(QRraw + qr_bytes);
  const int qh_pad = Tp + 16, qh_len = qlen + 2 * qh_pad;
  for (int i = tid; i < (Tp + 16) / 4; i += NT) ((uint32_t *)TQ8)[i] = 0;
  for (int i = tid; i < qr_bytes / 4; i += NT) ((uint32_t *)QRraw)[i] = 0;
  for (int i = tid; i < (qh_len + 1) / 2; i += NT) ((uint32_t *)QH)[i] = 0;
  __syncthreads();
  uint32_t any_n = 0;
  {
    const uint8_t *tb = tcodes + job.t_off, *qb = qcodes + job.q_off;
    for (int i = tid; i < tlen; i += NT) {
      const uint8_t c = tb[i];
      TQ8[i] = c, any_n |= c;
    }
    for (int i = tid; i < qlen; i += NT) {
      const uint8_t c = qb[qlen - 1 - i];
      QR8[i] = c, any_n |= c;
      QH[qh_pad + i] = c;
    }
  }
  const bool has_n = __syncthreads_or((any_n & 4u) != 0) != 0;  // ambiguous bases anywhere in the two windows

  uint8_t *P = p_arena + job.p_off;

  uint32_t U[KP], Y[KP], Y2[KP], V[KP], X[KP], X2[KP], TQ[KP];
  int32_t HL[KP], HH[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    U[k] = V[k] = lc.V0 * 0x00010001u, X[k] = lc.X0 * 0x00010001u, Y[k] = lc.Y0 * 0x00010001u;
    X2[k] = lc.X20 * 0x00010001u, Y2[k] = lc.Y20 * 0x00010001u;
    HL[k] = HH[k] = KSW_NEG_INF;
    const int t = 2 * (NT * k + tid);
    TQ[k] = t + 1 < Tp + 16 ? ((uint32_t)TQ8[t] | (uint32_t)TQ8[t + 1] << 16) : 0u;
  }
  EzState ez;
  ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
  ez.max = 0, ez.score = ez.mqe = ez.mte = KSW_NEG_INF;
  ez.zdropped = 0, ez.reach_end = 0;
  // first pass: the score is the sum of v along the last target column (see K5a); its owner adds them up
  uint32_t vsum = 0;
  const int p_end = (tlen - 1) >> 1, sh_end = ((tlen - 1) & 1) * 16;
  int r_done = 0;

  for (int r = 0; r < n_row; ++r) {
    r_done = r;
    const int par = r & 1;
    WideMail &in = mail[par ^ 1], &out = mail[par];
    const int st0 = r - qlen + 1 > 0 ? r - qlen + 1 : 0, en0 = r < tlen - 1 ? r : tlen - 1;
    const uint32_t UF = lc.ufirst(r);
    const int p_lo = st0 >> 1, p_hi = en0 >> 1;
    const int qbase = qh_pad + (qlen - 1 - r);  // halfword index of the query base that meets target position 0
    uint8_t *prow = P + (size_t)r * Tp;
    const int en1 = st0 + (en0 - st0) / 4 * 4, NB = (en1 - st0) >> 2;
    long long best = LLONG_MIN;
#pragma unroll
    for (int k = KP - 1; k >= 0; --k) {
      if (NT * k > p_hi || NT * k + NT - 1 < p_lo) continue;  // block-uniform
      const int p = NT * k + tid;
      uint32_t px = __shfl_up_sync(0xffffffffu, X[k], 1), pv = __shfl_up_sync(0xffffffffu, V[k], 1), px2 = __shfl_up_sync(0xffffffffu, X2[k], 1);
      int32_t ph_old = EXACT ? __shfl_up_sync(0xffffffffu, HH[k], 1) : 0;
      if (lane == 0) {  // previous warp's last pair, or the last pair of the previous slot, as of the previous anti-diagonal
        if (warp > 0) px = in.x[warp - 1][k], pv = in.v[warp - 1][k], px2 = in.x2[warp - 1][k], ph_old = EXACT ? in.hhi[warp - 1][k] : 0;
        else if (k > 0) px = in.x[NW - 1][k - 1], pv = in.v[NW - 1][k - 1], px2 = in.x2[NW - 1][k - 1], ph_old = EXACT ? in.hhi[NW - 1][k - 1] : 0;
        else px = lc.X0 << 16, pv = UF << 16, px2 = lc.X20 << 16;  // first column (:156-159)
      }
      const uint32_t XT1 = __funnelshift_l(px, X[k], 16), VT1 = __funnelshift_l(pv, V[k], 16), X2T1 = __funnelshift_l(px2, X2[k], 16);
      uint32_t Uo = U[k], Yo = Y[k], Y2o = Y2[k];
      if (en0 == r && (r >> 1) == p) {  // first row (:160-163)
        if (r & 1) Uo = (Uo & 0x0000ffffu) | UF << 16, Yo = (Yo & 0x0000ffffu) | lc.Y0 << 16, Y2o = (Y2o & 0x0000ffffu) | lc.Y20 << 16;
        else Uo = (Uo & 0xffff0000u) | UF, Yo = (Yo & 0xffff0000u) | lc.Y0, Y2o = (Y2o & 0xffff0000u) | lc.Y20;
      }
      const int qi = qbase + 2 * p;
      const uint32_t qq = qi + 1 < qh_len ? ((uint32_t)QH[qi] | (uint32_t)QH[qi + 1] << 16) : 0u;
      uint32_t Un, Vn;
      const uint32_t d16 = has_n ? pair_step<true>(lc, XT1, VT1, X2T1, Uo, Yo, Y2o, TQ[k], qq, Un, Vn, X[k], Y[k], X2[k], Y2[k])
                                 : pair_step<false>(lc, XT1, VT1, X2T1, Uo, Yo, Y2o, TQ[k], qq, Un, Vn, X[k], Y[k], X2[k], Y2[k]);
      U[k] = Un, V[k] = Vn;
      if (p <= p_hi) *(uint16_t *)(prow + 2 * p) = (uint16_t)d16;
      const int t_lo = 2 * p, t_hi = 2 * p + 1;
      if (EXACT) {  // H[t] += v[t] for st0 <= t < en0; H[en0] = H[en0-1] (previous anti-diagonal) + u[en0]   (:333-357)
        const int32_t hl_old = HL[k];
        if (r == 0) {
          if (p == 0) HL[k] = dec_lo(Vn) - qe;
        } else {
          if (t_lo >= st0 && t_lo < en0) HL[k] += dec_lo(Vn);
          else if (t_lo == en0) HL[k] = en0 > 0 ? ph_old + dec_lo(Un) : HL[k] + dec_lo(Vn);
          if (t_hi >= st0 && t_hi < en0) HH[k] += dec_hi(Vn);
          else if (t_hi == en0) HH[k] = hl_old + dec_hi(Un);
        }
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {
          const int t = hsel ? t_hi : t_lo;
          if (t >= st0 && t <= en0) {
            const int32_t h = hsel ? HH[k] : HL[k];
            const int d = t - st0;
            const int rank = t == en0 ? 0 : t < en1 ? 1 + (d & 3) * NB + (d >> 2) : 1 + 4 * NB + (t - en1);
            const long long key = (long long)h * 4294967296LL + (long long)(0x7fffffff - rank);
            best = best > key ? best : key;
            if (t == en0) out.hen = h;
            if (t == st0) out.hst0 = h;
          }
        }
      } else if (p == p_end && en0 == tlen - 1) vsum += (Vn >> sh_end) & 0xffffu;
      if (lane == 31) {
        out.x[warp][k] = X[k], out.v[warp][k] = V[k], out.x2[warp][k] = X2[k];
        if (EXACT) out.hhi[warp][k] = HH[k];
      }
    }
    if (EXACT) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = best > other ? best : other;
      }
      if (lane == 0) out.wbest[warp] = best;
    }
    __syncthreads();
    // ---- scalar bookkeeping, done by every thread from the mailbox ----
    if (EXACT) {
      long long b = lane < NW ? out.wbest[lane] : LLONG_MIN;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = b > other ? b : other;
      }
      const int rank = 0x7fffffff - (int)(uint32_t)(b & 0xffffffffLL);
      const int32_t max_H = (int32_t)(b >> 32);
      int32_t max_t;
      if (rank == 0) max_t = en0;
      else if (rank - 1 < 4 * NB) max_t = st0 + 4 * ((rank - 1) % NB) + (rank - 1) / NB;
      else max_t = en1 + (rank - 1 - 4 * NB);
      const int32_t hen = out.hen, hst0 = out.hst0;
      if (en0 == tlen - 1 && hen > ez.mte) ez.mte = hen, ez.mte_q = r - en0;
      if (r - st0 == qlen - 1 && hst0 > ez.mqe) ez.mqe = hst0, ez.mqe_t = st0;
      bool brk = false;
      if (max_H > ez.max) ez.max = max_H, ez.max_t = max_t, ez.max_q = r - max_t;
      else if (max_t >= ez.max_t && r - max_t >= ez.max_q) {
        const int tl = max_t - ez.max_t, ql = (r - max_t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
        if (job.zdrop >= 0 && ez.max - max_H > job.zdrop + l * e2) ez.zdropped = 1, brk = true;
      }
      if (!brk && r == n_row - 1 && en0 == tlen - 1) ez.score = hen;
      if (brk) break;
    }
  }
  if (!EXACT) {
    if ((tid % NT) == (p_end % NT)) {  // qlen halves of 8*v + 0x4000 each, plus the boundary H(tlen-1, -1)
      const int gap1 = lc.q + lc.e * tlen, gap2 = lc.q2 + lc.e2 * tlen;
      mail[0].d0 = ((int32_t)(vsum - (uint32_t)qlen * kLB) >> 3) - (gap1 < gap2 ? gap1 : gap2);
    }
  }
  __threadfence_block();
  __syncthreads();
  if (warp == 0) {
    const unsigned long long tr1 = trace::begin();
    // where the traceback starts (ksw2_extd2_sse.c:388-399 for the two flag sets of a fill: no KSW_EZ_EXTZ_ONLY)
    int wi = -1, wj = -1;
    if (!ez.zdropped) wi = tlen - 1, wj = qlen - 1;
    else if (ez.max_t >= 0 && ez.max_q >= 0) wi = ez.max_t, wj = ez.max_q;
    const int n_runs = wi >= 0 ? walk_warp(wi, wj, Tp, P, cig_arena + job.cig_off, lane) : 0;
    __syncwarp();
    if (tid == 0) {
      if (!EXACT) ez.score = mail[0].d0;
      finish_job(job, jid, ez, r_done + 1, tlen > qlen ? tlen : qlen, /*abs_layout=*/true, Tp, P, TQ8, QR8, sc, cig_arena, cig_packed,
                 cig_counter, outs, n_runs);
      trace::emit(EXACT ? 3 : 2, tr0, tr1, (unsigned)(r_done + 1));
      span_end(cig_counter);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(NT) ksw_extd2_kernel(const KswJob *__restrict__ jobs, const int *__restrict__ job_ids,
                                                       const uint8_t *__restrict__ qcodes,
                                                       const uint8_t *__restrict__ tcodes, KswScoring sc,
                                                       uint8_t *__restrict__ p_arena, uint32_t *__restrict__ cig_arena,
                                                       uint8_t *__restrict__ scratch, KswOut *__restrict__ outs,
                                                       uint32_t *__restrict__ cig_packed, unsigned long long *__restrict__ cig_counter) {
  extern __shared__ uint32_t dyn_smem[];
  __shared__ __align__(16) uint32_t misc[kMiscWords];
  const int tid = threadIdx.x;
  const int jid = job_ids[blockIdx.x];  // index within this wave
  const KswJob job = jobs[jid];
  const unsigned long long tr0 = trace::begin();
  span_begin(cig_counter);
  const int qlen = job.qlen, tlen = job.tlen, flag = job.flag;
  int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
  if (q2 + e2 < q + e) {  // ksw2_extd2_sse.c:73
    int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t;
  }
  const int w = job.w < 0 ? (tlen > qlen ? tlen : qlen) : job.w;
  const bool approx = flag & KSW_APPROX_MAX, right = flag & KSW_RIGHT;
  const Geom g = make_geom(qlen, tlen, w, flag);
  const int T = g.T;

  // ---- carve the per-problem state ----
  uint8_t *base = job.scr_off == ~0ull ? (uint8_t *)dyn_smem : scratch + job.scr_off;
  uint32_t *U32 = (uint32_t *)base, *Y32 = U32 + T / 4, *Y232 = Y32 + T / 4, *S32 = Y232 + T / 4;
  // v, x, x2 are read at t-1 by the neighbouring cell: two copies, swapped every anti-diagonal
  uint32_t *Vc32 = S32 + T / 4, *Vn32 = S32 + 2 * (T / 4);
  uint32_t *Xc32 = S32 + 3 * (T / 4), *Xn32 = S32 + 4 * (T / 4);
  uint32_t *X2c32 = S32 + 5 * (T / 4), *X2n32 = S32 + 6 * (T / 4);
  uint32_t *TQ32 = S32 + 7 * (T / 4);
  uint32_t *QR32 = TQ32 + T / 4;
  int32_t *H = (int32_t *)(QR32 + g.Qp / 4);
  int8_t *U8 = (int8_t *)U32;

  const uint32_t NQE = rep4(-q - e), NQE2 = rep4(-q2 - e2);
  for (int i = tid; i < T / 4; i += NT) {
    U32[i] = NQE, Y32[i] = NQE, Y232[i] = NQE2, S32[i] = 0;
    Vc32[i] = Vn32[i] = NQE;
    Xc32[i] = Xn32[i] = NQE;
    X2c32[i] = X2n32[i] = NQE2;
    TQ32[i] = 0;
    if (!approx) H[4 * i] = H[4 * i + 1] = H[4 * i + 2] = H[4 * i + 3] = KSW_NEG_INF;
  }
  for (int i = tid; i < g.Qp / 4; i += NT) QR32[i] = 0;
  cta_bar<NT>();
  {
    const bool rev = flag & KSW_JOB_REVSEQ;
    const uint8_t *tb = tcodes + job.t_off, *qb = qcodes + job.q_off;
    uint8_t *TQ8 = (uint8_t *)TQ32, *QR8 = (uint8_t *)QR32 + 4;
    for (int i = tid; i < tlen; i += NT) TQ8[i] = rev ? tb[tlen - 1 - i] : tb[i];
    for (int i = tid; i < qlen; i += NT) QR8[i] = rev ? qb[i] : qb[qlen - 1 - i];  // qr[i] = query[qlen-1-i]
  }
  cta_bar<NT>();

  const int long_thres0 = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;  // :99-102
  const int long_thres = (q2 + e2 + long_thres0 * e2 > q + e + long_thres0 * e) ? long_thres0 + 1 : long_thres0;
  const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
  const uint32_t MCH = rep4(sc.sc_mch), MIS = rep4(sc.sc_mis);
  const uint32_t SCN = rep4(sc.sc_ambi == 0 ? -e2 : -(sc.sc_ambi < 0 ? -sc.sc_ambi : sc.sc_ambi));
  const uint32_t Q4 = rep4(q), Q24 = rep4(q2), QE4 = rep4(q + e), QE24 = rep4(q2 + e2), N4 = rep4(4);
  const int qe = q + e;
  uint8_t *P = p_arena + job.p_off;
  const int stride = g.n_col16, n_row = g.n_row;

  EzState ez;
  ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
  ez.max = 0, ez.score = ez.mqe = ez.mte = KSW_NEG_INF;
  ez.zdropped = 0, ez.reach_end = 0;
  int32_t H0 = 0, last_H0_t = 0;  // thread 0 only
  int last_st = -1, last_en = -1;
  long long *wpart = (long long *)misc;     // [NT/32] warp partial keys
  int32_t *hsave = (int32_t *)&misc[64];    // H[en0-1] of the previous anti-diagonal
  volatile int32_t *stop = (int32_t *)&misc[65];
  if (tid == 0) *stop = 0;

  int r_done = 0;
  for (int r = 0; r < n_row; ++r) {
    int st0, en0;
    r_done = r;
    if (!band(r, qlen, tlen, w, st0, en0)) {  // :142-145
      ez.zdropped = 1;
      break;
    }
    const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
    const int8_t *Xc8 = (const int8_t *)Xc32, *Vc8 = (const int8_t *)Vc32, *X2c8 = (const int8_t *)X2c32;
    // stand-ins for position st-1 of the previous anti-diagonal, :149-159
    const int ufirst = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
    int x1, x21, v1;
    if (st > 0) {
      if (st - 1 >= last_st && st - 1 <= last_en) x1 = Xc8[st - 1], x21 = X2c8[st - 1], v1 = Vc8[st - 1];
      else x1 = -q - e, x21 = -q2 - e2, v1 = -q - e;
    } else x1 = -q - e, x21 = -q2 - e2, v1 = ufirst;
    // substitution scores are (re)written for [st0, s_hi] in runs of 16 from st0, :165-180
    const int s_hi = min(st0 + ((en0 - st0) / 16 + 1) * 16 - 1, T - 1);
    const int whi = max(en, s_hi) >> 2;
    const int en1 = st0 + (en0 - st0) / 4 * 4, NB = (en1 - st0) >> 2;
    long long best = LLONG_MIN;
    uint8_t *prow = P + (size_t)r * stride;

    for (int wi = (st >> 2) + tid; wi <= whi; wi += NT) {
      const int t0 = wi << 2;
      uint32_t S = S32[wi];
      if (t0 + 3 >= st0 && t0 <= s_hi) {
        const uint32_t tw = TQ32[wi];
        const int i = qlen - 1 - r + t0 + 4;  // byte index into QR8-4 (>= 1 because t0 >= st0-3)
        const uint32_t lo = QR32[i >> 2], hi = QR32[(i >> 2) + 1];
        const uint32_t qw = __funnelshift_r(lo, hi, (i & 3) * 8);
        const uint32_t eq = __vcmpeq4(tw, qw), nm = __vcmpeq4(tw, N4) | __vcmpeq4(qw, N4);
        const uint32_t scw = sel4(nm, SCN, sel4(eq, MCH, MIS));
        S = sel4(range_mask(t0, st0, s_hi), scw, S);
        S32[wi] = S;
      }
      if (t0 > en) continue;
      uint32_t Uo = U32[wi], Yo = Y32[wi], Y2o = Y232[wi];
      const uint32_t Vc = Vc32[wi], Xc = Xc32[wi], X2c = X2c32[wi];
      int xl, vl, x2l;
      if (t0 == st) xl = x1, vl = v1, x2l = x21;
      else xl = Xc8[t0 - 1], vl = Vc8[t0 - 1], x2l = X2c8[t0 - 1];
      const uint32_t XT1 = (Xc << 8) | (uint8_t)xl, VT1 = (Vc << 8) | (uint8_t)vl, X2T1 = (X2c << 8) | (uint8_t)x2l;
      if (en >= r && (r >> 2) == wi) {  // first row of the matrix, :160-163
        const int k = r & 3;
        Yo = set_byte(Yo, k, -q - e), Y2o = set_byte(Y2o, k, -q2 - e2), Uo = set_byte(Uo, k, ufirst);
      }
      uint32_t Z = S, A = __vadd4(XT1, VT1), B = __vadd4(Yo, Uo), A2 = __vadd4(X2T1, VT1), B2 = __vadd4(Y2o, Uo), D, m;
      if (!right) {  // :228-275
        m = __vcmpgts4(A, Z);  D = m & 0x01010101u;             Z = __vmaxs4(Z, A);
        m = __vcmpgts4(B, Z);  D = sel4(m, 0x02020202u, D);      Z = __vmaxs4(Z, B);
        m = __vcmpgts4(A2, Z); D = sel4(m, 0x03030303u, D);      Z = __vmaxs4(Z, A2);
        m = __vcmpgts4(B2, Z); D = sel4(m, 0x04040404u, D);      Z = __vmaxs4(Z, B2);
      } else {  // :276-322
        m = __vcmpgts4(Z, A);  D = ~m & 0x01010101u;             Z = __vmaxs4(Z, A);
        m = __vcmpgts4(Z, B);  D = sel4(m, D, 0x02020202u);      Z = __vmaxs4(Z, B);
        m = __vcmpgts4(Z, A2); D = sel4(m, D, 0x03030303u);      Z = __vmaxs4(Z, A2);
        m = __vcmpgts4(Z, B2); D = sel4(m, D, 0x04040404u);      Z = __vmaxs4(Z, B2);
      }
      Z = __vmins4(Z, MCH);
      const uint32_t Un = __vsub4(Z, VT1), Vn = __vsub4(Z, Uo);
      uint32_t tmp = __vsub4(Z, Q4);
      A = __vsub4(A, tmp), B = __vsub4(B, tmp);
      tmp = __vsub4(Z, Q24);
      A2 = __vsub4(A2, tmp), B2 = __vsub4(B2, tmp);
      uint32_t Xn, Yn, X2n, Y2n;
      if (!right) {
        m = __vcmpgts4(A, 0);  Xn = __vsub4(A & m, QE4);    D |= m & 0x08080808u;
        m = __vcmpgts4(B, 0);  Yn = __vsub4(B & m, QE4);    D |= m & 0x10101010u;
        m = __vcmpgts4(A2, 0); X2n = __vsub4(A2 & m, QE24); D |= m & 0x20202020u;
        m = __vcmpgts4(B2, 0); Y2n = __vsub4(B2 & m, QE24); D |= m & 0x40404040u;
      } else {
        m = ~__vcmpgts4(0, A);  Xn = __vsub4(A & m, QE4);    D |= m & 0x08080808u;
        m = ~__vcmpgts4(0, B);  Yn = __vsub4(B & m, QE4);    D |= m & 0x10101010u;
        m = ~__vcmpgts4(0, A2); X2n = __vsub4(A2 & m, QE24); D |= m & 0x20202020u;
        m = ~__vcmpgts4(0, B2); Y2n = __vsub4(B2 & m, QE24); D |= m & 0x40404040u;
      }
      U32[wi] = Un, Y32[wi] = Yn, Y232[wi] = Y2n;
      Vn32[wi] = Vn, Xn32[wi] = Xn, X2n32[wi] = X2n;
      *(uint32_t *)(prow + (t0 - st)) = D;
      if (!approx && r > 0) {  // H[t] += v[t] for st0 <= t < en0, :333-357
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int t = t0 + k;
          if (t >= st0 && t < en0) {
            int32_t h = H[t];
            if (t == en0 - 1) *hsave = h;
            h += byte_of(Vn, k);
            H[t] = h;
            // tie order of the reference: en0 first, then 4 interleaved lanes (first block wins inside a lane, lower
            // lane wins across lanes), then the scalar tail in increasing t
            const int d = t - st0;
            const int rank = t < en1 ? 1 + (d & 3) * NB + (d >> 2) : 1 + 4 * NB + (t - en1);
            const long long key = (long long)h * 4294967296LL + (long long)(0x7fffffff - rank);
            best = best > key ? best : key;
          }
        }
      }
    }
    if (!approx) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = best > other ? best : other;
      }
      if ((tid & 31) == 0) wpart[tid >> 5] = best;
    }
    cta_bar<NT>();
    if (tid == 0) {
      const int8_t *Vn8 = (const int8_t *)Vn32;
      if (!approx) {  // :323-366
        int32_t max_H, max_t;
        if (r > 0) {
          int32_t hen;
          if (en0 > 0) hen = (en0 - 1 >= st0 ? *hsave : H[en0 - 1]) + U8[en0];
          else hen = H[0] + Vn8[0];
          H[en0] = hen;
          max_H = hen, max_t = en0;
          long long b = wpart[0];
          for (int k = 1; k < NT / 32; ++k) b = b > wpart[k] ? b : wpart[k];
          if (b != LLONG_MIN && (int32_t)(b >> 32) > max_H) {
            const int rank = 0x7fffffff - (int)(uint32_t)b;
            max_H = (int32_t)(b >> 32);
            if (rank - 1 < 4 * NB) max_t = st0 + 4 * ((rank - 1) % NB) + (rank - 1) / NB;
            else max_t = en1 + (rank - 1 - 4 * NB);
          }
        } else H[0] = Vn8[0] - qe, max_H = H[0], max_t = 0;
        if (en0 == tlen - 1 && H[en0] > ez.mte) ez.mte = H[en0], ez.mte_q = r - en0;
        if (r - st0 == qlen - 1 && H[st0] > ez.mqe) ez.mqe = H[st0], ez.mqe_t = st0;
        bool brk = false;  // ksw_apply_zdrop with e2, ksw2.h:168-184
        if (max_H > ez.max) ez.max = max_H, ez.max_t = max_t, ez.max_q = r - max_t;
        else if (max_t >= ez.max_t && r - max_t >= ez.max_q) {
          const int tl = max_t - ez.max_t, ql = (r - max_t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
          if (job.zdrop >= 0 && ez.max - max_H > job.zdrop + l * e2) ez.zdropped = 1, brk = true;
        }
        if (!brk && r == n_row - 1 && en0 == tlen - 1) ez.score = H[tlen - 1];
        if (brk) *stop = 1;
      } else {  // one tracked cell, :367-383
        if (r > 0) {
          const bool in0 = last_H0_t >= st0 && last_H0_t <= en0, in1 = last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0;
          if (in0 && in1) {
            const int d0 = Vn8[last_H0_t], d1 = U8[last_H0_t + 1];
            if (d0 > d1) H0 += d0;
            else H0 += d1, ++last_H0_t;
          } else if (in0) H0 += Vn8[last_H0_t];
          else {
            ++last_H0_t;
            H0 += last_H0_t < T ? U8[last_H0_t] : 0;
          }
        } else H0 = Vn8[0] - qe, last_H0_t = 0;
        if (r == n_row - 1 && en0 == tlen - 1) ez.score = H0;
      }
    }
    cta_bar<NT>();
    if (*stop) break;
    last_st = st, last_en = en;
    uint32_t *sw;
    sw = Vc32, Vc32 = Vn32, Vn32 = sw;
    sw = Xc32, Xc32 = Xn32, Xn32 = sw;
    sw = X2c32, X2c32 = X2n32, X2n32 = sw;
  }

  if (tid == 0) {
    const unsigned long long tr1 = trace::begin();
    finish_job(job, jid, ez, r_done + 1, w, /*abs_layout=*/false, stride, P, (const uint8_t *)TQ32, (const uint8_t *)QR32 + 4, sc,
               cig_arena, cig_packed, cig_counter, outs);
    trace::emit(4, tr0, tr1, (unsigned)(r_done + 1));
    span_end(cig_counter);
  }
}

constexpr size_t kSmemMax = 200 * 1024;

template <int NT>
void launch_class(const std::vector<int> &ids, size_t smem, const int *d_ids_base, size_t ids_off, const KswJob *d_jobs,
                  const uint8_t *d_q, const uint8_t *d_t, const KswScoring &sc, uint8_t *p_arena, uint32_t *cig_arena,
                  uint8_t *scratch, KswOut *d_outs, uint32_t *cig_packed, unsigned long long *cig_counter, cudaStream_t stream) {
  if (ids.empty()) return;
  static bool attr_set = false;
  if (!attr_set) {
    PGMM_CUDA(cudaFuncSetAttribute(ksw_extd2_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    attr_set = true;
  }
  ksw_extd2_kernel<NT><<<(unsigned)ids.size(), NT, smem, stream>>>(d_jobs, d_ids_base + ids_off, d_q, d_t, sc, p_arena,
                                                                   cig_arena, scratch, d_outs, cig_packed, cig_counter);
  PGMM_CUDA(cudaGetLastError());
}

}  // namespace

// first-pass gap fills whose band cannot bind and whose target window fits one warp's registers: K5a
static inline bool is_small_fill(const KswJob &j) {
  static const bool off = getenv("PGMM_NO_FILL_KERNEL") != nullptr;
  if (off) return false;
  const int mx = j.qlen > j.tlen ? j.qlen : j.tlen;
  return j.flag == KSW_APPROX_MAX && (j.w < 0 || j.w >= mx) && j.tlen <= kFillMaxT && j.qlen <= kFillMaxQ && j.qlen > 0 && j.tlen > 0;
}

// wide problems whose band cannot bind: K5b.  Returns the width configuration (0..3) or -1.
static inline int wide_fill_cfg(const KswJob &j) {
  static const bool off = getenv("PGMM_NO_FILL_KERNEL") != nullptr;
  if (off || j.qlen <= 0 || j.tlen <= 0) return -1;
  if (j.flag != KSW_APPROX_MAX && j.flag != 0) return -1;
  const int mx = j.qlen > j.tlen ? j.qlen : j.tlen;
  if (!(j.w < 0 || j.w >= mx)) return -1;
  const size_t Tp = (size_t)(j.tlen + 15) / 16 * 16;
  const size_t smem = (Tp + 16) + (size_t)(j.qlen + 4 + 28 + 15) / 16 * 16 + 2 * ((size_t)j.qlen + 2 * (Tp + 16)) + 64;
  if (smem > 180 * 1024) return -1;
  return j.tlen <= 256 ? 0 : j.tlen <= 1024 ? 1 : j.tlen <= 4096 ? 2 : j.tlen <= 8192 ? 3 : -1;
}
static inline size_t wide_fill_smem(const KswJob &j) {
  const size_t Tp = (size_t)(j.tlen + 15) / 16 * 16;
  return (Tp + 16) + (size_t)(j.qlen + 4 + 28 + 15) / 16 * 16 + 2 * ((size_t)j.qlen + 2 * (Tp + 16)) + 64;
}

KswGeom ksw_geometry(int qlen, int tlen, int w, int flag) {
  Geom g = make_geom(qlen, tlen, w, flag);
  KswGeom o;
  o.T = g.T, o.n_col16 = g.n_col16, o.p_bytes = (int64_t)g.p_bytes, o.state_bytes = (int64_t)g.state_bytes;
  return o;
}

struct KswEngine::Impl {
  static constexpr int kClasses = 29;  // 5 CTA widths x 4 state-size tiers, K5a (20), K5b: 4 widths x {approximate, exact} (21..28)
  cudaStream_t cls_stream[kClasses] = {};
  int n_streams = kClasses, prio_lo = 0, prio_hi = 0;  // priorities: numerically lower = more urgent
  // PGMM_CLASS_STREAMS=k: the launch classes of an engine share k streams instead of one each.  The few long,
  // latency-bound problems (wide CTAs, long fills) decide when a wave ends: their CTAs go first; the thousands of small
  // fills soak up whatever is left.
  cudaStream_t stream_of(int c) {
    const int k = c % n_streams;
    if (!cls_stream[k]) {
      const bool small = (c == 20 || c < 4) && n_streams == kClasses;
      PGMM_CUDA(cudaStreamCreateWithPriority(&cls_stream[k], cudaStreamNonBlocking, small ? prio_lo : prio_hi));
    }
    return cls_stream[k];
  }
  cudaEvent_t cls_done[kClasses] = {}, fork = nullptr;
  DevBuf<KswJob> d_jobs;
  DevBuf<int> d_ids;
  DevBuf<KswOut> d_outs;
  DevBuf<uint8_t> p_arena, scratch;
  DevBuf<uint32_t> cig_arena, cig_packed;
  DevBuf<unsigned long long> d_counter;
  // pinned staging: copies in both directions are asynchronous and need no driver-side bounce buffer
  PinBuf<KswJob> h_jobs;
  PinBuf<int> h_ids;
  PinBuf<KswOut> h_outs;
  PinBuf<uint32_t> h_cigar;
  PinBuf<unsigned long long> h_counter;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t cls_t0[kClasses] = {}, cls_t1[kClasses] = {};  // timing of each class launch on its own stream
};

KswEngine::KswEngine() : impl_(new Impl) {
  PGMM_CUDA(cudaEventCreate(&impl_->ev0));
  PGMM_CUDA(cudaEventCreate(&impl_->ev1));
  PGMM_CUDA(cudaEventCreateWithFlags(&impl_->fork, cudaEventDisableTiming));
  // class streams are created on first use (stream_of): PGMM_NO_CLASS_STREAMS engines never create any
  static const int n_cls_streams = getenv("PGMM_CLASS_STREAMS") ? std::max(1, std::min((int)Impl::kClasses, atoi(getenv("PGMM_CLASS_STREAMS")))) : (int)Impl::kClasses;
  impl_->n_streams = n_cls_streams;
  PGMM_CUDA(cudaDeviceGetStreamPriorityRange(&impl_->prio_lo, &impl_->prio_hi));
  for (int c = 0; c < Impl::kClasses; ++c) {
    PGMM_CUDA(cudaEventCreateWithFlags(&impl_->cls_done[c], cudaEventDisableTiming | cudaEventBlockingSync));
    PGMM_CUDA(cudaEventCreate(&impl_->cls_t0[c]));
    PGMM_CUDA(cudaEventCreate(&impl_->cls_t1[c]));
  }
}
KswEngine::~KswEngine() {
  cudaEventDestroy(impl_->ev0);
  cudaEventDestroy(impl_->ev1);
  cudaEventDestroy(impl_->fork);
  for (int c = 0; c < Impl::kClasses; ++c) {
    if (impl_->cls_stream[c]) cudaStreamDestroy(impl_->cls_stream[c]);
    cudaEventDestroy(impl_->cls_done[c]), cudaEventDestroy(impl_->cls_t0[c]), cudaEventDestroy(impl_->cls_t1[c]);
  }
  delete impl_;
}

void KswEngine::run(std::vector<KswJob> &jobs, const uint8_t *d_q, const uint8_t *d_t, const KswScoring &sc,
                    KswBatchResult &res, cudaStream_t stream) {
  const size_t n = jobs.size();
  res.out.assign(n, KswOut{});
  res.cig_start.assign(n + 1, 0);
  res.cigar.clear();
  res.cells = 0, res.launches = 0, res.kernel_ms = 0.f;
  for (int f = 0; f < 3; ++f) res.fam_ms[f] = 0.f, res.fam_cells[f] = 0, res.fam_bases[f] = 0, res.fam_launches[f] = 0;
  if (n == 0) return;
  Impl &m = *impl_;
  timespec ts_entry;
  clock_gettime(CLOCK_MONOTONIC, &ts_entry);
  const double w_entry = ts_entry.tv_sec * 1e3 + ts_entry.tv_nsec * 1e-6;
  double w_prev_end = w_entry;

  // order by traceback size, largest first, so that waves are filled greedily and long problems start early
  std::vector<Geom> geo(n);
  std::vector<int> order;
  order.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    KswJob &j = jobs[i];
    if (j.qlen <= 0 || j.tlen <= 0) {  // ksw2_extd2_sse.c:71: reset only
      KswOut &o = res.out[i];
      o.max_q = o.max_t = o.mqe_t = o.mte_q = -1;
      o.score = o.mqe = o.mte = KSW_NEG_INF;
      continue;
    }
    geo[i] = make_geom(j.qlen, j.tlen, j.w, j.flag);
    if (is_small_fill(j) || wide_fill_cfg(j) >= 0) geo[i].p_bytes = (size_t)geo[i].n_row * (size_t)geo[i].T;  // rows indexed by target position
    order.push_back((int)i);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return geo[a].p_bytes > geo[b].p_bytes; });

  // problems whose state is larger than this keep it in the global slab (L2-resident) instead of shared memory
  static const size_t state_limit = getenv("PGMM_STATE_SMEM_KB") ? std::min(kSmemMax, (size_t)atoi(getenv("PGMM_STATE_SMEM_KB")) * 1024) : kSmemMax;
  LaneConsts lane_consts;
  lane_consts.init(sc);
  size_t pos = 0;
  while (pos < order.size()) {
    // ---- carve one wave ----
    size_t p_used = 0, cig_used = 0, scr_used = 0, end = pos;
    while (end < order.size()) {
      const int i = order[end];
      const size_t pb = (geo[i].p_bytes + 255) / 256 * 256;
      if (end > pos && p_used + pb > arena_budget_bytes) break;
      jobs[i].p_off = p_used, p_used += pb;
      jobs[i].cig_off = cig_used, cig_used += (size_t)jobs[i].qlen + jobs[i].tlen + 2;
      if (geo[i].state_bytes > state_limit) jobs[i].scr_off = scr_used, scr_used += (geo[i].state_bytes + 255) / 256 * 256;
      else jobs[i].scr_off = ~0ull;
      ++end;
    }
    const size_t nw = end - pos;
    // ---- launch classes: threads per CTA follow the width of the wavefront (one word of 4 cells per thread and
    // anti-diagonal for everything but the small fills), shared memory follows the length of the target ----
    constexpr int kClasses = Impl::kClasses;
    static const int max_nt_tier = getenv("PGMM_MAX_NT_TIER") ? atoi(getenv("PGMM_MAX_NT_TIER")) : 3;  // 256 threads: a 512-thread CTA (113 registers) owns its SM and leaves 70 % of its issue slots idle (profiles/r02_sweep_nt.txt)
    std::vector<int> cls[kClasses], generic;
    size_t cls_smem[kClasses] = {};
    KswJob *hj = m.h_jobs.ensure(nw);
    for (size_t k = 0; k < nw; ++k) {  // jobs of this wave are renumbered 0..nw-1 on the device
      const int i = order[pos + k];
      hj[k] = jobs[i];
      const size_t sb = geo[i].state_bytes;
      const int ww = jobs[i].w < 0 ? INT32_MAX : jobs[i].w;
      const int front = std::min(std::min(jobs[i].qlen, jobs[i].tlen), ww < INT32_MAX ? ww + 1 : INT32_MAX);
      const int words = (front + 32 + 3) / 4;
      if (is_small_fill(jobs[i])) {
        cls[20].push_back((int)k);
        cls_smem[20] = std::max<size_t>(cls_smem[20], (size_t)jobs[i].qlen);  // here: the longest query of the class
        res.cells += (uint64_t)jobs[i].qlen * jobs[i].tlen;
        res.fam_cells[1] += (uint64_t)jobs[i].qlen * jobs[i].tlen, res.fam_bases[1] += (uint64_t)jobs[i].qlen + jobs[i].tlen;
        continue;
      }
      if (const int wc = wide_fill_cfg(jobs[i]); wc >= 0) {
        const int c = 21 + 2 * wc + (jobs[i].flag == 0 ? 1 : 0);
        cls[c].push_back((int)k);
        cls_smem[c] = std::max(cls_smem[c], wide_fill_smem(jobs[i]));
        res.cells += (uint64_t)jobs[i].qlen * jobs[i].tlen;
        res.fam_cells[2] += (uint64_t)jobs[i].qlen * jobs[i].tlen, res.fam_bases[2] += (uint64_t)jobs[i].qlen + jobs[i].tlen;
        continue;
      }
      int nt_tier, sm_tier;
      if (sb <= 6 * 1024 && words <= 96) nt_tier = 0;  // small fills: one warp, a few words per lane
      else nt_tier = words <= 64 ? 1 : words <= 128 ? 2 : words <= 256 ? 3 : 4;
      if (nt_tier > max_nt_tier) nt_tier = max_nt_tier;
      sm_tier = sb > state_limit ? 3 : sb <= 12 * 1024 ? 0 : sb <= 48 * 1024 ? 1 : 2;
      const int c = nt_tier * 4 + sm_tier;
      cls[c].push_back((int)k);
      if (sm_tier < 3) cls_smem[c] = std::max(cls_smem[c], sb);
      const uint64_t band_cells = (uint64_t)std::min<int64_t>((int64_t)jobs[i].qlen * jobs[i].tlen,
                                                              (int64_t)geo[i].n_row * std::min(geo[i].n_col16, geo[i].T));
      res.cells += band_cells;  // asked for (upper bound: an extension that z-drops stops early)
      res.fam_bases[0] += (uint64_t)jobs[i].qlen + jobs[i].tlen;
      generic.push_back((int)k);
    }
    int *hid = m.h_ids.ensure(nw);
    size_t cls_off[kClasses], nid = 0;
    for (int c = 0; c < kClasses; ++c) {
      cls_off[c] = nid;
      for (int k : cls[c]) hid[nid++] = k;
    }
    m.d_jobs.ensure(nw), m.d_outs.ensure(nw), m.d_ids.ensure(nw), m.d_counter.ensure(8);
    // the traceback arena grows on demand up to its budget, in steps of at least half its size: a cudaMalloc stalls every
    // stream of the device, so growth has to stop after the first few waves
    if (p_used + 256 > m.p_arena.cap) {
      const size_t want = std::max<size_t>(p_used + 256, std::min<size_t>(arena_budget_bytes + 256, std::max<size_t>(m.p_arena.cap + m.p_arena.cap / 2, size_t(256) << 20)));
      if (getenv("PGMM_TRACE")) fprintf(stderr, "[pgmm trace] traceback arena grows from %zu to %zu MB (budget %zu MB)\n", m.p_arena.cap >> 20, want >> 20, arena_budget_bytes >> 20);
      m.p_arena.ensure_exact(want);
    }
    m.cig_arena.ensure(2 * cig_used + 4), m.cig_packed.ensure(2 * cig_used + 4), m.scratch.ensure(scr_used + 256);
    PGMM_CUDA(cudaMemcpyAsync(m.d_jobs.p, hj, nw * sizeof(KswJob), cudaMemcpyHostToDevice, stream));
    PGMM_CUDA(cudaMemcpyAsync(m.d_ids.p, hid, nw * sizeof(int), cudaMemcpyHostToDevice, stream));
    static const bool trace_spans = getenv("PGMM_TRACE") != nullptr;
    const auto wall = [] {
      timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    const double w_submit = wall();
    PGMM_CUDA(cudaMemsetAsync(m.d_counter.p, 0, 8 * sizeof(unsigned long long), stream));
    PGMM_CUDA(cudaMemsetAsync(m.d_counter.p + 1, 0xff, sizeof(unsigned long long), stream));
    if (trace_spans) stamp_kernel<<<1, 1, 0, stream>>>(m.d_counter.p + 3);
    PGMM_CUDA(cudaEventRecord(m.ev0, stream));
    // the size classes are independent launches: fork them onto their own streams so that the few long problems of the
    // large classes overlap with the many short ones instead of queueing behind each other
    PGMM_CUDA(cudaEventRecord(m.fork, stream));
    static const bool no_fork = getenv("PGMM_NO_CLASS_STREAMS") != nullptr;
    // PGMM_WIDE_SMEM_KB: shared memory a long fill asks for at least (a large value keeps other CTAs off its SM)
    static const size_t wide_min_smem = getenv("PGMM_WIDE_SMEM_KB") ? std::min(kSmemMax, (size_t)atoi(getenv("PGMM_WIDE_SMEM_KB")) * 1024) : 0;
    for (int c = kClasses - 1; c >= 0; --c) {  // widest / longest first
      if (cls[c].empty()) continue;
      cudaStream_t cs = no_fork ? stream : m.stream_of(c);
      if (!no_fork) PGMM_CUDA(cudaStreamWaitEvent(cs, m.fork, 0));
      PGMM_CUDA(cudaEventRecord(m.cls_t0[c], cs));
#define PGMM_LAUNCH(NT, SMEM)                                                                                                     \
  launch_class<NT>(cls[c], SMEM, m.d_ids.p, cls_off[c], m.d_jobs.p, d_q, d_t, sc, m.p_arena.p, m.cig_arena.p, m.scratch.p, m.d_outs.p, \
                   m.cig_packed.p, m.d_counter.p, cs)
      if (c == 20) {
        const int q_cap = ((int)cls_smem[20] + 15) / 16 * 16;
        const size_t per_warp = ((size_t)(kFillMaxT + 16) + (q_cap + 32) + 4 * (size_t)(q_cap + 4 * kFillMaxT) + 15) / 16 * 16;
        static bool attr_set = false;
        if (!attr_set) {
          PGMM_CUDA(cudaFuncSetAttribute(ksw_fill_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
          attr_set = true;
        }
        const int nj = (int)cls[20].size();
        ksw_fill_small_kernel<<<(nj + kFillWarps - 1) / kFillWarps, kFillWarps * 32, per_warp * kFillWarps, cs>>>(
            m.d_jobs.p, m.d_ids.p + cls_off[20], nj, d_q, d_t, sc, lane_consts, q_cap, m.p_arena.p, m.cig_arena.p, m.d_outs.p, m.cig_packed.p, m.d_counter.p);
        PGMM_CUDA(cudaGetLastError());
      } else if (c > 20) {
#define PGMM_WIDE(NW, KP, EX)                                                                                                   \
  do {                                                                                                                          \
    static bool attr_set = false;                                                                                               \
    if (!attr_set) {                                                                                                            \
      PGMM_CUDA(cudaFuncSetAttribute(ksw_fill_wide_kernel<NW, KP, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)); \
      attr_set = true;                                                                                                          \
    }                                                                                                                           \
    ksw_fill_wide_kernel<NW, KP, EX><<<(unsigned)cls[c].size(), NW * 32, std::max(cls_smem[c], wide_min_smem), cs>>>(                                    \
        m.d_jobs.p, m.d_ids.p + cls_off[c], d_q, d_t, sc, lane_consts, m.p_arena.p, m.cig_arena.p, m.d_outs.p, m.cig_packed.p, m.d_counter.p); \
    PGMM_CUDA(cudaGetLastError());                                                                                              \
  } while (0)
        switch (c - 21) {
          case 0: PGMM_WIDE(4, 1, false); break;
          case 1: PGMM_WIDE(4, 1, true); break;
          case 2: PGMM_WIDE(8, 2, false); break;
          case 3: PGMM_WIDE(8, 2, true); break;
          case 4: PGMM_WIDE(16, 4, false); break;
          case 5: PGMM_WIDE(16, 4, true); break;
          case 6: PGMM_WIDE(16, 8, false); break;
          default: PGMM_WIDE(16, 8, true); break;
        }
#undef PGMM_WIDE
      } else
      switch (c / 4) {
        case 0: PGMM_LAUNCH(32, cls_smem[c]); break;
        case 1: PGMM_LAUNCH(64, cls_smem[c]); break;
        case 2: PGMM_LAUNCH(128, cls_smem[c]); break;
        case 3: PGMM_LAUNCH(256, cls_smem[c]); break;
        default: PGMM_LAUNCH(512, cls_smem[c]); break;
      }
#undef PGMM_LAUNCH
      PGMM_CUDA(cudaEventRecord(m.cls_t1[c], cs));
      PGMM_CUDA(cudaEventRecord(m.cls_done[c], cs));
      ++res.launches;
    }
    // The classes are joined on the HOST.  A cudaStreamWaitEvent on the main stream would park that wait at the head of
    // one of the device's (at most 32) hardware queues for the whole wave, and the streams of other rounds that share the
    // queue would stall behind it: with two dozen rounds in flight that serialises rounds against each other.
    static const bool host_join = getenv("PGMM_DEVICE_JOIN") == nullptr;
    for (int c = 0; c < kClasses && !no_fork; ++c) {
      if (cls[c].empty()) continue;
      if (host_join) PGMM_CUDA(cudaEventSynchronize(m.cls_done[c]));
      else PGMM_CUDA(cudaStreamWaitEvent(stream, m.cls_done[c], 0));
    }
    PGMM_CUDA(cudaEventRecord(m.ev1, stream));
    const double w_joined = wall();
    if (trace_spans) stamp_kernel<<<1, 1, 0, stream>>>(m.d_counter.p + 4);

    // ---- results of this wave: ez + the number of CIGAR words, then exactly those words ----
    KswOut *ho = m.h_outs.ensure(nw);
    unsigned long long *hc = m.h_counter.ensure(8);
    PGMM_CUDA(cudaMemcpyAsync(ho, m.d_outs.p, nw * sizeof(KswOut), cudaMemcpyDeviceToHost, stream));
    PGMM_CUDA(cudaMemcpyAsync(hc, m.d_counter.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    PGMM_CUDA(cudaStreamSynchronize(stream));
    if (trace_spans) {
      const double w_done = wall();
      fprintf(stderr, "[pgmm trace] dp span: %zu jobs, %d launches; host: prep %.2f ms, enqueue+join %.2f ms, results %.2f ms; gpu: main stream reached the wave -> first CTA %.2f ms, "
              "first CTA -> last CTA end %.2f ms, last CTA end -> main stream after the join %.2f ms\n", nw, res.launches, w_submit - w_prev_end, w_joined - w_submit, w_done - w_joined,
              hc[1] == ~0ull ? 0.0 : ((double)hc[1] - (double)hc[3]) * 1e-6, hc[1] == ~0ull ? 0.0 : ((double)hc[2] - (double)hc[1]) * 1e-6,
              hc[1] == ~0ull ? 0.0 : ((double)hc[4] - (double)hc[2]) * 1e-6);
      w_prev_end = w_done;
    }
    float ms = 0.f;
    PGMM_CUDA(cudaEventElapsedTime(&ms, m.ev0, m.ev1));
    res.kernel_ms += ms;
    for (int c = 0; c < kClasses; ++c) {
      if (cls[c].empty()) continue;
      const int fam = c < 20 ? 0 : c == 20 ? 1 : 2;
      PGMM_CUDA(cudaEventElapsedTime(&ms, m.cls_t0[c], m.cls_t1[c]));
      res.fam_ms[fam] += ms, res.fam_launches[fam] += 1;
    }
    const size_t tot = (size_t)*hc, base = res.cigar.size();
    if (tot > 0) {
      uint32_t *hcig = m.h_cigar.ensure(tot);
      PGMM_CUDA(cudaMemcpyAsync(hcig, m.cig_packed.p, tot * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      PGMM_CUDA(cudaStreamSynchronize(stream));
      res.cigar.insert(res.cigar.end(), hcig, hcig + tot);
    }
    for (size_t k = 0; k < nw; ++k) {
      const int i = order[pos + k];
      res.out[i] = ho[k];
      res.cig_start[i] = base + ho[k].cig_pos;
    }
    // K5 generic: the in-band cells of the anti-diagonals that were actually evaluated (ksw2_extd2_sse.c:137-147)
    for (int k : generic) {
      const KswJob &j = jobs[order[pos + k]];
      const int w = j.w < 0 ? std::max(j.qlen, j.tlen) : j.w, nd = std::min(ho[k].n_diag, j.qlen + j.tlen - 1);
      uint64_t c = 0;
      for (int r = 0; r < nd; ++r) {
        const int st = std::max(std::max(0, r - j.qlen + 1), (r - w + 1) >> 1), en = std::min(std::min(j.tlen - 1, r), (r + w) >> 1);
        if (en >= st) c += (uint64_t)(en - st + 1);
      }
      res.fam_cells[0] += c;
    }
    pos = end;
  }
  if (getenv("PGMM_DUMP_JOBS")) {
    FILE *fp = fopen(getenv("PGMM_DUMP_JOBS"), "a");
    if (fp) {
      for (size_t i = 0; i < n; ++i)
        fprintf(fp, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", jobs[i].qlen, jobs[i].tlen, jobs[i].w, jobs[i].flag, res.out[i].n_diag, res.out[i].zdropped, res.out[i].reach_end, res.out[i].n_cigar);
      fprintf(fp, "#wave\t%f\n", res.kernel_ms);
      fclose(fp);
    }
  }
  res.cig_start[n] = res.cigar.size();
}

void trace_attach_ksw(PgmmCtaTraceRec *buf, unsigned long long *cnt, unsigned long long cap) { trace::attach(buf, cnt, cap); }

}  // namespace pgmm
