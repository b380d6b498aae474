// Option structs of the minimap2-sys boundary: defaults, presets, validation.
// Behaviour follows the reference's packages/minimap2-sys/minimap2/options.c (init :5-64, presets :88-162,
// checks :164-234); pangraph only ever selects asm5/asm10/asm20 (align_with_minimap2_lib.rs:42-47) but every preset
// name the reference accepts is accepted here so that mm_set_opt keeps its error behaviour.
#include "../../include/pgmm_b200.h"

#include <climits>
#include <cstdio>
#include <cstring>

extern "C" {

void mm_idxopt_init(mm_idxopt_t *o) {
  memset(o, 0, sizeof(*o));
  o->k = 15;
  o->w = 10;
  o->bucket_bits = 14;
  o->mini_batch_size = 50000000;
  o->batch_size = 8000000000ULL;
}

void mm_mapopt_init(mm_mapopt_t *o) {
  memset(o, 0, sizeof(*o));
  o->seed = 11;
  // seeding
  o->mid_occ_frac = 2e-4f, o->min_mid_occ = 10, o->max_mid_occ = 1000000, o->q_occ_frac = 0.01f;
  o->max_max_occ = 4095, o->occ_dist = 500;
  // chaining
  o->min_cnt = 3, o->min_chain_score = 40, o->bw = 500, o->bw_long = 20000, o->max_gap = 5000, o->max_gap_ref = -1;
  o->max_chain_skip = 25, o->max_chain_iter = 5000, o->chain_gap_scale = 0.8f, o->chain_skip_scale = 0.0f;
  o->rmq_inner_dist = 1000, o->rmq_size_cap = 100000, o->rmq_rescue_size = 1000, o->rmq_rescue_ratio = 0.1f;
  // hit selection
  o->mask_level = 0.5f, o->mask_len = INT_MAX, o->pri_ratio = 0.8f, o->best_n = 5, o->alt_drop = 0.15f;
  // base-level alignment
  o->a = 2, o->b = 4, o->q = 4, o->e = 2, o->q2 = 24, o->e2 = 1, o->sc_ambi = 1;
  o->zdrop = 400, o->zdrop_inv = 200, o->end_bonus = -1;
  o->min_dp_max = o->min_chain_score * o->a;
  o->min_ksw_len = 200, o->anchor_ext_len = 20, o->anchor_ext_shift = 6, o->max_clip_ratio = 1.0f;
  o->mini_batch_size = 500000000, o->max_sw_mat = 100000000, o->cap_kalloc = 1000000000;
  o->rank_min_len = 500, o->rank_frac = 0.9f;
  o->pe_ori = 0, o->pe_bonus = 33;
}

static void all_vs_all(mm_mapopt_t *mo) {
  mo->flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN;
  mo->min_chain_score = 100, mo->pri_ratio = 0.0f, mo->max_chain_skip = 25, mo->occ_dist = 0;
}

static void gaps(mm_mapopt_t *mo, int a, int b, int q, int e, int q2, int e2) {
  mo->a = a, mo->b = b, mo->q = q, mo->e = e, mo->q2 = q2, mo->e2 = e2;
}

int mm_set_opt(const char *preset, mm_idxopt_t *io, mm_mapopt_t *mo) {
  if (preset == nullptr) {
    mm_idxopt_init(io);
    mm_mapopt_init(mo);
    return 0;
  }
  const auto is = [&](const char *s) { return strcmp(preset, s) == 0; };
  if (is("map-ont")) return 0;
  if (is("ava-ont")) {
    io->flag = 0, io->k = 15, io->w = 5;
    all_vs_all(mo);
    mo->bw = mo->bw_long = 2000;
  } else if (is("map10k") || is("map-pb")) {
    io->flag |= MM_I_HPC, io->k = 19;
  } else if (is("ava-pb")) {
    io->flag |= MM_I_HPC, io->k = 19, io->w = 5;
    all_vs_all(mo);
    mo->bw_long = mo->bw;
  } else if (is("map-hifi") || is("map-ccs")) {
    io->flag = 0, io->k = 19, io->w = 19;
    mo->max_gap = 10000;
    gaps(mo, 1, 4, 6, 2, 26, 1);
    mo->occ_dist = 500, mo->min_mid_occ = 50, mo->max_mid_occ = 500, mo->min_dp_max = 200;
  } else if (strncmp(preset, "asm", 3) == 0) {
    // the three presets pangraph can reach; all of them turn on the RMQ chainer
    io->flag = 0, io->k = 19, io->w = 19;
    mo->bw = 1000, mo->bw_long = 100000, mo->max_gap = 10000;
    mo->flag |= MM_F_RMQ;
    mo->min_mid_occ = 50, mo->max_mid_occ = 500, mo->min_dp_max = 200, mo->best_n = 50;
    if (is("asm5")) gaps(mo, 1, 19, 39, 3, 81, 1);
    else if (is("asm10")) gaps(mo, 1, 9, 16, 2, 41, 1);
    else if (is("asm20")) gaps(mo, 1, 4, 6, 2, 26, 1), io->w = 10;
    else return -1;
    mo->zdrop = mo->zdrop_inv = 200;
  } else if (is("short") || is("sr")) {
    io->flag = 0, io->k = 21, io->w = 11;
    mo->flag |= MM_F_SR | MM_F_FRAG_MODE | MM_F_NO_PRINT_2ND | MM_F_2_IO_THREADS | MM_F_HEAP_SORT;
    mo->pe_ori = 0 << 1 | 1;
    gaps(mo, 2, 8, 12, 2, 24, 1);
    mo->zdrop = mo->zdrop_inv = 100, mo->end_bonus = 10, mo->max_frag_len = 800, mo->max_gap = 100;
    mo->bw = mo->bw_long = 100, mo->pri_ratio = 0.5f, mo->min_cnt = 2, mo->min_chain_score = 25, mo->min_dp_max = 40;
    mo->best_n = 20, mo->mid_occ = 1000, mo->max_occ = 5000, mo->mini_batch_size = 50000000;
  } else if (strncmp(preset, "splice", 6) == 0 || is("cdna")) {
    io->flag = 0, io->k = 15, io->w = 5;
    mo->flag |= MM_F_SPLICE | MM_F_SPLICE_FOR | MM_F_SPLICE_REV | MM_F_SPLICE_FLANK;
    mo->max_sw_mat = 0, mo->max_gap = 2000, mo->max_gap_ref = mo->bw = mo->bw_long = 200000;
    gaps(mo, 1, 2, 2, 1, 32, 0);
    mo->noncan = 9, mo->junc_bonus = 9, mo->zdrop = 200, mo->zdrop_inv = 100;
    if (is("splice:hq")) mo->junc_bonus = 5, mo->b = 4, mo->q = 6, mo->q2 = 24;
  } else
    return -1;
  return 0;
}

int mm_check_opt(const mm_idxopt_t *io, const mm_mapopt_t *mo) {
  // same order and codes as the reference, so the first failing rule decides the return value
  const auto fail = [](int code, const char *msg) {
    fprintf(stderr, "[ERROR] %s\n", msg);
    return code;
  };
  if (mo->bw > mo->bw_long) return fail(-8, "with '-rNUM1,NUM2', NUM1 can't be larger than NUM2");
  if ((mo->flag & MM_F_RMQ) && (mo->flag & (MM_F_SR | MM_F_SPLICE))) return fail(-7, "--rmq doesn't work with --sr or --splice");
  if (mo->split_prefix && (mo->flag & (MM_F_OUT_CS | MM_F_OUT_MD))) return fail(-6, "--cs or --MD doesn't work with --split-prefix");
  if (io->k <= 0 || io->w <= 0) return fail(-5, "-k and -w must be positive");
  if (mo->best_n < 0) return fail(-4, "-N must be no less than 0");
  if (mo->pri_ratio < 0.0f || mo->pri_ratio > 1.0f) return fail(-4, "-p must be within 0 and 1 (including 0 and 1)");
  if ((mo->flag & MM_F_FOR_ONLY) && (mo->flag & MM_F_REV_ONLY)) return fail(-3, "--for-only and --rev-only can't be applied at the same time");
  if (mo->e <= 0 || mo->q <= 0) return fail(-1, "-O and -E must be positive");
  if ((mo->q != mo->q2 || mo->e != mo->e2) && !(mo->e > mo->e2 && mo->q + mo->e < mo->q2 + mo->e2))
    return fail(-2, "dual gap penalties violating E1>E2 and O1+E1<O2+E2");
  if ((mo->q + mo->e) + (mo->q2 + mo->e2) > 127) return fail(-1, "scoring system violating ({-O}+{-E})+({-O2}+{-E2}) <= 127");
  if (mo->zdrop < mo->zdrop_inv) return fail(-5, "Z-drop should not be less than inversion-Z-drop");
  if ((mo->flag & MM_F_NO_PRINT_2ND) && (mo->flag & MM_F_ALL_CHAINS)) return fail(-5, "-X/-P and --secondary=no can't be applied at the same time");
  if ((mo->flag & MM_F_QSTRAND) && ((mo->flag & (MM_F_OUT_SAM | MM_F_SPLICE | MM_F_FRAG_MODE)) || (io->flag & MM_I_HPC)))
    return fail(-5, "--qstrand doesn't work with -a, -H, --frag or --splice");
  return 0;
}

}  // extern "C"
