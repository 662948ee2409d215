/*
 * bg_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See bg_oracle.h.
 *
 * Plain-C restatement of the reference hot path.  Every function cites the reference
 * file:line it follows.  The control flow is kept literal (including the reference's
 * quirks Q1/Q2 of SURVEY.md appendix A.3 and the doubles "only move at its level" flags),
 * so that bit-exactness against the Python reference is a property of this file and the
 * CUDA kernels are then checked against it.
 */
#include "bg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* board basics                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* src/backgammon/board/immutable_board.py:26-70 */
void bgo_initial_board(bgo_board* b) {
  memset(b, 0, sizeof(*b));
  b->p[0][0] = 2;
  b->p[0][11] = 5;
  b->p[0][16] = 3;
  b->p[0][18] = 5;
  b->p[1][23] = 2;
  b->p[1][12] = 5;
  b->p[1][7] = 3;
  b->p[1][5] = 5;
}

/* src/backgammon/moves/conditions.py:152-194 */
static int all_checkers_home(const bgo_board* b, int player) {
  if (b->bar[player] > 0) return 0;
  int lo = player == 1 ? 0 : 18, hi = player == 1 ? 6 : 24;
  int total = 0;
  for (int idx = 0; idx < BGO_NPOINTS; ++idx) {
    int n = b->p[player][idx];
    if (n > 0) {
      if (idx >= lo && idx < hi)
        total += n;
      else
        return 0;
    }
  }
  return total + b->off[player] == 15;
}

enum { ST_NORMAL, ST_ON_BAR, ST_BEAR_OFF, ST_GAME_OVER };

/* src/backgammon/moves/conditions.py:5-22 */
static int compute_board_state(const bgo_board* b, int player) {
  if (b->off[player] == 15) return ST_GAME_OVER;
  if (b->bar[player] > 0) return ST_ON_BAR;
  if (all_checkers_home(b, player)) return ST_BEAR_OFF;
  return ST_NORMAL;
}

/* src/backgammon/moves/get_moves_one_die.py:13-251 */
int bgo_moves_one_die(const bgo_board* b, int die, int player, bgo_submove* out) {
  int n = 0;
  int opp = 1 - player;
  int dir = player == 0 ? 1 : -1;
  int st = compute_board_state(b, player);
  if (st == ST_NORMAL) { /* :40-83 */
    for (int idx = 0; idx < BGO_NPOINTS; ++idx) {
      if (b->p[player][idx] > 0) {
        int dst = idx + die * dir;
        if (dst >= 0 && dst < BGO_NPOINTS) {
          if (b->p[opp][dst] < 2) { /* conditions.py:25-61 valid_move */
            out[n].start = (uint8_t)idx;
            out[n].end = (uint8_t)dst;
            out[n].hit = b->p[opp][dst] == 1; /* conditions.py:64-89 */
            ++n;
          }
        }
      }
    }
  } else if (st == ST_ON_BAR) { /* :86-130 */
    int dst = player == 0 ? die - 1 : 24 - die;
    int ok_range = player == 0 ? (dst >= 0 && dst < 6) : (dst >= 18 && dst < 24);
    if (ok_range && b->p[opp][dst] < 2) { /* conditions.py:92-119 */
      out[n].start = BGO_BAR;
      out[n].end = (uint8_t)dst;
      out[n].hit = b->p[opp][dst] == 1;
      ++n;
    }
  } else if (st == ST_BEAR_OFF) { /* :133-251 */
    int lo = player == 0 ? 18 : 0, hi = player == 0 ? 24 : 6;
    int last = player == 0 ? 18 : 5;
    /* 1. in-board moves inside home, ascending index for both players (:164-189) */
    for (int idx = lo; idx < hi; ++idx) {
      if (b->p[player][idx] > 0) {
        int dst = idx + die * dir;
        if (dst >= 0 && dst < BGO_NPOINTS) {
          if (b->p[opp][dst] < 2) {
            out[n].start = (uint8_t)idx;
            out[n].end = (uint8_t)dst;
            out[n].hit = b->p[opp][dst] == 1;
            ++n;
          }
        }
      }
    }
    /* 2. farthest checker from the exit (:192-203) */
    if (player == 0) {
      for (int idx = lo; idx < hi; ++idx)
        if (b->p[0][idx] > 0) {
          last = idx;
          break;
        }
    } else {
      for (int idx = hi - 1; idx >= lo; --idx)
        if (b->p[1][idx] > 0) {
          last = idx;
          break;
        }
    }
    /* 3. bear-off moves (:206-249) */
    if (player == 0) {
      if (last + die * dir >= BGO_NPOINTS) {
        out[n].start = (uint8_t)last;
        out[n].end = BGO_BEAR_OFF;
        out[n].hit = 0;
        ++n;
      }
      int ps = BGO_NPOINTS - die;
      if (ps != last && ps >= lo && ps < hi && b->p[0][ps] > 0) {
        out[n].start = (uint8_t)ps;
        out[n].end = BGO_BEAR_OFF;
        out[n].hit = 0;
        ++n;
      }
    } else {
      if (last + die * dir < 0) {
        out[n].start = (uint8_t)last;
        out[n].end = BGO_BEAR_OFF;
        out[n].hit = 0;
        ++n;
      }
      int ps = die - 1;
      if (ps != last && ps >= lo && ps < hi && b->p[1][ps] > 0) {
        out[n].start = (uint8_t)ps;
        out[n].end = BGO_BEAR_OFF;
        out[n].hit = 0;
        ++n;
      }
    }
  }
  return n;
}

/* src/backgammon/board/immutable_board.py:183-258 */
void bgo_move_checker(const bgo_board* b, int player, bgo_submove sm, bgo_board* out) {
  bgo_board t = *b;
  int opp = 1 - player;
  if (sm.start == BGO_BAR) {
    if (t.bar[player] <= 0) {
      *out = *b;
      return;
    }
    t.bar[player] -= 1;
  } else {
    if (t.p[player][sm.start] <= 0) {
      *out = *b;
      return;
    }
    t.p[player][sm.start] -= 1;
  }
  if (sm.hit) {
    if (t.p[opp][sm.end] == 1) {
      t.p[opp][sm.end] -= 1;
      t.bar[opp] += 1;
    } else {
      *out = *b;
      return;
    }
  }
  if (sm.end == BGO_BEAR_OFF)
    t.off[player] += 1;
  else
    t.p[player][sm.end] += 1;
  *out = t;
}

/* src/environments/env_helper.py:27-91 */
void bgo_execute_full_move(const bgo_board* b, int player, const bgo_fullmove* fm, bgo_board* out) {
  bgo_board t = *b;
  int opp = 1 - player;
  for (int i = 0; i < fm->n; ++i) {
    bgo_submove s = fm->sm[i];
    if (s.start == BGO_BAR)
      t.bar[player] -= 1;
    else
      t.p[player][s.start] -= 1;
    if (s.hit) {
      t.p[opp][s.end] -= 1;
      t.bar[opp] += 1;
    }
    if (s.end == BGO_BEAR_OFF)
      t.off[player] += 1;
    else
      t.p[player][s.end] += 1;
  }
  *out = t;
}

/* ------------------------------------------------------------------------------------------ */
/* full-move generation                                                                       */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  bgo_fullmove* moves;
  bgo_board* boards;
  int n, cap;
  int32_t* table; /* open addressing over `boards`, -1 empty */
  int32_t* slot_of;
  int tcap;
} gen_ctx;

static __thread gen_ctx g_ctx;

static uint64_t board_hash(const bgo_board* b) {
  const uint8_t* p = (const uint8_t*)b;
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < BGO_BOARD_BYTES; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h ^ (h >> 29);
}

static void ctx_reset(gen_ctx* c) {
  if (!c->moves) {
    c->cap = 1024;
    c->moves = (bgo_fullmove*)malloc(sizeof(bgo_fullmove) * c->cap);
    c->boards = (bgo_board*)malloc(sizeof(bgo_board) * c->cap);
    c->slot_of = (int32_t*)malloc(sizeof(int32_t) * c->cap);
    c->tcap = 4096;
    c->table = (int32_t*)malloc(sizeof(int32_t) * c->tcap);
    memset(c->table, 0xff, sizeof(int32_t) * c->tcap);
  } else {
    for (int i = 0; i < c->n; ++i) c->table[c->slot_of[i]] = -1; /* clear only the used slots */
  }
  c->n = 0;
}

static void ctx_grow(gen_ctx* c) {
  c->cap *= 2;
  c->moves = (bgo_fullmove*)realloc(c->moves, sizeof(bgo_fullmove) * c->cap);
  c->boards = (bgo_board*)realloc(c->boards, sizeof(bgo_board) * c->cap);
  c->slot_of = (int32_t*)realloc(c->slot_of, sizeof(int32_t) * c->cap);
  free(c->table);
  c->tcap *= 2;
  c->table = (int32_t*)malloc(sizeof(int32_t) * c->tcap);
  memset(c->table, 0xff, sizeof(int32_t) * c->tcap);
  for (int i = 0; i < c->n; ++i) {
    uint64_t h = board_hash(&c->boards[i]) & (uint64_t)(c->tcap - 1);
    while (c->table[h] >= 0) h = (h + 1) & (uint64_t)(c->tcap - 1);
    c->table[h] = i;
    c->slot_of[i] = (int32_t)h;
  }
}

/* src/backgammon/moves/handle_move_types.py:196-221 */
static void add_unique_board(gen_ctx* c, const bgo_board* board, const bgo_submove* sm, int nsm) {
  uint64_t h = board_hash(board) & (uint64_t)(c->tcap - 1);
  while (c->table[h] >= 0) {
    if (memcmp(&c->boards[c->table[h]], board, BGO_BOARD_BYTES) == 0) return; /* board in unique_boards */
    h = (h + 1) & (uint64_t)(c->tcap - 1);
  }
  if (c->n == c->cap) {
    ctx_grow(c);
    h = board_hash(board) & (uint64_t)(c->tcap - 1);
    while (c->table[h] >= 0) h = (h + 1) & (uint64_t)(c->tcap - 1);
  }
  c->table[h] = c->n;
  c->slot_of[c->n] = (int32_t)h;
  c->boards[c->n] = *board;
  bgo_fullmove* fm = &c->moves[c->n];
  memset(fm, 0, sizeof(*fm));
  for (int i = 0; i < nsm; ++i) fm->sm[i] = sm[i];
  fm->n = (uint8_t)nsm;
  c->n += 1;
}

/* src/backgammon/moves/handle_move_types.py:7-81 */
static void handle_non_doubles(gen_ctx* c, const bgo_board* board, int die_first, int die_second,
                               int player) {
  bgo_submove first[32], second[32], seq[2];
  int n1 = bgo_moves_one_die(board, die_first, player, first);
  int two_generated = 0;
  for (int i = 0; i < n1; ++i) {
    bgo_board b1, b2;
    bgo_move_checker(board, player, first[i], &b1);
    int n2 = bgo_moves_one_die(&b1, die_second, player, second);
    if (n2) {
      two_generated = 1;
      for (int j = 0; j < n2; ++j) {
        bgo_move_checker(&b1, player, second[j], &b2);
        seq[0] = first[i];
        seq[1] = second[j];
        add_unique_board(c, &b2, seq, 2);
      }
    }
  }
  if (!two_generated) {
    for (int i = 0; i < n1; ++i) {
      bgo_board b1;
      bgo_move_checker(board, player, first[i], &b1);
      add_unique_board(c, &b1, &first[i], 1);
    }
  }
}

/* src/backgammon/moves/handle_move_types.py:84-193 */
static void handle_doubles(gen_ctx* c, const bgo_board* board, int die, int player) {
  bgo_submove m1[32], m2[32], m3[32], m4[32], seq[4];
  int n1 = bgo_moves_one_die(board, die, player, m1);
  int len4_possible = 0;
  for (int a = 0; a < n1; ++a) {
    bgo_board b1;
    bgo_move_checker(board, player, m1[a], &b1);
    seq[0] = m1[a];
    int n2 = bgo_moves_one_die(&b1, die, player, m2);
    if (!n2 && n1 == 1 && !len4_possible) add_unique_board(c, &b1, seq, 1);
    for (int bq = 0; bq < n2; ++bq) {
      bgo_board b2;
      bgo_move_checker(&b1, player, m2[bq], &b2);
      seq[1] = m2[bq];
      int n3 = bgo_moves_one_die(&b2, die, player, m3);
      if (!n3 && n2 == 1 && !len4_possible) add_unique_board(c, &b2, seq, 2);
      for (int cq = 0; cq < n3; ++cq) {
        bgo_board b3;
        bgo_move_checker(&b2, player, m3[cq], &b3);
        seq[2] = m3[cq];
        int n4 = bgo_moves_one_die(&b3, die, player, m4);
        if (!n4 && n3 == 1 && !len4_possible) add_unique_board(c, &b3, seq, 3);
        for (int dq = 0; dq < n4; ++dq) {
          bgo_board b4;
          bgo_move_checker(&b3, player, m4[dq], &b4);
          seq[3] = m4[dq];
          add_unique_board(c, &b4, seq, 4);
          len4_possible = 1;
        }
      }
    }
  }
}

/* src/backgammon/moves/generate_all_moves.py:7-90 */
int bgo_get_all_possible_moves(int player, const bgo_board* b, int d0, int d1, bgo_fullmove* out_moves,
                               bgo_board* out_boards, int cap) {
  gen_ctx* c = &g_ctx;
  ctx_reset(c);
  if (d0 != d1) {
    int hi = d0 > d1 ? d0 : d1, lo = d0 > d1 ? d1 : d0;
    handle_non_doubles(c, b, hi, lo, player);
    /* :40-50  "if not full_moves or not (len == 1 and len(sub_moves) == 1)" */
    if (c->n == 0 || !(c->n == 1 && c->moves[0].n == 1)) handle_non_doubles(c, b, lo, hi, player);
  } else {
    handle_doubles(c, b, d0, player);
  }
  /* filter_full_moves_by_max_submoves :69-90 */
  int maxn = 0;
  for (int i = 0; i < c->n; ++i)
    if (c->moves[i].n > maxn) maxn = c->moves[i].n;
  int k = 0;
  for (int i = 0; i < c->n; ++i) {
    if (c->moves[i].n == maxn) {
      if (k < cap) {
        if (out_moves) out_moves[k] = c->moves[i];
        if (out_boards) out_boards[k] = c->boards[i];
      }
      ++k;
    }
  }
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* features + value net                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* src/backgammon/board/immutable_board.py:86-128 */
void bgo_features(const bgo_board* b, int flag_player, float* out) {
  for (int pl = 0; pl < 2; ++pl)
    for (int pt = 0; pt < 24; ++pt) {
      int cnt = b->p[pl][pt];
      float* f = out + pl * 96 + pt * 4;
      f[0] = cnt >= 1 ? 1.0f : 0.0f;
      f[1] = cnt >= 2 ? 1.0f : 0.0f;
      f[2] = cnt >= 3 ? 1.0f : 0.0f;
      int ex = cnt - 3;
      if (ex < 0) ex = 0;
      f[3] = (float)ex / 2.0f;
    }
  out[192] = (float)((double)b->bar[0] / 2.0);
  out[193] = (float)((double)b->off[0] / 15.0);
  out[194] = (float)((double)b->bar[1] / 2.0);
  out[195] = (float)((double)b->off[1] / 15.0);
  out[196] = flag_player == 0 ? 1.0f : 0.0f;
  out[197] = flag_player == 1 ? 1.0f : 0.0f;
}

/* src/agents/policy_network.py:53-70; accumulation in double (the reference's fp32 summation order
 * is torch-internal and unspecified; double is the neutral anchor for the 1e-5 tolerance). */
void bgo_forward(const float* packed, int H, const float* feats, int64_t n, float* out_v) {
  const float* W1t = packed;
  const float* b1 = packed + (int64_t)BGO_NFEAT * H;
  const float* w2 = b1 + H;
  const float b2 = w2[H];
  double* z = (double*)malloc(sizeof(double) * (size_t)H);
  for (int64_t r = 0; r < n; ++r) {
    const float* x = feats + r * BGO_NFEAT;
    for (int h = 0; h < H; ++h) z[h] = (double)b1[h];
    for (int f = 0; f < BGO_NFEAT; ++f) {
      if (x[f] != 0.0f) {
        double xf = (double)x[f];
        const float* w = W1t + (int64_t)f * H;
        for (int h = 0; h < H; ++h) z[h] += xf * (double)w[h];
      }
    }
    double v = (double)b2;
    for (int h = 0; h < H; ++h) v += (double)w2[h] / (1.0 + exp(-z[h]));
    out_v[r] = (float)v;
  }
  free(z);
}

/* fast fp32 path used for CPU-baseline timing (same maths, float accumulation over non-zero features) */
static void eval_board_f32(const float* packed, int H, const bgo_board* b, int flag, float* zbuf, float* out) {
  float x[BGO_NFEAT];
  bgo_features(b, flag, x);
  const float* W1t = packed;
  const float* b1 = packed + (int64_t)BGO_NFEAT * H;
  const float* w2 = b1 + H;
  for (int h = 0; h < H; ++h) zbuf[h] = b1[h];
  for (int f = 0; f < BGO_NFEAT; ++f) {
    float xf = x[f];
    if (xf != 0.0f) {
      const float* w = W1t + (int64_t)f * H;
      for (int h = 0; h < H; ++h) zbuf[h] += xf * w[h];
    }
  }
  float v = w2[H];
  for (int h = 0; h < H; ++h) v += w2[h] / (1.0f + expf(-zbuf[h]));
  *out = v;
}

void bgo_eval_boards(const float* packed, int H, const int8_t* boards, const uint8_t* flags, int64_t n,
                     float* out_v) {
  float* z = (float*)malloc(sizeof(float) * (size_t)H);
  for (int64_t i = 0; i < n; ++i)
    eval_board_f32(packed, H, (const bgo_board*)(boards + i * BGO_BOARD_BYTES), flags[i], z, &out_v[i]);
  free(z);
}

/* ------------------------------------------------------------------------------------------ */
/* terminal / shaping predicates (src/environments/env_helper.py:113-242)                     */
/* ------------------------------------------------------------------------------------------ */

int bgo_check_game_over(const bgo_board* b, int player) { return b->off[player] >= 15; }

int bgo_check_gammon(const bgo_board* b, int player) { return b->off[1 - player] == 0; }

int bgo_check_backgammon(const bgo_board* b, int player) {
  int opp = 1 - player;
  if (b->off[opp] > 0) return 0;
  int lo = player == 0 ? 18 : 0, hi = player == 0 ? 24 : 6;
  for (int idx = lo; idx < hi; ++idx)
    if (b->p[opp][idx] > 0) return 1;
  if (b->bar[opp] > 0) return 1;
  return 0;
}

int bgo_made_five_prime(const bgo_board* b, int player) {
  int opp = 1 - player;
  int run = 0;
  if (player == 0) {
    for (int idx = 0; idx < 24; ++idx) {
      if (b->p[0][idx] >= 2)
        ++run;
      else
        run = 0;
      if (run >= 5) {
        for (int i = idx + 1; i < 24; ++i)
          if (b->p[opp][i] > 0) return 1;
      }
    }
  } else {
    for (int idx = 23; idx >= 0; --idx) {
      if (b->p[1][idx] >= 2)
        ++run;
      else
        run = 0;
      if (run >= 5) {
        for (int i = 0; i < idx; ++i)
          if (b->p[opp][i] > 0) return 1;
      }
    }
  }
  return 0;
}

int bgo_is_closed_out(const bgo_board* b, int player) {
  int opp = 1 - player;
  if (b->bar[opp] == 0) return 0;
  int lo = player == 0 ? 18 : 0, hi = player == 0 ? 24 : 6;
  for (int idx = lo; idx < hi; ++idx)
    if (b->p[player][idx] < 2) return 0;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* batch helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */

static int resolve_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  return nthreads;
#else
  (void)nthreads;
  return 1;
#endif
}

int64_t bgo_movegen_batch(const int8_t* boards, const uint8_t* players, const uint8_t* rolls, int64_t B,
                          int64_t pool_cap, int8_t* out_boards, uint8_t* out_submoves, int64_t* out_offsets,
                          int nthreads) {
  nthreads = resolve_threads(nthreads);
  /* pass 1: counts */
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
  for (int64_t i = 0; i < B; ++i) {
    out_offsets[i + 1] = bgo_get_all_possible_moves(players[i], (const bgo_board*)(boards + i * BGO_BOARD_BYTES),
                                                    rolls[2 * i], rolls[2 * i + 1], NULL, NULL, 0);
  }
  out_offsets[0] = 0;
  for (int64_t i = 0; i < B; ++i) out_offsets[i + 1] += out_offsets[i];
  int64_t total = out_offsets[B];
  if (total > pool_cap) return -total;
  if (!out_boards && !out_submoves) return total;
    /* pass 2: emit */
#pragma omp parallel num_threads(nthreads)
  {
    bgo_fullmove* mv = (bgo_fullmove*)malloc(sizeof(bgo_fullmove) * BGO_MAX_MOVES);
    bgo_board* bd = (bgo_board*)malloc(sizeof(bgo_board) * BGO_MAX_MOVES);
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < B; ++i) {
      int n = bgo_get_all_possible_moves(players[i], (const bgo_board*)(boards + i * BGO_BOARD_BYTES), rolls[2 * i],
                                         rolls[2 * i + 1], mv, bd, BGO_MAX_MOVES);
      int64_t o = out_offsets[i];
      for (int k = 0; k < n; ++k) {
        if (out_boards) memcpy(out_boards + (o + k) * BGO_BOARD_BYTES, &bd[k], BGO_BOARD_BYTES);
        if (out_submoves) {
          uint8_t* s = out_submoves + (o + k) * 12;
          for (int q = 0; q < 4; ++q) {
            if (q < mv[k].n) {
              s[3 * q] = mv[k].sm[q].start;
              s[3 * q + 1] = mv[k].sm[q].end;
              s[3 * q + 2] = mv[k].sm[q].hit;
            } else {
              s[3 * q] = 255;
              s[3 * q + 1] = 255;
              s[3 * q + 2] = 0;
            }
          }
        }
      }
    }
    free(mv);
    free(bd);
  }
  return total;
}

void bgo_encode_batch(const int8_t* boards, const uint8_t* flags, int64_t n, float* out, int nthreads) {
  nthreads = resolve_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < n; ++i)
    bgo_features((const bgo_board*)(boards + i * BGO_BOARD_BYTES), flags[i], out + i * BGO_NFEAT);
}

/* ------------------------------------------------------------------------------------------ */
/* 2-ply (src/multi/two_ply.py:10-35, 93-150)                                                 */
/* ------------------------------------------------------------------------------------------ */

static const uint8_t DICE_ROLLS[21][2] = {{1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 2},
                                          {2, 3}, {2, 4}, {2, 5}, {2, 6}, {3, 3}, {3, 4}, {3, 5},
                                          {3, 6}, {4, 4}, {4, 5}, {4, 6}, {5, 5}, {5, 6}, {6, 6}};
static const int ROLL_COUNTS[21] = {1, 2, 2, 2, 2, 2, 1, 2, 2, 2, 2, 1, 2, 2, 2, 1, 2, 2, 1, 2, 1};

static int cmp_desc(const void* a, const void* b) {
  float x = *(const float*)a, y = *(const float*)b;
  return (x < y) - (x > y);
}

float bgo_weighted_opponent_response(const bgo_board* cand, int opponent, const float* packed, int H,
                                     int top_k, int64_t* n_replies) {
  bgo_board* bd = (bgo_board*)malloc(sizeof(bgo_board) * BGO_MAX_MOVES);
  float* feats = (float*)malloc(sizeof(float) * BGO_NFEAT * BGO_MAX_MOVES);
  float* vals = (float*)malloc(sizeof(float) * BGO_MAX_MOVES);
  double total = 0.0;
  int64_t nrep = 0;
  for (int r = 0; r < 21; ++r) {
    int n = bgo_get_all_possible_moves(opponent, cand, DICE_ROLLS[r][0], DICE_ROLLS[r][1], NULL, bd, BGO_MAX_MOVES);
    if (n > BGO_MAX_MOVES) n = BGO_MAX_MOVES;
    if (n) {
      nrep += n;
      for (int i = 0; i < n; ++i) bgo_features(&bd[i], opponent, feats + (size_t)i * BGO_NFEAT);
      bgo_forward(packed, H, feats, n, vals);
      qsort(vals, (size_t)n, sizeof(float), cmp_desc);
      int k = n < top_k ? n : top_k;
      float s = 0.0f;
      for (int i = 0; i < k; ++i) s += vals[i];
      float mean = s / (float)k; /* torch fp32 mean */
      total += (double)mean * ((double)ROLL_COUNTS[r] / 36.0);
    }
  }
  free(bd);
  free(feats);
  free(vals);
  if (n_replies) *n_replies = nrep;
  return (float)total;
}

/* src/multi/two_ply.py:44-90: score = alpha * S - beta * W (opponent = 1 - mover) */
void bgo_two_ply_batch(const int8_t* cand_boards, const uint8_t* mover, const float* S, int64_t n,
                       const float* packed, int H, int top_k, float alpha, float beta, float* out_score,
                       int64_t* out_replies, int nthreads) {
  nthreads = resolve_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int64_t i = 0; i < n; ++i) {
    int64_t nrep = 0;
    float W = bgo_weighted_opponent_response((const bgo_board*)(cand_boards + i * BGO_BOARD_BYTES), 1 - mover[i],
                                             packed, H, top_k, &nrep);
    out_score[i] = (float)((double)alpha * (double)S[i] - (double)beta * (double)W);
    if (out_replies) out_replies[i] = nrep;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* environment (src/environments/backgammon_env.py:92-329)                                    */
/* ------------------------------------------------------------------------------------------ */

struct bgo_env {
  bgo_board board;
  int player;
  int roll[2];
  int game_over;
  int win_type;
  int close_out_given[2], prime_given[2];
  int max_legal_moves;
  int num_moves;
  bgo_fullmove* moves;
  bgo_board* after;
  const uint8_t* tape;
  int64_t tape_len, tape_pos;
  uint64_t rng;
};

static uint64_t splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

bgo_env* bgo_env_create(int max_legal_moves, const uint8_t* tape, int64_t tape_len, uint64_t seed) {
  bgo_env* e = (bgo_env*)calloc(1, sizeof(bgo_env));
  e->max_legal_moves = max_legal_moves;
  e->moves = (bgo_fullmove*)malloc(sizeof(bgo_fullmove) * BGO_MAX_MOVES);
  e->after = (bgo_board*)malloc(sizeof(bgo_board) * BGO_MAX_MOVES);
  e->tape = tape;
  e->tape_len = tape_len;
  e->rng = seed;
  bgo_initial_board(&e->board);
  return e;
}

void bgo_env_destroy(bgo_env* e) {
  if (!e) return;
  free(e->moves);
  free(e->after);
  free(e);
}

/* :310-311 */
static void env_roll(bgo_env* e) {
  if (e->tape) {
    int64_t p = e->tape_pos < e->tape_len ? e->tape_pos : e->tape_len - 1;
    e->roll[0] = e->tape[2 * p];
    e->roll[1] = e->tape[2 * p + 1];
    e->tape_pos += 1;
  } else {
    e->roll[0] = 1 + (int)(splitmix64(&e->rng) % 6);
    e->roll[1] = 1 + (int)(splitmix64(&e->rng) % 6);
  }
}

/* :223-272 (truncation to max_legal_moves keeps the first entries) */
static void env_update_legal_moves(bgo_env* e) {
  int n = bgo_get_all_possible_moves(e->player, &e->board, e->roll[0], e->roll[1], e->moves, e->after,
                                     BGO_MAX_MOVES);
  if (n > BGO_MAX_MOVES) n = BGO_MAX_MOVES;
  if (n > e->max_legal_moves) n = e->max_legal_moves;
  e->num_moves = n;
}

/* :92-128 */
void bgo_env_reset(bgo_env* e) {
  bgo_initial_board(&e->board);
  e->game_over = 0;
  e->win_type = 0;
  env_roll(e);
  while (e->roll[0] == e->roll[1]) env_roll(e);
  e->player = e->roll[0] < e->roll[1] ? 1 : 0;
  env_roll(e);
  while (e->roll[0] == e->roll[1]) env_roll(e);
  env_update_legal_moves(e);
  e->close_out_given[0] = e->close_out_given[1] = 0;
  e->prime_given[0] = e->prime_given[1] = 0;
}

/* :130-221 */
int bgo_env_step(bgo_env* e, int action, float* reward, int* info_bits) {
  *info_bits = 0;
  *reward = 0.0f;
  if (e->game_over) return 1;
  if (e->num_moves == 0) { /* :139-151 */
    e->player = 1 - e->player;
    env_roll(e);
    env_update_legal_moves(e);
    *info_bits = 1;
    return 0;
  }
  if (action < 0 || action >= e->num_moves) { /* :153-158 */
    *reward = -1.0f;
    *info_bits = 2;
    return 0;
  }
  bgo_board nb;
  bgo_execute_full_move(&e->board, e->player, &e->moves[action], &nb);
  e->board = nb;
  if (bgo_check_game_over(&e->board, e->player)) { /* :167-193 */
    if (bgo_check_backgammon(&e->board, e->player)) {
      *reward = 2.5f;
      e->win_type = 3;
    } else if (bgo_check_gammon(&e->board, e->player)) {
      *reward = 2.0f;
      e->win_type = 2;
    } else {
      *reward = 1.0f;
      e->win_type = 1;
    }
    e->game_over = 1;
    *info_bits = 16;
    return 1;
  }
  float r = 0.0f; /* :195-218 */
  if (bgo_is_closed_out(&e->board, e->player) && !e->close_out_given[e->player]) {
    r += 0.30f;
    e->close_out_given[e->player] = 1;
    *info_bits |= 4;
  }
  if (bgo_made_five_prime(&e->board, e->player) && !e->prime_given[e->player]) {
    r += 0.20f;
    e->prime_given[e->player] = 1;
    *info_bits |= 8;
  }
  *reward = r;
  e->player = 1 - e->player;
  env_roll(e);
  env_update_legal_moves(e);
  return 0;
}

int bgo_env_num_moves(const bgo_env* e) { return e->num_moves; }
int bgo_env_player(const bgo_env* e) { return e->player; }
int bgo_env_win_type(const bgo_env* e) { return e->win_type; }
void bgo_env_roll(const bgo_env* e, int* d0, int* d1) {
  *d0 = e->roll[0];
  *d1 = e->roll[1];
}
const bgo_board* bgo_env_board(const bgo_env* e) { return &e->board; }
const bgo_board* bgo_env_afterstates(const bgo_env* e) { return e->after; }
const bgo_fullmove* bgo_env_moves(const bgo_env* e) { return e->moves; }
int64_t bgo_env_tape_pos(const bgo_env* e) { return e->tape_pos; }

/* ------------------------------------------------------------------------------------------ */
/* episode loop (src/multi/worker.py:78-174)                                                  */
/* ------------------------------------------------------------------------------------------ */

void bgo_play_episode(bgo_env* e, const float* packed, int H, float temperature, uint64_t* rng_state,
                      int max_steps, bgo_episode_stats* st, int32_t* tr_nmoves, int32_t* tr_action,
                      uint8_t* tr_roll, uint8_t* tr_player, float* tr_v, float* tr_vnext, float* tr_reward,
                      int8_t* tr_after_boards) {
  memset(st, 0, sizeof(*st));
  float* vals = (float*)malloc(sizeof(float) * (BGO_MAX_MOVES + 1));
  float* z = (float*)malloc(sizeof(float) * (size_t)H);
  bgo_env_reset(e);
  int done = 0, step = 0, nd = 0;
  while (!done && step < max_steps) {
    int n = e->num_moves;
    float reward;
    int info;
    if (n == 0) { /* :103-113 pass: no experience */
      done = bgo_env_step(e, -1, &reward, &info);
      st->n_passes += 1;
      step += 1;
      continue;
    }
    /* one forward over [obs; afterstates] (:116-125): obs flag = player to move, afterstate flag = mover */
    eval_board_f32(packed, H, &e->board, e->player, z, &vals[0]);
    for (int i = 0; i < n; ++i) eval_board_f32(packed, H, &e->after[i], e->player, z, &vals[1 + i]);
    st->n_afterstates += n;
    int a = 0;
    if (temperature <= 0.0f) { /* greedy: torch.argmax, lowest index on ties */
      float best = vals[1];
      for (int i = 1; i < n; ++i)
        if (vals[1 + i] > best) {
          best = vals[1 + i];
          a = i;
        }
    } else { /* :137-143 softmax(V/T) + Categorical sample */
      float m = vals[1] / temperature;
      for (int i = 1; i < n; ++i) {
        float t = vals[1 + i] / temperature;
        if (t > m) m = t;
      }
      double sum = 0.0;
      for (int i = 0; i < n; ++i) sum += exp((double)(vals[1 + i] / temperature - m));
      double u = (double)(splitmix64(rng_state) >> 11) * (1.0 / 9007199254740992.0) * sum;
      double acc = 0.0;
      a = n - 1;
      for (int i = 0; i < n; ++i) {
        acc += exp((double)(vals[1 + i] / temperature - m));
        if (u < acc) {
          a = i;
          break;
        }
      }
    }
    int player = e->player;
    if (tr_nmoves) tr_nmoves[nd] = n;
    if (tr_action) tr_action[nd] = a;
    if (tr_roll) {
      tr_roll[2 * nd] = (uint8_t)e->roll[0];
      tr_roll[2 * nd + 1] = (uint8_t)e->roll[1];
    }
    if (tr_player) tr_player[nd] = (uint8_t)player;
    if (tr_v) tr_v[nd] = vals[0];
    if (tr_vnext) tr_vnext[nd] = vals[1 + a];
    done = bgo_env_step(e, a, &reward, &info);
    if (tr_reward) tr_reward[nd] = reward;
    if (tr_after_boards) memcpy(tr_after_boards + (size_t)nd * BGO_BOARD_BYTES, &e->board, BGO_BOARD_BYTES);
    st->total_reward += reward;
    if (done) st->winner = player;
    nd += 1;
    step += 1;
  }
  st->n_steps = step;
  st->n_decisions = nd;
  st->win_type = e->win_type;
  free(vals);
  free(z);
}

int64_t bgo_selfplay_bench(const float* packed, int H, float temperature, int64_t n_games, uint64_t seed,
                           int nthreads, int64_t* out_steps, int64_t* out_decisions) {
  nthreads = resolve_threads(nthreads);
  int64_t after = 0, steps = 0, decs = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : after, steps, decs)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    bgo_env* e = bgo_env_create(500, NULL, 0, seed * 1315423911ull + (uint64_t)tid * 2654435761ull + 1);
    uint64_t rng = seed ^ (0xabcdef12345ull * (uint64_t)(tid + 1));
#pragma omp for schedule(dynamic, 1)
    for (int64_t g = 0; g < n_games; ++g) {
      bgo_episode_stats st;
      bgo_play_episode(e, packed, H, temperature, &rng, 300, &st, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
      after += st.n_afterstates;
      steps += st.n_steps;
      decs += st.n_decisions;
    }
    bgo_env_destroy(e);
  }
  if (out_steps) *out_steps = steps;
  if (out_decisions) *out_decisions = decs;
  return after;
}

int64_t bgo_movegen_eval_bench(const int8_t* boards, const uint8_t* players, const uint8_t* rolls, int64_t B,
                               const float* packed, int H, int nthreads, double* out_vsum) {
  nthreads = resolve_threads(nthreads);
  int64_t total = 0;
  double vsum = 0.0;
#pragma omp parallel num_threads(nthreads) reduction(+ : total, vsum)
  {
    bgo_board* bd = (bgo_board*)malloc(sizeof(bgo_board) * BGO_MAX_MOVES);
    float* z = (float*)malloc(sizeof(float) * (size_t)H);
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < B; ++i) {
      int n = bgo_get_all_possible_moves(players[i], (const bgo_board*)(boards + i * BGO_BOARD_BYTES), rolls[2 * i],
                                         rolls[2 * i + 1], NULL, bd, BGO_MAX_MOVES);
      if (n > BGO_MAX_MOVES) n = BGO_MAX_MOVES;
      for (int k = 0; k < n; ++k) {
        float v;
        eval_board_f32(packed, H, &bd[k], players[i], z, &v);
        vsum += v;
      }
      total += n;
    }
    free(bd);
    free(z);
  }
  if (out_vsum) *out_vsum = vsum;
  return total;
}

/* ------------------------------------------------------------------------------------------ */
/* synthetic position set                                  */
/* ------------------------------------------------------------------------------------------ */

/* One position sampled uniformly from the plies of a uniform-random-legal-move playout from the
 * initial board; playout g is keyed by (seed, g) only, so the set is independent of threading. */
void bgo_random_positions(int64_t n, uint64_t seed, int8_t* out_boards, uint8_t* out_players, int nthreads) {
  nthreads = resolve_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    bgo_env* e = bgo_env_create(500, NULL, 0, 0);
    bgo_board* traj = (bgo_board*)malloc(sizeof(bgo_board) * 1024);
    uint8_t* trajp = (uint8_t*)malloc(1024);
#pragma omp for schedule(dynamic, 64)
    for (int64_t g = 0; g < n; ++g) {
      uint64_t s = seed * 0x9e3779b97f4a7c15ull + (uint64_t)g * 0xd1b54a32d192ed03ull + 0x1234567ull;
      e->rng = splitmix64(&s);
      uint64_t r2 = splitmix64(&s);
      bgo_env_reset(e);
      int len = 0;
      traj[len] = e->board;
      trajp[len++] = (uint8_t)e->player;
      for (int step = 0; step < 1000 && len < 1024; ++step) {
        int nm = e->num_moves;
        int a = nm == 0 ? -1 : (int)(splitmix64(&r2) % (uint64_t)nm);
        float rew;
        int info;
        if (bgo_env_step(e, a, &rew, &info)) break;
        traj[len] = e->board;
        trajp[len++] = (uint8_t)e->player;
      }
      int pick = (int)(splitmix64(&r2) % (uint64_t)len);
      memcpy(out_boards + g * BGO_BOARD_BYTES, &traj[pick], BGO_BOARD_BYTES);
      out_players[g] = trajp[pick];
    }
    free(traj);
    free(trajp);
    bgo_env_destroy(e);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* learner: Trainer.update (src/agents/trainer.py:48-166)                                     */
/* ------------------------------------------------------------------------------------------ */

/* The reference processes the episodes of a batch SEQUENTIALLY: for each episode one forward pass over its
 * T observations (trainer.py:104-108), TD(0) targets r_t + gamma * V(x_{t+1}).detach() with the last target = r_T
 * (:110-115), loss = mse (:118), backward (:121-122), clip_grad_norm_(1.0) (:125-128), the logged gradient norm of the
 * clipped gradients (:131-136) and one torch.optim.Adam step (lr 1e-3, betas (0.9, 0.999), eps 1e-8, :27-29,139).
 * The arithmetic (matmul, mse, clip, Adam) is PyTorch's; its fp32 summation order is unspecified, so this
 * restatement accumulates every gradient sum in double and rounds once to fp32 (the neutral anchor), and keeps
 * parameters and Adam moments in fp32 with torch's operation order (torch/optim/adam.py _single_tensor_adam:
 * lerp, mul+addcmul, sqrt/bias_correction2_sqrt + eps, addcdiv with step_size = lr / bias_correction1). */
struct bgo_learner {
  int H;
  int64_t n_params;
  float* p; /* packed [W1t(198,H) | b1 | w2 | b2] */
  float* m;
  float* v;
  int64_t step;
  float lr, gamma, grad_clip;
};

bgo_learner* bgo_learner_create(const float* packed, int H, float lr, float gamma, float grad_clip) {
  bgo_learner* L = (bgo_learner*)calloc(1, sizeof(bgo_learner));
  L->H = H;
  L->n_params = (int64_t)(BGO_NFEAT + 2) * H + 1;
  L->p = (float*)malloc(sizeof(float) * (size_t)L->n_params);
  L->m = (float*)calloc((size_t)L->n_params, sizeof(float));
  L->v = (float*)calloc((size_t)L->n_params, sizeof(float));
  memcpy(L->p, packed, sizeof(float) * (size_t)L->n_params);
  L->lr = lr;
  L->gamma = gamma;
  L->grad_clip = grad_clip;
  return L;
}

void bgo_learner_destroy(bgo_learner* L) {
  if (!L) return;
  free(L->p);
  free(L->m);
  free(L->v);
  free(L);
}

void bgo_learner_get(const bgo_learner* L, float* packed, float* m, float* v, int64_t* step) {
  if (packed) memcpy(packed, L->p, sizeof(float) * (size_t)L->n_params);
  if (m) memcpy(m, L->m, sizeof(float) * (size_t)L->n_params);
  if (v) memcpy(v, L->v, sizeof(float) * (size_t)L->n_params);
  if (step) *step = L->step;
}

/* sub-range [lo,hi) of the packed vector is one torch parameter: sum of squares in double */
static double sumsq(const double* g, int64_t lo, int64_t hi) {
  double s = 0.0;
  for (int64_t i = lo; i < hi; ++i) s += g[i] * g[i];
  return s;
}

void bgo_learner_update(bgo_learner* L, const int8_t* obs_boards, const uint8_t* obs_flags, const float* reward,
                        const int64_t* ep_offsets, int64_t n_eps, float* out_metrics) {
  const int H = L->H;
  const int64_t NP = L->n_params;
  const int64_t oB1 = (int64_t)BGO_NFEAT * H, oW2 = oB1 + H, oB2 = oW2 + H;
  double* g = (double*)malloc(sizeof(double) * (size_t)NP);
  for (int64_t e = 0; e < n_eps; ++e) {
    const int64_t lo = ep_offsets[e], T = ep_offsets[e + 1] - lo;
    float* met = out_metrics ? out_metrics + e * BGO_LEARNER_NMETRICS : NULL;
    if (T <= 0) { /* the reference would fail on torch.stack([]); never produced by the worker (worker.py:149-158) */
      if (met) memset(met, 0, sizeof(float) * BGO_LEARNER_NMETRICS);
      continue;
    }
    float* x = (float*)malloc(sizeof(float) * (size_t)T * BGO_NFEAT);
    double* h = (double*)malloc(sizeof(double) * (size_t)T * (size_t)H);
    float* Y = (float*)malloc(sizeof(float) * (size_t)T);
    float* tgt = (float*)malloc(sizeof(float) * (size_t)T);
    const float* W1t = L->p;
    const float* b1 = L->p + oB1;
    const float* w2 = L->p + oW2;
    /* forward (policy_network.py:53-70) */
    for (int64_t t = 0; t < T; ++t) {
      bgo_features((const bgo_board*)(obs_boards + (lo + t) * BGO_BOARD_BYTES), obs_flags[lo + t], x + t * BGO_NFEAT);
      double* ht = h + t * H;
      for (int j = 0; j < H; ++j) ht[j] = (double)b1[j];
      for (int f = 0; f < BGO_NFEAT; ++f) {
        float xf = x[t * BGO_NFEAT + f];
        if (xf != 0.0f)
          for (int j = 0; j < H; ++j) ht[j] += (double)xf * (double)W1t[(int64_t)f * H + j];
      }
      double y = (double)L->p[oB2];
      for (int j = 0; j < H; ++j) {
        ht[j] = 1.0 / (1.0 + exp(-ht[j]));
        y += (double)w2[j] * ht[j];
      }
      Y[t] = (float)y;
    }
    /* targets (trainer.py:110-115) and loss (:118) */
    double loss = 0.0, tdabs = 0.0, ysum = 0.0, rsum = 0.0;
    for (int64_t t = 0; t < T; ++t) {
      tgt[t] = reward[lo + t];
      if (t + 1 < T) tgt[t] = tgt[t] + L->gamma * Y[t + 1];
      double d = (double)Y[t] - (double)tgt[t];
      loss += d * d;
      tdabs += fabs((double)tgt[t] - (double)Y[t]);
      ysum += (double)Y[t];
      rsum += (double)reward[lo + t];
    }
    loss /= (double)T;
    /* backward of mean((Y - tgt)^2) through value_head, sigmoid, fc1 */
    memset(g, 0, sizeof(double) * (size_t)NP);
    for (int64_t t = 0; t < T; ++t) {
      double dY = 2.0 * ((double)Y[t] - (double)tgt[t]) / (double)T;
      const double* ht = h + t * H;
      g[oB2] += dY;
      for (int j = 0; j < H; ++j) {
        g[oW2 + j] += dY * ht[j];
        double dz = dY * (double)w2[j] * ht[j] * (1.0 - ht[j]);
        g[oB1 + j] += dz;
        for (int f = 0; f < BGO_NFEAT; ++f) {
          float xf = x[t * BGO_NFEAT + f];
          if (xf != 0.0f) g[(int64_t)f * H + j] += dz * (double)xf;
        }
      }
    }
    /* torch.nn.utils.clip_grad_norm_ (trainer.py:125-128): norm of the per-parameter norms, coef = max_norm / (norm + 1e-6)
     * clamped to 1, gradients always multiplied by it */
    double total = sqrt(sumsq(g, 0, oB1) + sumsq(g, oB1, oW2) + sumsq(g, oW2, oB2) + sumsq(g, oB2, NP));
    float coef = 1.0f;
    if (L->grad_clip > 0.0f) {
      coef = L->grad_clip / ((float)total + 1e-6f);
      if (coef > 1.0f) coef = 1.0f;
    }
    /* Adam (torch/optim/adam.py, defaults of trainer.py:27-29) */
    L->step += 1;
    const double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
    const double bc1 = 1.0 - pow(beta1, (double)L->step), bc2 = 1.0 - pow(beta2, (double)L->step);
    const float step_size = (float)((double)L->lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    double gn2 = 0.0;
    for (int64_t i = 0; i < NP; ++i) {
      float gi = (float)g[i] * coef;
      gn2 += (double)gi * (double)gi;
      L->m[i] = L->m[i] + (float)(1.0 - beta1) * (gi - L->m[i]);
      L->v[i] = L->v[i] * (float)beta2 + (float)(1.0 - beta2) * gi * gi;
      float denom = sqrtf(L->v[i]) / bc2_sqrt + (float)eps;
      L->p[i] = L->p[i] - step_size * (L->m[i] / denom);
    }
    if (met) {
      met[0] = (float)loss;                 /* loss.item()                       trainer.py:148 */
      met[1] = (float)(tdabs / (double)T);  /* TD_error.abs().mean()             :142-144 */
      met[2] = (float)sqrt(gn2);            /* norm of the clipped gradients     :131-136 */
      met[3] = (float)(ysum / (double)T);   /* Y_values.mean()                   :149 */
      met[4] = (float)rsum;                 /* rewards.sum()                     :150 */
      met[5] = (float)T;                    /* seq_len                           :151 */
    }
    free(x);
    free(h);
    free(Y);
    free(tgt);
  }
  free(g);
}
