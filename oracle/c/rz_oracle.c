/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's AlphaZero search on k-in-a-row
 * boards, for parity checks at sizes the Python oracle cannot finish (8192 games x 800 playouts).
 *
 * Follows, function by function (paths under /root/reference):
 *   board / step / legal list     rlzero/games/gomoku/gomoku_env.py:19-70
 *   has_a_winner                  rlzero/games/gomoku/gomoku_env.py:116-170
 *   game_end_winner               rlzero/games/gomoku/gomoku_env.py:196-203
 *   TreeNode select / uct_value   rlzero/mcts/node.py:32-42,75-88   (PUCT: deepmind_mcts.py:149-151)
 *   TreeNode expand               rlzero/mcts/node.py:44-73 (no noise)
 *   update_recursive              rlzero/mcts/node.py:119-144
 *   _playout / simulate           rlzero/mcts/alphazero_mcts.py:42-94
 *   update_with_move              rlzero/mcts/alphazero_mcts.py:96-103
 * with the closed-form evaluators of oracle/evaluators.py as policy_value_fn.  It shares no code with
 * the CUDA library and none with the Python oracle; tests/test_oracle_c.py pins it against the Python
 * restatement and the golden vectors generated from the live reference (tests/golden/mcts_kat.json).
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: Python evaluates W/n + c*sqrt(ln(Np)/n) with
 * separately rounded operations) -fopenmp -shared -fPIC; see oracle/build_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXC 361

typedef struct {
  int rows, cols, k;
  int gravity;            /* 0: an action is a square (Gomoku, gomoku_env.py); 1: an action is a column and the stone
                             drops to the lowest empty row (Connect Four: oracle/pyoracle.py ConnectFourBoard, the CPU
                             definition of RZ_GAME_CONNECT4 -- the reference has no such game) */
  int8_t cell[MAXC];      /* -1 empty, 0 / 1; square index r * cols + c (row 0 at the bottom for gravity games) */
  int8_t height[19];      /* gravity: stones in each column */
  int stones, last_move, to_move;
} board_t;

typedef struct node_s {
  struct node_s* parent;
  struct node_s** child;  /* [A] by action, NULL = no such child; allocated on expand */
  int expanded;
  int n;                  /* explore_count */
  double w;               /* total_reward */
  double prior;
} node_t;

/* simple arena so a search frees everything at once */
typedef struct { char* base; size_t used, cap; } arena_t;
static void* arena_alloc(arena_t* a, size_t sz) {
  sz = (sz + 15) & ~(size_t)15;
  if (a->used + sz > a->cap) {
    size_t nc = a->cap * 2 + sz;
    /* chunks are never moved: allocate a new block and chain it through the first bytes */
    char* nb = (char*)malloc(nc + 16);
    if (!nb) return NULL;
    *(char**)nb = a->base;
    a->base = nb; a->used = 16; a->cap = nc + 16;
  }
  void* p = a->base + a->used;
  a->used += sz;
  return p;
}
static void arena_free(arena_t* a) {
  char* b = a->base;
  while (b) { char* nx = *(char**)b; free(b); b = nx; }
  a->base = NULL;
}
static int arena_init(arena_t* a, size_t cap) {
  a->base = (char*)malloc(cap + 16);
  if (!a->base) return -1;
  *(char**)a->base = NULL;
  a->used = 16; a->cap = cap + 16;
  return 0;
}

static void board_reset(board_t* b, int rows, int cols, int k, int gravity) {
  b->rows = rows; b->cols = cols; b->k = k; b->gravity = gravity;
  memset(b->cell, -1, sizeof(b->cell));
  memset(b->height, 0, sizeof(b->height));
  b->stones = 0; b->last_move = -1; b->to_move = 0;
}
static int board_actions(const board_t* b) { return b->gravity ? b->cols : b->rows * b->cols; }
static int board_legal(const board_t* b, int a) { return b->gravity ? b->height[a] < b->rows : b->cell[a] < 0; }
static int board_n_legal(const board_t* b) {
  if (!b->gravity) return b->rows * b->cols - b->stones;
  int n = 0;
  for (int c = 0; c < b->cols; ++c) n += b->height[c] < b->rows;
  return n;
}
static void board_step(board_t* b, int a) {
  int cell = a;
  if (b->gravity) { cell = b->height[a] * b->cols + a; b->height[a] += 1; }
  b->cell[cell] = (int8_t)b->to_move;
  b->stones += 1; b->last_move = cell; b->to_move ^= 1;
}
/* gomoku_env.py:116-170: for every stone, the four directions with the reference's edge guards */
static int board_winner(const board_t* b) {
  const int n = b->cols, nr = b->rows, k = b->k;     /* square boards: n == nr, the reference's width == height */
  if (b->stones < 2 * k - 1) return -1;
  for (int m = 0; m < nr * n; ++m) {
    const int p = b->cell[m];
    if (p < 0) continue;
    const int h = m / n, w = m % n;
    int ok;
    if (w <= n - k) { ok = 1; for (int i = 1; i < k && ok; ++i) ok = b->cell[m + i] == p; if (ok) return p; }
    if (h <= nr - k) { ok = 1; for (int i = 1; i < k && ok; ++i) ok = b->cell[m + i * n] == p; if (ok) return p; }
    if (w <= n - k && h <= nr - k) { ok = 1; for (int i = 1; i < k && ok; ++i) ok = b->cell[m + i * (n + 1)] == p; if (ok) return p; }
    if (w >= k - 1 && h <= nr - k) { ok = 1; for (int i = 1; i < k && ok; ++i) ok = b->cell[m + i * (n - 1)] == p; if (ok) return p; }
  }
  return -1;
}
/* game_end_winner: 1 = ended; *winner = player or -1 (tie) */
static int board_end(const board_t* b, int* winner) {
  *winner = board_winner(b);
  if (*winner >= 0) return 1;
  return b->stones >= b->rows * b->cols;
}

/* oracle/evaluators.py */
static uint32_t board_hash(const board_t* b) {
  uint32_t h = 0;
  for (int m = 0; m < b->rows * b->cols; ++m)
    if (b->cell[m] >= 0) h += (uint32_t)(m + 1) * (uint32_t)(m + 1) * (3u + 4u * (uint32_t)b->cell[m]);
  h += 7u * (uint32_t)(b->last_move + 1);
  return h * 2654435761u;
}
static double eval_value(const board_t* b, int eval_id) {
  if (eval_id == 1) return (double)((17 * b->stones + 31 * (b->last_move + 1)) % 13 - 6) / 8.0;
  if (eval_id == 2) return (double)((int)((board_hash(b) >> 16) % 129u) - 64) / 64.0;
  return 0.0;
}
static double eval_prior(const board_t* b, int eval_id, int action, int n_legal, uint32_t h) {
  if (eval_id == 2) return (double)(((uint32_t)action * 29u + (h >> 8)) % 32u + 1u) / 256.0;
  return (double)(1.0f / (float)n_legal);    /* np.float32(1) / np.float32(len(legal)) */
}

static double node_score(const node_t* c, double cpuct, int rule) {
  if (rule == 1)   /* deepmind_mcts.py:149-151 */
    return (c->n ? c->w / (double)c->n : 0.0) + cpuct * c->prior * sqrt((double)c->parent->n) / (double)(c->n + 1);
  if (c->parent->n == 0 || c->n == 0) return INFINITY;      /* node.py:76-80 */
  return c->w / (double)c->n + cpuct * sqrt(log((double)c->parent->n) / (double)c->n);
}

static int playout(arena_t* ar, node_t* root, board_t b, int A, double cpuct, int rule, int eval_id) {
  node_t* node = root;
  while (node->expanded) {
    int best_a = -1; double best_s = 0.0;
    for (int a = 0; a < A; ++a) {                 /* children in ascending action order, first maximum */
      node_t* c = node->child[a];
      if (!c) continue;
      const double s = node_score(c, cpuct, rule);
      if (best_a < 0 || s > best_s) { best_a = a; best_s = s; }
    }
    if (best_a < 0) return -2;                    /* ValueError('Node has no children.') */
    node = node->child[best_a];
    board_step(&b, best_a);
  }
  double v = eval_value(&b, eval_id);             /* evaluated on terminal leaves too (:59) */
  int winner;
  if (!board_end(&b, &winner)) {
    node->child = (node_t**)arena_alloc(ar, sizeof(node_t*) * (size_t)A);
    if (!node->child) return -1;
    const uint32_t h = eval_id == 2 ? board_hash(&b) : 0u;
    const int n_legal = board_n_legal(&b);
    for (int a = 0; a < A; ++a) {
      node->child[a] = NULL;
      if (!board_legal(&b, a)) continue;
      node_t* c = (node_t*)arena_alloc(ar, sizeof(node_t));
      if (!c) return -1;
      c->parent = node; c->child = NULL; c->expanded = 0; c->n = 0; c->w = 0.0;
      c->prior = eval_prior(&b, eval_id, a, n_legal, h);
      node->child[a] = c;
    }
    node->expanded = 1;
  } else if (winner < 0) {
    v = 0.0;
  } else {
    v = winner == b.to_move ? 1.0 : -1.0;
  }
  v = -v;                                          /* update_recursive(-leaf_value) */
  for (node_t* p = node; p; p = p->parent) { p->n += 1; p->w += v; v = -v; }
  return 0;
}

/* One game: play `moves`, then for each of n_searches: n_playout playouts, report the root children's
 * visits / value sums, then commit follow[j] (keeping the subtree, update_with_move) if j < n_follow.
 * visits/w: [n_searches][A]; root_n/root_w: [n_searches].  Returns 0, or <0 on error. */
int rzo_search_game(int size, int k, const int32_t* moves, int n_moves, int n_playout, double cpuct, int rule,
                    int eval_id, int n_searches, const int32_t* follow, int32_t* visits, double* w,
                    int32_t* root_n, double* root_w) {
  if (size < 1 || size * size > MAXC) return -3;
  const int A = size * size;
  board_t b;
  board_reset(&b, size, size, k, 0);
  for (int i = 0; i < n_moves; ++i) {
    if (moves[i] < 0 || moves[i] >= A || b.cell[moves[i]] >= 0) return -4;
    board_step(&b, moves[i]);
  }
  arena_t ar;
  if (arena_init(&ar, (size_t)1 << 22)) return -1;
  node_t* root = (node_t*)arena_alloc(&ar, sizeof(node_t));
  memset(root, 0, sizeof(*root));
  root->prior = 1.0;
  int rc = 0;
  for (int j = 0; j < n_searches && rc == 0; ++j) {
    for (int i = 0; i < n_playout && rc == 0; ++i) rc = playout(&ar, root, b, A, cpuct, rule, eval_id);
    if (rc) break;
    for (int a = 0; a < A; ++a) {
      node_t* c = root->expanded ? root->child[a] : NULL;
      visits[(size_t)j * A + a] = c ? c->n : 0;
      w[(size_t)j * A + a] = c ? c->w : 0.0;
    }
    root_n[j] = root->n; root_w[j] = root->w;
    if (j + 1 < n_searches) {
      const int m = follow[j];
      if (m < 0 || m >= A || b.cell[m] >= 0) { rc = -4; break; }
      board_step(&b, m);
      if (root->expanded && root->child[m]) { root = root->child[m]; root->parent = NULL; }
      else { root = (node_t*)arena_alloc(&ar, sizeof(node_t)); memset(root, 0, sizeof(*root)); root->prior = 1.0; }
    }
  }
  arena_free(&ar);
  return rc;
}

/* GomokuEnv.step / game_end_winner (gomoku_env.py:49-70,196-203) over whole games: play moves[g][0..] until the game
 * is over; end_ply[g] = number of moves played (n_moves[g] if it never ended), winner[g] = the winner, -1 for a tie
 * or an unfinished game, ended[g] = 1 if it ended. */
int rzo_replay_games(int G, int size, int k, const int32_t* moves, const int32_t* n_moves, int max_moves,
                     int32_t* end_ply, int32_t* winner, int32_t* ended) {
  if (size < 1 || size * size > MAXC) return -3;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 64)
  for (int g = 0; g < G; ++g) {
    board_t b;
    board_reset(&b, size, size, k, 0);
    int t = 0, w = -1, e = 0;
    for (; t < n_moves[g]; ++t) {
      const int a = moves[(size_t)g * max_moves + t];
      if (a < 0 || a >= size * size || b.cell[a] >= 0) {
#pragma omp atomic write
        bad = -4;
        break;
      }
      board_step(&b, a);
      if (board_end(&b, &w)) { e = 1; t += 1; break; }
    }
    end_ply[g] = t; winner[g] = e ? w : -1; ended[g] = e;
  }
  return bad;
}

/* Two searches per game with the subtree kept in between (update_with_move, alphazero_mcts.py:96-103): after the
 * first search the most visited root child is played (lowest action on ties, numpy argmax), then n_playout more
 * playouts from the re-rooted tree.  Outputs the SECOND stage and the move played; games whose position after the move
 * is over report move and zeros. */
int rzo_search_batch_reuse(int G, int size, int k, const int32_t* moves, const int32_t* n_moves, int max_moves,
                           int n_playout, double cpuct, int rule, int eval_id, int32_t* move_out, int32_t* visits,
                           double* w, int32_t* root_n, double* root_w) {
  const int A = size * size;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int g = 0; g < G; ++g) {
    int32_t* v1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)A * 2);
    double* w1 = (double*)malloc(sizeof(double) * (size_t)A * 2);
    int32_t rn[2]; double rw[2]; int32_t follow[1];
    int rc = (v1 && w1) ? rzo_search_game(size, k, moves + (size_t)g * max_moves, n_moves[g], n_playout, cpuct, rule,
                                          eval_id, 1, NULL, v1, w1, rn, rw) : -1;
    if (rc == 0) {
      int best = 0;
      for (int a = 1; a < A; ++a) if (v1[a] > v1[best]) best = a;
      move_out[g] = best;
      follow[0] = best;
      /* is the game over after that move?  replay on a scratch board */
      board_t b; board_reset(&b, size, size, k, 0);
      for (int i = 0; i < n_moves[g]; ++i) board_step(&b, moves[(size_t)g * max_moves + i]);
      board_step(&b, best);
      int winner;
      if (board_end(&b, &winner)) {
        for (int a = 0; a < A; ++a) { visits[(size_t)g * A + a] = 0; w[(size_t)g * A + a] = 0.0; }
        root_n[g] = 0; root_w[g] = 0.0;
      } else {
        rc = rzo_search_game(size, k, moves + (size_t)g * max_moves, n_moves[g], n_playout, cpuct, rule, eval_id, 2,
                             follow, v1, w1, rn, rw);
        if (rc == 0) {
          memcpy(visits + (size_t)g * A, v1 + A, sizeof(int32_t) * (size_t)A);
          memcpy(w + (size_t)g * A, w1 + A, sizeof(double) * (size_t)A);
          root_n[g] = rn[1]; root_w[g] = rw[1];
        }
      }
    }
    free(v1); free(w1);
    if (rc) {
#pragma omp atomic write
      bad = rc;
    }
  }
  return bad;
}

/* ---- the product's opt-in leaf-parallel wave (include/rlzero_b200.h, rz_tree_desc.leaves_per_tree) -------------
 * NOT a reference algorithm (the reference has no virtual loss: PARITY UNPINNED for this mode); restated here, as in
 * oracle/pyoracle.py Search.wave, so that the kernels can be checked at full size.  `budget` descents, each followed
 * by one virtual visit and -vl on the value sum of every node of its path (+1 visit on the root); then the virtual
 * statistics are taken off in reverse order (saved sums restored), and the leaves are expanded / backed up in order;
 * a leaf reached twice is expanded once and backed up twice. */
typedef struct { node_t* node; double w; } undo_t;

static int expand_and_backup(arena_t* ar, node_t* node, board_t* b, int A, int eval_id) {
  double v = eval_value(b, eval_id);
  int winner;
  if (!board_end(b, &winner)) {
    if (!node->expanded) {
      node->child = (node_t**)arena_alloc(ar, sizeof(node_t*) * (size_t)A);
      if (!node->child) return -1;
      const uint32_t h = eval_id == 2 ? board_hash(b) : 0u;
      const int n_legal = board_n_legal(b);
      for (int a = 0; a < A; ++a) {
        node->child[a] = NULL;
        if (!board_legal(b, a)) continue;
        node_t* c = (node_t*)arena_alloc(ar, sizeof(node_t));
        if (!c) return -1;
        c->parent = node; c->child = NULL; c->expanded = 0; c->n = 0; c->w = 0.0;
        c->prior = eval_prior(b, eval_id, a, n_legal, h);
        node->child[a] = c;
      }
      node->expanded = 1;
    }
  } else if (winner < 0) {
    v = 0.0;
  } else {
    v = winner == b->to_move ? 1.0 : -1.0;
  }
  v = -v;
  for (node_t* p = node; p; p = p->parent) { p->n += 1; p->w += v; v = -v; }
  return 0;
}

static int wave(arena_t* ar, node_t* root, const board_t* board, int A, double cpuct, int rule, int eval_id,
                int budget, double vl, node_t** leaf_node, board_t* leaf_board, undo_t* undo) {
  int n_undo = 0;
  for (int k = 0; k < budget; ++k) {
    board_t b = *board;
    node_t* node = root;
    const int first = n_undo;
    while (node->expanded) {
      int best_a = -1; double best_s = 0.0;
      for (int a = 0; a < A; ++a) {
        node_t* c = node->child[a];
        if (!c) continue;
        const double s = node_score(c, cpuct, rule);
        if (best_a < 0 || s > best_s) { best_a = a; best_s = s; }
      }
      if (best_a < 0) return -2;
      node = node->child[best_a];
      board_step(&b, best_a);
      undo[n_undo].node = node; n_undo += 1;
    }
    leaf_node[k] = node; leaf_board[k] = b;
    for (int i = first; i < n_undo; ++i) {
      node_t* nd = undo[i].node;
      undo[i].w = nd->w;
      nd->w = nd->n > 0 ? nd->w - vl : -vl;
      nd->n += 1;
    }
    root->n += 1;
  }
  for (int i = n_undo - 1; i >= 0; --i) { undo[i].node->w = undo[i].w; undo[i].node->n -= 1; }
  root->n -= budget;
  for (int k = 0; k < budget; ++k) {
    const int rc = expand_and_backup(ar, leaf_node[k], &leaf_board[k], A, eval_id);
    if (rc) return rc;
  }
  return 0;
}

/* one fresh search of n_playout playouts in waves of up to K leaves (the first wave of a fresh root takes one) */
static int search_game_vl(int rows, int cols, int gravity, int k, const int32_t* moves, int n_moves, int n_playout,
                          double cpuct, int rule, int eval_id, int K, double vl, int32_t* visits, double* w,
                          int32_t* root_n, double* root_w) {
  if (rows < 1 || cols < 1 || cols > 19 || rows * cols > MAXC || K < 1 || K > 256) return -3;
  board_t b;
  board_reset(&b, rows, cols, k, gravity);
  const int A = board_actions(&b);
  for (int i = 0; i < n_moves; ++i) {
    if (moves[i] < 0 || moves[i] >= A || !board_legal(&b, moves[i])) return -4;
    board_step(&b, moves[i]);
  }
  arena_t ar;
  if (arena_init(&ar, (size_t)1 << 22)) return -1;
  node_t* root = (node_t*)arena_alloc(&ar, sizeof(node_t));
  memset(root, 0, sizeof(*root));
  root->prior = 1.0;
  node_t** leaf_node = (node_t**)malloc(sizeof(node_t*) * (size_t)K);
  board_t* leaf_board = (board_t*)malloc(sizeof(board_t) * (size_t)K);
  undo_t* undo = (undo_t*)malloc(sizeof(undo_t) * (size_t)K * (size_t)(rows * cols + 1));
  int rc = (leaf_node && leaf_board && undo) ? 0 : -1;
  const int target = root->n + n_playout;
  while (rc == 0 && root->n < target) {
    if (K == 1) { rc = playout(&ar, root, b, A, cpuct, rule, eval_id); continue; }   /* the sequential reference search */
    int budget = target - root->n < K ? target - root->n : K;
    if (!root->expanded) budget = 1;
    rc = wave(&ar, root, &b, A, cpuct, rule, eval_id, budget, vl, leaf_node, leaf_board, undo);
  }
  if (rc == 0) {
    for (int a = 0; a < A; ++a) {
      node_t* c = root->expanded ? root->child[a] : NULL;
      visits[a] = c ? c->n : 0;
      w[a] = c ? c->w : 0.0;
    }
    *root_n = root->n; *root_w = root->w;
  }
  free(leaf_node); free(leaf_board); free(undo);
  arena_free(&ar);
  return rc;
}

int rzo_search_batch_vl(int G, int size, int k, const int32_t* moves, const int32_t* n_moves, int max_moves,
                        int n_playout, double cpuct, int rule, int eval_id, int K, double vl, int32_t* visits,
                        double* w, int32_t* root_n, double* root_w) {
  const int A = size * size;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int g = 0; g < G; ++g) {
    const int rc = search_game_vl(size, size, 0, k, moves + (size_t)g * max_moves, n_moves[g], n_playout, cpuct, rule,
                                  eval_id, K, vl, visits + (size_t)g * A, w + (size_t)g * A, root_n + g, root_w + g);
    if (rc) {
#pragma omp atomic write
      bad = rc;
    }
  }
  return bad;
}

/* ---- the reference's second search driver, DeepMindMCTS (rlzero/mcts/deepmind_mcts.py:65-175,384-646) -------------
 *   SearchNode.uct_value / puct_value with the outcome shortcut   :106-151
 *   sort_key / best_child                                         :153-175
 *   _apply_tree_policy (children created on a node's second visit, first maximum; no shuffle, no noise here) :477-528
 *   mcts_search (returns indexed by the player who moved, terminal outcomes, MCTS-Solver, early stop)       :554-646
 * returns_mode 0: GomokuEnv.returns() literally (gomoku_env.py:216-219 tests winner == 1 / == 2 although the players
 * are 0 / 1: a win of player 1 yields [1,-1], everything else [0,0]); 1: the zero-sum intent. */
typedef struct dnode_s {
  int action, player, n, n_children, has_outcome;
  int o[2];
  double prior, w;
  struct dnode_s* children;
} dnode_t;

static double dnode_score(const dnode_t* c, int parent_n, double uct_c, int rule) {
  if (c->has_outcome) return (double)c->o[c->player];
  if (rule == 1) return (c->n ? c->w / (double)c->n : 0.0) + uct_c * c->prior * sqrt((double)parent_n) / (double)(c->n + 1);
  if (c->n == 0) return INFINITY;
  return c->w / (double)c->n + uct_c * sqrt(log((double)parent_n) / (double)c->n);
}

static void board_returns(const board_t* b, int mode, int winner, int* r) {
  r[0] = 0; r[1] = 0;
  if (winner < 0) return;
  if (mode == 0) { if (winner == 1) { r[0] = 1; r[1] = -1; } return; }
  r[0] = winner == 0 ? 1 : -1; r[1] = -r[0];
  (void)b;
}

/* One search; outputs per action a: visits[a] (-1 = not a root child), w[a], outcome code[a] (0 none, else
 * 0x100 | (o0+1) | (o1+1) << 2), and root_n, root_w, root outcome code, best_child action. */
static int dm_search_game(int size, int k, const int32_t* moves, int n_moves, int sims, double uct_c, int rule,
                          int solve, int returns_mode, int eval_id, int32_t* visits, double* w, int32_t* outcome,
                          int32_t* root_n, double* root_w, int32_t* root_o, int32_t* best_a) {
  if (size < 1 || size * size > MAXC) return -3;
  const int A = size * size;
  board_t b0;
  board_reset(&b0, size, size, k, 0);
  for (int i = 0; i < n_moves; ++i) {
    if (moves[i] < 0 || moves[i] >= A || b0.cell[moves[i]] >= 0) return -4;
    board_step(&b0, moves[i]);
  }
  arena_t ar;
  if (arena_init(&ar, (size_t)1 << 22)) return -1;
  dnode_t* root = (dnode_t*)arena_alloc(&ar, sizeof(dnode_t));
  memset(root, 0, sizeof(*root));
  root->action = -1; root->player = b0.to_move; root->prior = 1.0;
  dnode_t** path = (dnode_t**)malloc(sizeof(dnode_t*) * (size_t)(A + 2));
  int rc = path ? 0 : -1;
  for (int s = 0; s < sims && rc == 0; ++s) {
    board_t work = b0;
    dnode_t* node = root;
    int depth = 0, winner = -1, ended;
    path[depth++] = node;
    while (!(ended = board_end(&work, &winner)) && node->n > 0) {
      if (!node->children) {
        const int nl = board_n_legal(&work);
        node->children = (dnode_t*)arena_alloc(&ar, sizeof(dnode_t) * (size_t)nl);
        if (!node->children) { rc = -1; break; }
        const uint32_t h = eval_id == 2 ? board_hash(&work) : 0u;
        int j = 0;
        for (int a = 0; a < A; ++a) {
          if (!board_legal(&work, a)) continue;
          dnode_t* c = &node->children[j++];
          memset(c, 0, sizeof(*c));
          c->action = a; c->player = work.to_move; c->prior = eval_prior(&work, eval_id, a, nl, h);
        }
        node->n_children = nl;
      }
      dnode_t* best = NULL; double best_s = 0.0;
      for (int j = 0; j < node->n_children; ++j) {
        const double sc = dnode_score(&node->children[j], node->n, uct_c, rule);
        if (!best || sc > best_s) { best = &node->children[j]; best_s = sc; }
      }
      board_step(&work, best->action);
      node = best;
      path[depth++] = node;
    }
    if (rc) break;
    double ret[2];
    int solved = 0;
    if (ended) {
      int r[2];
      board_returns(&work, returns_mode, winner, r);
      ret[0] = (double)r[0]; ret[1] = (double)r[1];
      dnode_t* leaf = path[depth - 1];
      leaf->has_outcome = 1; leaf->o[0] = r[0]; leaf->o[1] = r[1];
      solved = solve;
    } else {
      const double v = eval_value(&work, eval_id);
      ret[work.to_move] = v; ret[1 - work.to_move] = -v;
    }
    while (depth > 0) {
      dnode_t* nd = path[--depth];
      nd->w += ret[nd->player];
      nd->n += 1;
      if (solved && nd->children) {
        const int player = nd->children[0].player;
        dnode_t* best = NULL; int all_solved = 1;
        for (int j = 0; j < nd->n_children; ++j) {
          dnode_t* c = &nd->children[j];
          if (!c->has_outcome) all_solved = 0;
          else if (!best || c->o[player] > best->o[player]) best = c;
        }
        if (best && (all_solved || best->o[player] == 1)) { nd->has_outcome = 1; nd->o[0] = best->o[0]; nd->o[1] = best->o[1]; }
        else solved = 0;
      }
    }
    if (root->has_outcome) break;
  }
  if (rc == 0) {
    for (int a = 0; a < A; ++a) { visits[a] = -1; w[a] = 0.0; outcome[a] = 0; }
    dnode_t* best = NULL;
    for (int j = 0; j < root->n_children; ++j) {
      dnode_t* c = &root->children[j];
      visits[c->action] = c->n; w[c->action] = c->w;
      outcome[c->action] = c->has_outcome ? (0x100 | (c->o[0] + 1) | ((c->o[1] + 1) << 2)) : 0;
      if (!best) { best = c; continue; }
      /* sort_key (:153-171): (outcome[player] or 0, explore_count, total_reward), first maximum */
      const int ko = c->has_outcome ? c->o[c->player] : 0, kb = best->has_outcome ? best->o[best->player] : 0;
      if (ko > kb || (ko == kb && (c->n > best->n || (c->n == best->n && c->w > best->w)))) best = c;
    }
    *root_n = root->n; *root_w = root->w;
    *root_o = root->has_outcome ? (0x100 | (root->o[0] + 1) | ((root->o[1] + 1) << 2)) : 0;
    *best_a = best ? best->action : -1;
  }
  free(path);
  arena_free(&ar);
  return rc;
}

int rzo_dm_search_batch(int G, int size, int k, const int32_t* moves, const int32_t* n_moves, int max_moves, int sims,
                        double uct_c, int rule, int solve, int returns_mode, int eval_id, int32_t* visits, double* w,
                        int32_t* outcome, int32_t* root_n, double* root_w, int32_t* root_o, int32_t* best_a) {
  const int A = size * size;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int g = 0; g < G; ++g) {
    const int rc = dm_search_game(size, k, moves + (size_t)g * max_moves, n_moves[g], sims, uct_c, rule, solve,
                                  returns_mode, eval_id, visits + (size_t)g * A, w + (size_t)g * A,
                                  outcome + (size_t)g * A, root_n + g, root_w + g, root_o + g, best_a + g);
    if (rc) {
#pragma omp atomic write
      bad = rc;
    }
  }
  return bad;
}

/* Connect Four (gravity, rows x cols, actions = columns): K = 1 is the reference's sequential search over that game,
 * K > 1 the leaf-parallel wave.  visits / w: [G][cols]. */
int rzo_search_batch_c4(int G, int rows, int cols, int k, const int32_t* moves, const int32_t* n_moves, int max_moves,
                        int n_playout, double cpuct, int rule, int eval_id, int K, double vl, int32_t* visits,
                        double* w, int32_t* root_n, double* root_w) {
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int g = 0; g < G; ++g) {
    const int rc = search_game_vl(rows, cols, 1, k, moves + (size_t)g * max_moves, n_moves[g], n_playout, cpuct, rule,
                                  eval_id, K, vl, visits + (size_t)g * cols, w + (size_t)g * cols, root_n + g,
                                  root_w + g);
    if (rc) {
#pragma omp atomic write
      bad = rc;
    }
  }
  return bad;
}

/* G independent games in parallel over the host cores (OpenMP): moves [G][max_moves] with n_moves[G]. */
int rzo_search_batch(int G, int size, int k, const int32_t* moves, const int32_t* n_moves, int max_moves,
                     int n_playout, double cpuct, int rule, int eval_id, int32_t* visits, double* w,
                     int32_t* root_n, double* root_w) {
  const int A = size * size;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int g = 0; g < G; ++g) {
    const int rc = rzo_search_game(size, k, moves + (size_t)g * max_moves, n_moves[g], n_playout, cpuct, rule,
                                   eval_id, 1, NULL, visits + (size_t)g * A, w + (size_t)g * A, root_n + g,
                                   root_w + g);
    if (rc) {
#pragma omp atomic write
      bad = rc;
    }
  }
  return bad;
}
