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
