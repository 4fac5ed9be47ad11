// rlzero_b200 -- a Go position held by one warp, one board row per lane.
//
// The reference plays Go through rlzero/games/go/go_env.py, a wrapper over pettingzoo's go_base
// (MiniGo's go.py; third-party, un-vendored -- see oracle/go_oracle.py for the restated rules and the
// call sites).  Lane r (< H) keeps row r of each colour as a bitmask, like the Gomoku board
// (rz_board.cuh), so every rule becomes bit-parallel set algebra over the whole board:
//   neighbours     shifts inside the row + one shuffle up / down
//   group of a stone  flood fill = Kogge-Stone occluded fill along the rows, one shuffle step between
//                     rows, iterated to the fixpoint (a handful of rounds for real groups)
//   liberties      dilate(group) & empty
//   play_move      go_env.py:172 -> place, capture liberty-less opponent neighbour groups, ko point
//   all_legal_moves go_env.py:193 -> empty & has-empty-neighbour, plus a per-point check of the few
//                     "surrounded" points (MiniGo's own shortcut), minus the ko point
//   score          Tromp-Taylor: empty regions reaching one colour only, two floods over the empties
//   observation    go_env.py:156-178: 16 history planes (mover, opponent) x 8 + player plane
#pragma once
#include "rz_common.cuh"

struct rz_goboard {
  uint32_t p[2];               // this lane's row: black (player 0) / white (player 1)
  uint32_t hist[RZ_GO_HIST];   // board_history planes 2..15 (planes 0,1 are p[player^1], p[player])
  int player;                  // to move: 0 black, 1 white
  int last_move;               // last ACTION (cells = pass), -1 before the first move
  int stones;                  // moves played, passes included (Position.n)
  int ko;                      // square that may not be retaken, -1 none
  int passes;                  // consecutive passes
};

// stones of the player to move / of the other player (selects, so the board stays in registers)
__device__ __forceinline__ uint32_t rz_go_mine(const rz_goboard& b) { return b.player ? b.p[1] : b.p[0]; }
__device__ __forceinline__ uint32_t rz_go_theirs(const rz_goboard& b) { return b.player ? b.p[0] : b.p[1]; }

__device__ __forceinline__ uint32_t rz_go_rowmask(const rz_geom& q) {
  return rz_lane() < q.H ? ((1u << q.W) - 1u) : 0u;   // W <= 19
}
// bits of the row above (r+1) / below (r-1) brought to this lane
__device__ __forceinline__ uint32_t rz_go_from_above(uint32_t x) {
  const uint32_t v = __shfl_down_sync(RZ_FULL, x, 1);
  return rz_lane() == 31 ? 0u : v;
}
__device__ __forceinline__ uint32_t rz_go_from_below(uint32_t x) {
  const uint32_t v = __shfl_up_sync(RZ_FULL, x, 1);
  return rz_lane() == 0 ? 0u : v;
}
// the 4-neighbourhood of a set (the set itself excluded unless adjacent members cover it)
__device__ __forceinline__ uint32_t rz_go_neighbours(uint32_t x, uint32_t mask) {
  return ((x << 1) | (x >> 1) | rz_go_from_above(x) | rz_go_from_below(x)) & mask;
}

// connected component(s) of `seed` inside `region`
__device__ __forceinline__ uint32_t rz_go_flood(uint32_t seed, uint32_t region) {
  uint32_t g = seed & region;
  for (;;) {
    uint32_t l = g, r = g, m = region;
    l |= m & (l << 1); m &= m << 1;
    l |= m & (l << 2); m &= m << 2;
    l |= m & (l << 4); m &= m << 4;
    l |= m & (l << 8); m &= m << 8;
    l |= m & (l << 16);
    m = region;
    r |= m & (r >> 1); m &= m >> 1;
    r |= m & (r >> 2); m &= m >> 2;
    r |= m & (r >> 4); m &= m >> 4;
    r |= m & (r >> 8); m &= m >> 8;
    r |= m & (r >> 16);
    const uint32_t h = l | r;
    const uint32_t n = (h | rz_go_from_above(h) | rz_go_from_below(h)) & region;
    if (!__any_sync(RZ_FULL, n != g)) return g;
    g = n;
  }
}

// lowest (row, column) member of a non-empty distributed set, as a one-bit set
__device__ __forceinline__ uint32_t rz_go_pick(uint32_t set, int& row, int& col) {
  const unsigned who = __ballot_sync(RZ_FULL, set != 0u);
  row = __ffs(who) - 1;
  const uint32_t w = __shfl_sync(RZ_FULL, set, row);
  col = __ffs(w) - 1;
  return rz_lane() == row ? (1u << col) : 0u;
}

__device__ __forceinline__ int rz_go_count(uint32_t set) { return rz_warp_sum_i32(__popc(set)); }

__device__ __forceinline__ void rz_go_load(rz_goboard& b, const uint32_t* __restrict__ rows,
                                           const uint32_t* __restrict__ hist,
                                           const int32_t* __restrict__ meta, int H) {
  const int lane = rz_lane();
  b.p[0] = lane < H ? rows[lane] : 0u;
  b.p[1] = lane < H ? rows[H + lane] : 0u;
#pragma unroll
  for (int i = 0; i < RZ_GO_HIST; ++i) b.hist[i] = (hist && lane < H) ? hist[i * H + lane] : 0u;
  b.player = meta[RZ_META_PLAYER];
  b.last_move = meta[RZ_META_LAST_MOVE];
  b.stones = meta[RZ_META_STONES];
  b.ko = meta[RZ_META_KO];
  b.passes = meta[RZ_META_PASSES];
}

__device__ __forceinline__ void rz_go_store(const rz_goboard& b, uint32_t* __restrict__ rows,
                                            uint32_t* __restrict__ hist, int H) {
  const int lane = rz_lane();
  if (lane < H) {
    rows[lane] = b.p[0];
    rows[H + lane] = b.p[1];
    if (hist) {
#pragma unroll
      for (int i = 0; i < RZ_GO_HIST; ++i) hist[i * H + lane] = b.hist[i];
    }
  }
}

// position words of the meta record (lane 0 calls)
__device__ __forceinline__ void rz_go_store_meta(const rz_goboard& b, int32_t* __restrict__ m) {
  m[RZ_META_PLAYER] = b.player;
  m[RZ_META_LAST_MOVE] = b.last_move;
  m[RZ_META_STONES] = b.stones;
  m[RZ_META_KO] = b.ko;
  m[RZ_META_PASSES] = b.passes;
}

// Position.play_move + the history shift of GoEnv.step (go_env.py:172-178).  a is warp-uniform and
// LEGAL (0 <= a <= cells; cells = pass).
__device__ __forceinline__ void rz_go_play(rz_goboard& b, int a, const rz_geom& q) {
#pragma unroll
  for (int i = RZ_GO_HIST - 1; i >= 2; --i) b.hist[i] = b.hist[i - 2];
  b.hist[0] = rz_go_theirs(b);
  b.hist[1] = rz_go_mine(b);
  if (a >= q.cells) {                      // pass_move: ko cleared
    b.ko = -1;
    b.passes += 1;
  } else {
    const uint32_t mask = rz_go_rowmask(q);
    const int r = a / q.W, c = a - r * q.W;
    const uint32_t bit = rz_lane() == r ? (1u << c) : 0u;
    uint32_t me = rz_go_mine(b), opp = rz_go_theirs(b);
    const uint32_t nbrs = rz_go_neighbours(bit, mask) & ~bit;
    // is_koish: every neighbour of the empty point is an opponent stone
    const bool koish = !__any_sync(RZ_FULL, (nbrs & ~opp) != 0u);
    me |= bit;
    const uint32_t empty = mask & ~(me | opp);
    uint32_t captured = 0u;
    uint32_t todo = nbrs & opp;
    while (__any_sync(RZ_FULL, todo != 0u)) {
      int sr, sc;
      const uint32_t seed = rz_go_pick(todo, sr, sc);
      const uint32_t grp = rz_go_flood(seed, opp);
      const uint32_t libs = rz_go_neighbours(grp, mask) & empty;
      if (!__any_sync(RZ_FULL, libs != 0u)) captured |= grp;
      todo &= ~grp;
    }
    opp &= ~captured;
    b.ko = -1;
    if (koish && rz_go_count(captured) == 1) {
      int kr, kc;
      rz_go_pick(captured, kr, kc);
      b.ko = kr * q.W + kc;
    }
    b.p[0] = b.player ? opp : me;
    b.p[1] = b.player ? me : opp;
    b.passes = 0;
  }
  b.player ^= 1;
  b.last_move = a;
  b.stones += 1;
}

// Position.all_legal_moves without the pass entry: this lane's row of legal points.
__device__ __forceinline__ uint32_t rz_go_legal_rows(const rz_goboard& b, const rz_geom& q) {
  const uint32_t mask = rz_go_rowmask(q);
  const uint32_t me = rz_go_mine(b), opp = rz_go_theirs(b);
  const uint32_t empty = mask & ~(me | opp);
  const uint32_t near_empty = rz_go_neighbours(empty, mask);
  uint32_t legal = empty & near_empty;            // a liberty of its own: never suicide
  uint32_t cand = empty & ~near_empty;            // MiniGo's "surrounded spots": checked one by one
  while (__any_sync(RZ_FULL, cand != 0u)) {
    int xr, xc;
    const uint32_t x = rz_go_pick(cand, xr, xc);
    const uint32_t nbrs = rz_go_neighbours(x, mask) & ~x;
    bool ok = false;
    // friendly neighbour groups keep a liberty besides x  <=>  their joint liberties count >= 2
    const uint32_t fn = nbrs & me;
    if (__any_sync(RZ_FULL, fn != 0u)) {
      const uint32_t grp = rz_go_flood(fn, me);
      ok = rz_go_count(rz_go_neighbours(grp, mask) & empty) >= 2;
    }
    // or the move captures: some opponent neighbour group has x as its only liberty
    uint32_t todo = ok ? 0u : (nbrs & opp);
    while (!ok && __any_sync(RZ_FULL, todo != 0u)) {
      int sr, sc;
      const uint32_t seed = rz_go_pick(todo, sr, sc);
      const uint32_t grp = rz_go_flood(seed, opp);
      ok = rz_go_count(rz_go_neighbours(grp, mask) & empty) == 1;
      todo &= ~grp;
    }
    if (ok) legal |= x;
    cand &= ~x;
  }
  if (b.ko >= 0 && rz_lane() == b.ko / q.W) legal &= ~(1u << (b.ko - (b.ko / q.W) * q.W));
  return legal;
}

// is action a (warp-uniform) legal here?  Position.is_move_legal; the pass always is
__device__ __forceinline__ bool rz_go_action_legal(const rz_goboard& b, int a, const rz_geom& q) {
  if (a < 0 || a > q.cells) return false;
  if (a == q.cells) return true;
  const uint32_t legal = rz_go_legal_rows(b, q);
  const int r = a / q.W;
  return (__shfl_sync(RZ_FULL, legal, r) >> (a - r * q.W)) & 1u;
}

// legality of action slot s (per-lane s) given the distributed legal rows
__device__ __forceinline__ bool rz_go_slot_legal(uint32_t legal_rows, int s, const rz_geom& q) {
  const int sc = s < q.cells ? s : 0;
  const int r = sc / q.W;
  const uint32_t row = __shfl_sync(RZ_FULL, legal_rows, r);
  return s == q.cells || (s < q.cells && ((row >> (sc - r * q.W)) & 1u));
}

// Position.score(): black area - white area - komi (Tromp-Taylor)
__device__ __forceinline__ double rz_go_score(const rz_goboard& b, const rz_geom& q) {
  const uint32_t mask = rz_go_rowmask(q);
  const uint32_t black = b.p[0], white = b.p[1];
  const uint32_t empty = mask & ~(black | white);
  const uint32_t rb = rz_go_flood(rz_go_neighbours(black, mask) & empty, empty);
  const uint32_t rw = rz_go_flood(rz_go_neighbours(white, mask) & empty, empty);
  const int diff = rz_go_count(black | (rb & ~rw)) - rz_go_count(white | (rw & ~rb));
  return (double)diff - q.komi;
}

// is_game_over (two consecutive passes; or the engine's move cap) + the winner the reference's
// rewards encode: black iff result() == 1, else white (go_env.py:142-143, a zero score goes to white)
__device__ __forceinline__ int rz_go_status(const rz_goboard& b, const rz_geom& q, int& winner) {
  winner = -1;
  if (b.passes >= 2 || (q.max_moves > 0 && b.stones >= q.max_moves)) {
    winner = rz_go_score(b, q) > 0.0 ? 0 : 1;
    return RZ_ENDED_WIN;
  }
  return RZ_ACTIVE;
}

// ---- the interface the search kernels (rz_tree.cu) and evaluators are written against -------------
struct rz_go_game {
  typedef rz_goboard board;
  static __device__ __forceinline__ void load_root(board& b, const rz_tree_desc& t, int g) {
    const int H = t.game.board_size;
    rz_go_load(b, t.root_rows + (size_t)g * 2 * H, t.root_hist + (size_t)g * RZ_GO_HIST * H,
               t.root_meta + (size_t)g * RZ_META_STRIDE, H);
  }
  static __device__ __forceinline__ void load_leaf(board& b, const rz_tree_desc& t, int g) {
    const int H = t.game.board_size;
    rz_go_load(b, t.leaf_rows + (size_t)g * 2 * H, t.leaf_hist + (size_t)g * RZ_GO_HIST * H,
               t.leaf_meta + (size_t)g * RZ_META_STRIDE, H);
  }
  static __device__ __forceinline__ void store_root(const board& b, const rz_tree_desc& t, int g) {
    const int H = t.game.board_size;
    rz_go_store(b, t.root_rows + (size_t)g * 2 * H, t.root_hist + (size_t)g * RZ_GO_HIST * H, H);
  }
  static __device__ __forceinline__ void store_leaf(const board& b, const rz_tree_desc& t, int g) {
    const int H = t.game.board_size;
    rz_go_store(b, t.leaf_rows + (size_t)g * 2 * H, t.leaf_hist + (size_t)g * RZ_GO_HIST * H, H);
  }
  static __device__ __forceinline__ void store_meta(const board& b, int32_t* m) { rz_go_store_meta(b, m); }
  static __device__ __forceinline__ void play(board& b, int a, const rz_geom& q) { rz_go_play(b, a, q); }
  static __device__ __forceinline__ int status(const board& b, const rz_geom& q, int& winner) {
    return rz_go_status(b, q, winner);
  }
  static __device__ __forceinline__ bool action_illegal(const board& b, int a, const rz_geom& q) {
    return !rz_go_action_legal(b, a, q);
  }
  static __device__ __forceinline__ uint32_t legal_ctx(const rz_tree_desc& t, int g, const rz_geom& q) {
    board b;
    const int H = q.H;
    rz_go_load(b, t.leaf_rows + (size_t)g * 2 * H, nullptr, t.leaf_meta + (size_t)g * RZ_META_STRIDE, H);
    return rz_go_legal_rows(b, q);
  }
  static __device__ __forceinline__ uint32_t legal_ctx(const board& b, const rz_geom& q) { return rz_go_legal_rows(b, q); }
  static __device__ __forceinline__ bool slot_legal(uint32_t ctx, int s, const rz_geom& q) {
    return rz_go_slot_legal(ctx, s, q);
  }
  static __device__ __forceinline__ void clear_root(const rz_tree_desc& t, int g) {
    const int lane = rz_lane(), H = t.game.board_size;
    if (lane < H) {
      t.root_rows[(size_t)g * 2 * H + lane] = 0u;
      t.root_rows[(size_t)g * 2 * H + H + lane] = 0u;
      for (int i = 0; i < RZ_GO_HIST; ++i) t.root_hist[((size_t)g * RZ_GO_HIST + i) * H + lane] = 0u;
    }
    if (lane == 0) {
      t.root_meta[(size_t)g * RZ_META_STRIDE + RZ_META_KO] = -1;
      t.root_meta[(size_t)g * RZ_META_STRIDE + RZ_META_PASSES] = 0;
    }
  }
  static __device__ __forceinline__ int stone_count(const board& b) { return rz_go_count(b.p[0] | b.p[1]); }
  static __device__ __forceinline__ uint32_t move_candidates(const board& b, const rz_geom& q) {
    return rz_go_legal_rows(b, q);
  }
  static __device__ __forceinline__ int candidate_action(int row, int col, const rz_geom& q) {
    return row * q.W + col;
  }
  static constexpr bool kHasPass = true;      // the pass is one more legal action (go_env.py:193-194)
};
