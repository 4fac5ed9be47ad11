// rlzero_b200 -- Gomoku position held by one warp, one board row per lane.
//
// Lane r (< H) keeps row r of each colour as a bitmask (bit w = column w); lanes >= H
// keep zero.  Every rule of rlzero/games/gomoku/gomoku_env.py becomes a handful of
// shifts, ANDs and warp shuffles:
//   step            gomoku_env.py:49-70   -> one lane sets one bit
//   has_a_winner    gomoku_env.py:116-170 -> AND of k shifted copies along 4 directions
//   leagel_actions  gomoku_env.py:72-73   -> ~(p0|p1) & row mask
//   current_state   gomoku_env.py:95-114  -> plane bits read straight from the rows
#pragma once
#include "rz_common.cuh"

struct rz_wboard {
  uint32_t p[2];    // this lane's row of player 0 / player 1
  int player;       // to move
  int last_move;    // -1 on an empty board
  int stones;
};

__device__ __forceinline__ void rz_board_load(rz_wboard& b, const uint32_t* __restrict__ rows,
                                              const int32_t* __restrict__ meta, int H) {
  const int lane = rz_lane();
  b.p[0] = lane < H ? rows[lane] : 0u;
  b.p[1] = lane < H ? rows[H + lane] : 0u;
  b.player = meta[RZ_META_PLAYER];
  b.last_move = meta[RZ_META_LAST_MOVE];
  b.stones = meta[RZ_META_STONES];
}

__device__ __forceinline__ void rz_board_store_rows(const rz_wboard& b, uint32_t* __restrict__ rows,
                                                    int H) {
  const int lane = rz_lane();
  if (lane < H) {
    rows[lane] = b.p[0];
    rows[H + lane] = b.p[1];
  }
}

// square an action lands on (warp-uniform a): the square itself, or for a gravity game the lowest
// empty row of column a (stones of a column are contiguous from row 0, so that row is their count)
__device__ __forceinline__ int rz_board_action_cell(const rz_wboard& b, int a, const rz_geom& q) {
  if (!q.gravity) return a;
  const unsigned col = __ballot_sync(RZ_FULL, ((b.p[0] | b.p[1]) >> a) & 1u);
  return __popc(col) * q.W + a;
}

// is action a (warp-uniform, 0 <= a < A) illegal?  occupied square / full column
__device__ __forceinline__ bool rz_board_occupied(const rz_wboard& b, int a, const rz_geom& q) {
  const int r = q.gravity ? q.H - 1 : a / q.W, c = q.gravity ? a : a - (a / q.W) * q.W;
  const uint32_t occ = __shfl_sync(RZ_FULL, b.p[0] | b.p[1], r);
  return (occ >> c) & 1u;
}

// gomoku_env.py:55-57,67-68 -- place the mover's stone and flip the player (a is warp-uniform)
__device__ __forceinline__ void rz_board_play(rz_wboard& b, int a, const rz_geom& q) {
  const int cell = rz_board_action_cell(b, a, q);
  const int r = cell / q.W, c = cell - r * q.W;
  if (rz_lane() == r) b.p[b.player] |= (1u << c);
  b.player ^= 1;
  b.last_move = cell;
  b.stones += 1;
}

// k stones in a row for one colour (x = this lane's row of that colour); warp-uniform result.
__device__ __forceinline__ bool rz_rows_have_line(uint32_t x, int k) {
  uint32_t h = x, v = x, d = x, e = x;
  for (int i = 1; i < k; ++i) {
    const uint32_t below = __shfl_down_sync(RZ_FULL, x, i);  // row r+i (0 beyond the board)
    const uint32_t up = (rz_lane() + i < 32) ? below : 0u;
    h &= (x >> i);          // (r, c..c+k-1)          gomoku_env.py:139-142
    v &= up;                // (r..r+k-1, c)          gomoku_env.py:144-150
    d &= (up >> i);         // (r+i, c+i)             gomoku_env.py:152-159
    e &= (up << i);         // (r+i, c-i)             gomoku_env.py:161-168
  }
  return __any_sync(RZ_FULL, (h | v | d | e) != 0u);
}

// game_end_winner (gomoku_env.py:196-203) incl. the stones < 2k-1 early-out of
// has_a_winner (gomoku_env.py:131-133).  Returns rz_status; winner via reference.
__device__ __forceinline__ int rz_board_status(const rz_wboard& b, const rz_geom& q, int& winner) {
  winner = -1;
  if (b.stones >= 2 * q.k - 1) {
    const bool w0 = rz_rows_have_line(b.p[0], q.k);
    const bool w1 = rz_rows_have_line(b.p[1], q.k);
    if (w0 || w1) {
      winner = w0 ? 0 : 1;
      return RZ_ENDED_WIN;
    }
  }
  if (b.stones >= q.cells) return RZ_ENDED_TIE;
  return RZ_ACTIVE;
}

// legality of action slot s (per-lane s, all lanes must call): empty square / column not full
__device__ __forceinline__ bool rz_board_slot_legal(const rz_wboard& b, int s, const rz_geom& q) {
  const int sc = s < q.A ? s : 0;
  const int r = q.gravity ? q.H - 1 : sc / q.W, c = q.gravity ? sc : sc - (sc / q.W) * q.W;
  const uint32_t occ = __shfl_sync(RZ_FULL, b.p[0] | b.p[1], r);
  return s < q.A && !((occ >> c) & 1u);
}
// the same from the combined occupancy rows (lane r holds row r)
__device__ __forceinline__ bool rz_occ_slot_legal(uint32_t myocc, int s, const rz_geom& q) {
  const int sc = s < q.A ? s : 0;
  const int r = q.gravity ? q.H - 1 : sc / q.W, c = q.gravity ? sc : sc - (sc / q.W) * q.W;
  const uint32_t occ = __shfl_sync(RZ_FULL, myocc, r);
  return s < q.A && !((occ >> c) & 1u);
}

// ---- the interface the search kernels (rz_tree.cu) and evaluators are written against -------------
struct rz_line_game {
  typedef rz_wboard board;
  static __device__ __forceinline__ void load_root(board& b, const rz_tree_desc& t, int g) {
    const int H = t.game.board_size;
    rz_board_load(b, t.root_rows + (size_t)g * 2 * H, t.root_meta + (size_t)g * RZ_META_STRIDE, H);
  }
  static __device__ __forceinline__ void load_leaf(board& b, const rz_tree_desc& t, int g) {
    const int H = t.game.board_size;
    rz_board_load(b, t.leaf_rows + (size_t)g * 2 * H, t.leaf_meta + (size_t)g * RZ_META_STRIDE, H);
  }
  static __device__ __forceinline__ void store_root(const board& b, const rz_tree_desc& t, int g) {
    rz_board_store_rows(b, t.root_rows + (size_t)g * 2 * t.game.board_size, t.game.board_size);
  }
  static __device__ __forceinline__ void store_leaf(const board& b, const rz_tree_desc& t, int g) {
    rz_board_store_rows(b, t.leaf_rows + (size_t)g * 2 * t.game.board_size, t.game.board_size);
  }
  static __device__ __forceinline__ void store_meta(const board& b, int32_t* m) {   // lane 0
    m[RZ_META_PLAYER] = b.player;
    m[RZ_META_LAST_MOVE] = b.last_move;
    m[RZ_META_STONES] = b.stones;
  }
  static __device__ __forceinline__ void play(board& b, int a, const rz_geom& q) { rz_board_play(b, a, q); }
  static __device__ __forceinline__ int status(const board& b, const rz_geom& q, int& winner) {
    return rz_board_status(b, q, winner);
  }
  static __device__ __forceinline__ bool action_illegal(const board& b, int a, const rz_geom& q) {
    return a >= q.A || rz_board_occupied(b, a, q);                 // gomoku_env.py:51
  }
  // legality of the leaf's action slots (expand): context = this lane's occupancy row
  static __device__ __forceinline__ uint32_t legal_ctx(const rz_tree_desc& t, int g, const rz_geom& q) {
    const int lane = rz_lane(), H = q.H;
    return lane < H ? (t.leaf_rows[(size_t)g * 2 * H + lane] | t.leaf_rows[(size_t)g * 2 * H + H + lane]) : 0u;
  }
  static __device__ __forceinline__ uint32_t legal_ctx(const board& b, const rz_geom& q) { return b.p[0] | b.p[1]; }
  static __device__ __forceinline__ bool slot_legal(uint32_t ctx, int s, const rz_geom& q) {
    return rz_occ_slot_legal(ctx, s, q);
  }
  static __device__ __forceinline__ void clear_root(const rz_tree_desc& t, int g) {  // fresh episode position
    const int lane = rz_lane(), H = t.game.board_size;
    if (lane < H) { t.root_rows[(size_t)g * 2 * H + lane] = 0u; t.root_rows[(size_t)g * 2 * H + H + lane] = 0u; }
  }
  static __device__ __forceinline__ int stone_count(const board& b) { return b.stones; }
  // random playouts: this lane's row of squares a random mover may pick, and the action of (row, col)
  static __device__ __forceinline__ uint32_t move_candidates(const board& b, const rz_geom& q) {
    const uint32_t rowmask = (q.W >= 32) ? 0xffffffffu : ((1u << q.W) - 1u);
    const int lane = rz_lane();
    return (q.gravity ? lane == q.H - 1 : lane < q.H) ? (~(b.p[0] | b.p[1]) & rowmask) : 0u;
  }
  static __device__ __forceinline__ int candidate_action(int row, int col, const rz_geom& q) {
    return q.gravity ? col : row * q.W + col;
  }
  static constexpr bool kHasPass = false;
};
