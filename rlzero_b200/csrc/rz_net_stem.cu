// rlzero_b200 -- fused observation encoder + first (stem) convolution of the trunk.
//
// Reference ops: GomokuEnv.current_state (rlzero/games/gomoku/gomoku_env.py:95-114) followed by
// the first nn.Conv2d(4, C, 3, padding=1) + ReLU of the trunk
// (rlzero/games/gomoku/policy_value_net.py:14,36; stem of the ResNet-N trunk).
//
// The 4 observation planes are bits, so the im2col row of a position (9 taps x 4 planes = 36
// values, padded to K = 64) is built straight from the bitboards in registers and written to
// shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads; the planes never
// exist in HBM.  One 128 x 128 x 64 MMA group per 128 positions, epilogue as in rz_net_tc2.cu
// (TMEM -> +bias -> ReLU -> pad mask -> bf16 -> 256-bit global stores from registers; the Go stem
// below still stages its tile for a TMA store).  The kernel is bound
// by its 64 KB/board output write; several CTAs per SM overlap build / MMA / drain phases.
#include <cuda_bf16.h>

#include "rz_common.cuh"
#include "rz_tc.cuh"

namespace {

constexpr int STEM_THREADS = 128;
constexpr int OFF_B = 0;                 // [128 cout][64 k] bf16, SW128
constexpr int OFF_A = 16384;             // [128 pos][64 k] bf16, SW128
constexpr int OFF_CTRL = 32768;          // (the output goes to HBM straight from registers: no staging tile)
constexpr int STEM_SMEM = OFF_CTRL + 2048 + 1024;   // control block + alignment slack

struct StemParams {
  const uint32_t* rows;   // [n][2][H]
  const int32_t* meta;    // [n][RZ_META_STRIDE]
  const float* planes;    // kPlanes: [n][4][H][H] float observation planes instead of bitboards
  const float* bias;      // [128]
  __nv_bfloat16* out;     // [rows_alloc][128]
  long long rows_alloc;   // n_tiles * 128
  int n_tiles;            // 128-row tiles of the padded position layout
  int n_boards;
  int H;                  // rows
  int W;                  // columns
  int relu;
};

// kS: row stride of the padded position layout (8: boards up to 7x7 such as Connect Four 6x7, 16: up to 15x15,
// 20: up to 19x19); a board owns kS*kS consecutive rows; a tile of 128 rows holds exactly two boards when kS = 8
// and may straddle two boards when kS = 20
template <bool kPlanes, int kS>
__global__ void __launch_bounds__(STEM_THREADS)
rz_stem_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const StemParams p) {
  rz::grid_dep_wait();     // programmatic dependent launch: see rz_common.cuh
  rz::grid_dep_launch();
  static_assert(kS * kS >= 64, "a 128-row tile must not touch more than two boards");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (rz::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - rz::smem_u32(smem_raw));
  const uint32_t bar_w = base + OFF_CTRL, bar_mma = base + OFF_CTRL + 8;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(al + OFF_CTRL + 16);
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(al + OFF_CTRL + 32);    // 2 boards x ([2][32] rows + 4 meta words)
  float* s_bias = reinterpret_cast<float*>(al + OFF_CTRL + 1024);        // [128]
  constexpr int P = kS * kS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W;

  if (tid == 0) {
    rz::tma_prefetch_desc(&tmap_w);
    rz::mbar_init(bar_w, 1);
    rz::mbar_init(bar_mma, 1);
    rz::fence_barrier_init();
    rz::mbar_expect_tx(bar_w, 16384);
    rz::tma_load_2d(base + OFF_B, &tmap_w, bar_w, 0, 0);
  }
  if (warp == 0) { rz::tmem_alloc(rz::smem_u32(tmem_holder), 128); rz::tmem_relinquish(); }
  s_bias[tid] = p.bias[tid];
  rz::tc_fence_before();
  __syncthreads();
  rz::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  constexpr uint32_t idesc = rz::umma_idesc_bf16(128, 128);
  constexpr uint32_t ONE = 0x3F80u;  // bf16 1.0

  uint32_t phase = 0;
  // the board words of a tile are fetched ONE TILE AHEAD into registers (their L2 / HBM latency, ~1 us of a ~4 us tile,
  // used to sit between the tile's first barrier and its im2col phase)
  auto fetch_board = [&](int tile_, uint32_t& word, uint32_t& metaw) {
    word = 0u; metaw = 0u;
    if (kPlanes || tile_ >= p.n_tiles) return;
    const int sel = tid >> 6, t = tid & 63, bb = (tile_ * 128) / P + sel;
    const int c = t >> 5, y = t & 31;
    if (y < H && bb < p.n_boards) word = p.rows[((size_t)bb * 2 + c) * H + y];
    if (t < 4 && bb < p.n_boards) metaw = (uint32_t)p.meta[(size_t)bb * RZ_META_STRIDE + t];
  };
  uint32_t nxt_word, nxt_meta;
  fetch_board(blockIdx.x, nxt_word, nxt_meta);
  int it = 0;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    const int b_first = (tile * 128) / P;
    const int my_row = tile * 128 + tid;
    const int b = my_row / P, bsel = b - b_first;       // this thread's board (0 or 1 within the tile)
    // ---- the (at most two) boards of this tile -> shared memory (rows beyond H / boards beyond n are zero)
    uint32_t* s_rows_t = s_rows;     // (every read of the previous tile's words is behind the barrier that ends its iteration)
    if (!kPlanes) {
      const int sel = tid >> 6, t = tid & 63;
      s_rows_t[sel * 68 + t] = nxt_word;
      if (t < 4) s_rows_t[sel * 68 + 64 + t] = nxt_meta;
      fetch_board(tile + gridDim.x, nxt_word, nxt_meta);
    }
    __syncthreads();
    // ---- im2col row of position r: k = tap*4 + plane (gomoku_env.py:95-114 per tap)
    {
      const int pos = my_row - b * P;
      const int y = pos / kS, x = pos - y * kS;
      const uint32_t* brd = s_rows_t + bsel * 68;
      const int player = (int)brd[64] & 1, last = (int)brd[65], stones = (int)brd[66];
      const uint32_t colour = (stones & 1) ? 0u : ONE;
      const bool out_inside = (x < W) && (y < H) && b < p.n_boards;
      uint32_t w[20];  // 5 chunks x 4 words (2 bf16 each): taps 0..8 (+ one empty tap slot)
#pragma unroll
      for (int i = 0; i < 20; ++i) w[i] = 0u;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        const bool in = out_inside && yy >= 0 && yy < H && xx >= 0 && xx < W;
        uint32_t f0 = 0u, f1 = 0u, f2 = 0u, f3 = 0u;
        if (kPlanes) {
          if (in) {
            const float* src = p.planes + (size_t)b * 4 * H * W + yy * W + xx;
            f0 = __bfloat16_as_ushort(__float2bfloat16_rn(src[0]));
            f1 = __bfloat16_as_ushort(__float2bfloat16_rn(src[H * W]));
            f2 = __bfloat16_as_ushort(__float2bfloat16_rn(src[2 * H * W]));
            f3 = __bfloat16_as_ushort(__float2bfloat16_rn(src[3 * H * W]));
          }
        } else {
          const uint32_t mine = brd[player * 32 + (yy & 31)], theirs = brd[(player ^ 1) * 32 + (yy & 31)];
          f0 = in ? ((mine >> (xx & 31)) & 1u) * ONE : 0u;
          f1 = in ? ((theirs >> (xx & 31)) & 1u) * ONE : 0u;
          f2 = (in && stones > 0 && last == yy * W + xx) ? ONE : 0u;
          f3 = in ? colour : 0u;
        }
        w[tap * 2 + 0] = f0 | (f1 << 16);
        w[tap * 2 + 1] = f2 | (f3 << 16);
      }
      const uint32_t arow = base + OFF_A + (uint32_t)tid * 128u;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t dst = arow + (((uint32_t)c ^ ((uint32_t)tid & 7u)) << 4);
        if (c < 5) rz::st_shared_v4(dst, w[c * 4], w[c * 4 + 1], w[c * 4 + 2], w[c * 4 + 3]);
        else       rz::st_shared_v4(dst, 0u, 0u, 0u, 0u);
      }
    }
    rz::fence_proxy_async();
    rz::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      rz::tc_fence_after();
      if (phase == 0) rz::mbar_wait(bar_w, 0);
      const uint64_t adesc = rz::umma_desc_sw128(base + OFF_A), bdesc = rz::umma_desc_sw128(base + OFF_B);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        rz::umma_bf16(tmem_base, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk > 0 ? 1u : 0u);
      rz::umma_commit(bar_mma);
    }
    rz::mbar_wait(bar_mma, phase & 1u);
    rz::tc_fence_after();
    // ---- epilogue: row per thread, 256-bit global stores straight from registers (L1::no_allocate, as in the
    // trunk convolution): no staging tile, no TMA store to wait for before the next tile
    if (p.relu & 2) {
      // float32-accurate 32-channel stem (the reference's own conv1 4 -> 32, policy_value_net.py:14): the weight
      // rows hold the bf16 high parts of the 32 filters (rows 0..31) and the rounding residues (rows 32..63), the
      // inputs are 0/1, so acc[c] + acc[32 + c] is the float32 convolution.  The activation leaves as a bf16
      // (high, low) pair per channel in the layout the next layer's single K = 128 pass consumes:
      // [hi 0..31 | lo 0..31 | hi 0..31 | lo 0..31] against weights [Whi | Whi | Wlo | Wlo].
      const int pos = my_row - b * P;
      const bool valid = (pos % kS < W) && (pos / kS < H) && b < p.n_boards;
      __nv_bfloat16* orow = p.out + (size_t)my_row * 128;
      uint32_t a_hi[32], a_lo[32];
      rz::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), a_hi);
      rz::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u, a_lo);
      rz::tmem_ld_wait();
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float v0 = (__uint_as_float(a_hi[2 * e]) + __uint_as_float(a_lo[2 * e])) + s_bias[2 * e];
        float v1 = (__uint_as_float(a_hi[2 * e + 1]) + __uint_as_float(a_lo[2 * e + 1])) + s_bias[2 * e + 1];
        if (p.relu & 1) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
        if (!valid) { v0 = 0.0f; v1 = 0.0f; }
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
        hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __low2float(h2), v1 - __high2float(h2));
        lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      if (my_row < p.rows_alloc) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          uint32_t v8[8];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = hi[j * 8 + e];
            rz::st_global_v8(orow + rep * 64 + j * 16, v8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = lo[j * 8 + e];
            rz::st_global_v8(orow + rep * 64 + 32 + j * 16, v8);
          }
        }
      }
    } else {
      const int pos = my_row - b * P;
      const bool valid = (pos % kS < W) && (pos / kS < H) && b < p.n_boards;
      __nv_bfloat16* orow = p.out + (size_t)my_row * 128;
      // two 32-column chunks in flight per wait (four serial load / wait round trips cost ~0.4 us of a ~4 us tile)
#pragma unroll
      for (int cp = 0; cp < 2; ++cp) {
        uint32_t acc2[2][32];
        rz::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cp * 64), acc2[0]);
        rz::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cp * 64 + 32), acc2[1]);
        rz::tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = cp * 2 + h;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t packed[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = j * 16 + e * 2;
              float v0 = __uint_as_float(acc2[h][c]) + s_bias[ch * 32 + c];
              float v1 = __uint_as_float(acc2[h][c + 1]) + s_bias[ch * 32 + c + 1];
              if (p.relu & 1) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
              if (!valid) { v0 = 0.0f; v1 = 0.0f; }
              const __nv_bfloat162 o2 = __floats2bfloat162_rn(v0, v1);
              packed[e] = *reinterpret_cast<const uint32_t*>(&o2);
            }
            if (my_row < p.rows_alloc) rz::st_global_v8(orow + ch * 32 + j * 16, packed);
          }
        }
      }
    }
    rz::tc_fence_before();
    __syncthreads();          // every warp has read its accumulator rows: the next tile's MMAs may overwrite TMEM
    ++phase;
  }
  rz::tc_fence_before();
  __syncthreads();
  if (warp == 0) { rz::tc_fence_after(); rz::tmem_dealloc(tmem_base, 128); }
}


// ---------------------------------------------------------------------------------------------
// Go: GoEnv.observe (rlzero/games/go/go_env.py:156-178; 16 history planes + player plane) fused with
// the stem conv3x3(17 -> 128).  Same structure as above with a wider im2col row: k = tap*17 + plane,
// 153 values padded to K = 192 = three 64-wide SW128 atoms (10 of the 12 k-steps are issued).
// ---------------------------------------------------------------------------------------------
constexpr int GO_PLANES = 17;
constexpr int GO_OFF_B = 0;               // 3 x [128 cout][64 k] bf16, SW128
constexpr int GO_OFF_A = 49152;           // 3 x [128 pos][64 k]
constexpr int GO_OFF_STAGE = GO_OFF_A;    // 2 x [128 pos][64 cout]: the A tile is dead once its MMAs are done, so
                                          // the output is staged over it (105 KB per CTA -> two CTAs per SM)
constexpr int GO_OFF_CTRL = 98304;
constexpr int GO_BOARD_WORDS = 16 * 32 + 4;                 // 16 planes x 32 rows + (player, pad)
constexpr int GO_STEM_SMEM = GO_OFF_CTRL + 8192 + 1024;

struct GoStemParams {
  const uint32_t* rows;   // [n][2][H]
  const uint32_t* hist;   // [n][RZ_GO_HIST][H]
  const int32_t* meta;    // [n][RZ_META_STRIDE]
  const float* planes;    // kPlanes: [n][17][H][W] float observation planes
  const float* bias;
  int n_tiles, n_boards, H, W, relu;
};

template <bool kPlanes, int kS>
__global__ void __launch_bounds__(STEM_THREADS)
rz_stem_go_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                     const GoStemParams p) {
  rz::grid_dep_wait();     // programmatic dependent launch: see rz_common.cuh
  rz::grid_dep_launch();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (rz::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - rz::smem_u32(smem_raw));
  const uint32_t bar_w = base + GO_OFF_CTRL, bar_mma = base + GO_OFF_CTRL + 8;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(al + GO_OFF_CTRL + 16);
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(al + GO_OFF_CTRL + 64);   // 2 boards x GO_BOARD_WORDS
  float* s_bias = reinterpret_cast<float*>(al + GO_OFF_CTRL + 4608);       // [128]
  constexpr int P = kS * kS;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int H = p.H, W = p.W;

  if (tid == 0) {
    rz::tma_prefetch_desc(&tmap_w);
    rz::tma_prefetch_desc(&tmap_out);
    rz::mbar_init(bar_w, 1);
    rz::mbar_init(bar_mma, 1);
    rz::fence_barrier_init();
    rz::mbar_expect_tx(bar_w, 3 * 16384);
    for (int a = 0; a < 3; ++a) rz::tma_load_2d(base + GO_OFF_B + a * 16384, &tmap_w, bar_w, a * 64, 0);
  }
  if (warp == 0) { rz::tmem_alloc(rz::smem_u32(tmem_holder), 128); rz::tmem_relinquish(); }
  s_bias[tid] = p.bias[tid];
  rz::tc_fence_before();
  __syncthreads();
  rz::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  constexpr uint32_t idesc = rz::umma_idesc_bf16(128, 128);
  constexpr uint32_t ONE = 0x3F80u;  // bf16 1.0

  uint32_t phase = 0;
  bool store_pending = false;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int b_first = (tile * 128) / P;
    const int my_row = tile * 128 + tid;
    const int b = my_row / P, bsel = b - b_first;
    if (!kPlanes) {
      // the (at most two) boards of this tile: plane 0 = stones of the last mover, 1 = of the player
      // to move, 2.. = history (go_env.py:174-178); rows beyond H / boards beyond n are zero
      for (int i = tid; i < 2 * 16 * 32; i += STEM_THREADS) {
        const int sel = i >> 9, pl = (i >> 5) & 15, y = i & 31, bb = b_first + sel;
        uint32_t v = 0u;
        if (y < H && bb < p.n_boards) {
          const int player = p.meta[(size_t)bb * RZ_META_STRIDE + RZ_META_PLAYER] & 1;
          if (pl < 2) v = p.rows[((size_t)bb * 2 + ((pl == 0) ? (player ^ 1) : player)) * H + y];
          else v = p.hist[((size_t)bb * RZ_GO_HIST + (pl - 2)) * H + y];
        }
        s_rows[sel * GO_BOARD_WORDS + pl * 32 + y] = v;
      }
      if (tid < 2) {
        const int bb = b_first + tid;
        s_rows[tid * GO_BOARD_WORDS + 512] =
            bb < p.n_boards ? (uint32_t)(p.meta[(size_t)bb * RZ_META_STRIDE + RZ_META_PLAYER] & 1) : 0u;
      }
    }
    if (tid == 0 && store_pending) rz::tma_store_wait_read();
    __syncthreads();
    {
      const int pos = my_row - b * P;
      const int y = pos / kS, x = pos - y * kS;
      const uint32_t* brd = s_rows + bsel * GO_BOARD_WORDS;
      const bool out_inside = (x < W) && (y < H) && b < p.n_boards;
      uint32_t w[96];
#pragma unroll
      for (int i = 0; i < 96; ++i) w[i] = 0u;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        const bool in = out_inside && yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
        for (int pl = 0; pl < GO_PLANES; ++pl) {
          uint32_t f = 0u;
          if (kPlanes) {
            if (in) f = __bfloat16_as_ushort(__float2bfloat16_rn(
                        p.planes[((size_t)b * GO_PLANES + pl) * H * W + yy * W + xx]));
          } else if (pl < 16) {
            f = in ? ((brd[pl * 32 + (yy & 31)] >> (xx & 31)) & 1u) * ONE : 0u;
          } else {
            f = (in && brd[512]) ? ONE : 0u;
          }
          const int k = tap * GO_PLANES + pl;
          w[k >> 1] |= f << (16 * (k & 1));
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const uint32_t arow = base + GO_OFF_A + (uint32_t)a * 16384u + (uint32_t)tid * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t dst = arow + (((uint32_t)c ^ ((uint32_t)tid & 7u)) << 4);
          const int o = a * 32 + c * 4;
          rz::st_shared_v4(dst, w[o], w[o + 1], w[o + 2], w[o + 3]);
        }
      }
    }
    rz::fence_proxy_async();
    rz::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      rz::tc_fence_after();
      if (phase == 0) rz::mbar_wait(bar_w, 0);
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        const uint64_t adesc = rz::umma_desc_sw128(base + GO_OFF_A + (kk >> 2) * 16384);
        const uint64_t bdesc = rz::umma_desc_sw128(base + GO_OFF_B + (kk >> 2) * 16384);
        rz::umma_bf16(tmem_base, adesc + (uint64_t)(2 * (kk & 3)), bdesc + (uint64_t)(2 * (kk & 3)), idesc,
                      kk > 0 ? 1u : 0u);
      }
      rz::umma_commit(bar_mma);
    }
    rz::mbar_wait(bar_mma, phase & 1u);
    rz::tc_fence_after();
    {
      const int pos = my_row - b * P;
      const bool valid = (pos % kS < W) && (pos / kS < H) && b < p.n_boards;
      const uint32_t stage_row = base + GO_OFF_STAGE + (uint32_t)tid * 128u;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t acc[32];
        rz::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32), acc);
        rz::tmem_ld_wait();
        const uint32_t srow = stage_row + (uint32_t)(ch >> 1) * 16384u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t packed[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = j * 8 + e * 2;
            float v0 = __uint_as_float(acc[c]) + s_bias[ch * 32 + c];
            float v1 = __uint_as_float(acc[c + 1]) + s_bias[ch * 32 + c + 1];
            if (p.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
            if (!valid) { v0 = 0.0f; v1 = 0.0f; }
            const __nv_bfloat162 o2 = __floats2bfloat162_rn(v0, v1);
            packed[e] = *reinterpret_cast<const uint32_t*>(&o2);
          }
          const uint32_t chunk = (uint32_t)((ch & 1) * 4 + j);
          rz::st_shared_v4(srow + ((chunk ^ ((uint32_t)tid & 7u)) << 4), packed[0], packed[1], packed[2], packed[3]);
        }
      }
    }
    rz::fence_proxy_async();
    rz::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      rz::tma_store_2d(&tmap_out, base + GO_OFF_STAGE, 0, tile * 128);
      rz::tma_store_2d(&tmap_out, base + GO_OFF_STAGE + 16384, 64, tile * 128);
      rz::tma_store_commit();
    }
    store_pending = true;
    ++phase;
  }
  if (tid == 0 && store_pending) rz::tma_store_wait_all();
  rz::tc_fence_before();
  __syncthreads();
  if (warp == 0) { rz::tc_fence_after(); rz::tmem_dealloc(tmem_base, 128); }
}

}  // namespace

static int stem_launch(const rz_game_desc* g, const uint32_t* rows, const int32_t* meta, const float* planes,
                       const void* weight, const float* bias, void* act_out, int n_boards, int relu,
                       int n_ctas, void* stream) {
  if (n_boards == 0) return 0;
  const int H = g->board_size, W = g->width > 0 ? g->width : g->board_size;
  const int S = rz_row_stride(H, W, g->row_stride);
  RZ_REQUIRE(S != 0, "rz_net_stem_tc: row_stride %d does not hold a %dx%d board", g->row_stride, H, W);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rz_stem_tc_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_tc_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_tc_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_tc_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_tc_kernel<false, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_tc_kernel<true, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e != cudaSuccess) { rz_set_error("rz_net_stem_tc: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    attr_set = true;
  }
  // the output tensor is padded to a multiple of 256 rows (the convolutions work on pairs of 128-row tiles)
  const long long rows_alloc = ((long long)n_boards * S * S + 255) / 256 * 256;
  CUtensorMap tmap_w;
  if (rz::make_tmap_2d(&tmap_w, weight, 128, 64, 128)) return -1;
  StemParams p;
  p.rows = rows; p.meta = meta; p.planes = planes; p.bias = bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(act_out); p.rows_alloc = rows_alloc;
  p.n_tiles = (int)(rows_alloc / 128); p.n_boards = n_boards; p.H = H; p.W = W; p.relu = relu;
  int ctas = n_ctas > 0 ? n_ctas : 148 * 4;   // 35 KB of shared memory and 128 TMEM columns per CTA: four per SM
  if (ctas > p.n_tiles) ctas = p.n_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 8) {
    if (planes) rz_launch_pdl(rz_stem_tc_kernel<true, 8>, ctas, STEM_THREADS, STEM_SMEM, st, tmap_w, p);
    else        rz_launch_pdl(rz_stem_tc_kernel<false, 8>, ctas, STEM_THREADS, STEM_SMEM, st, tmap_w, p);
  } else if (S == 16) {
    if (planes) rz_launch_pdl(rz_stem_tc_kernel<true, 16>, ctas, STEM_THREADS, STEM_SMEM, st, tmap_w, p);
    else        rz_launch_pdl(rz_stem_tc_kernel<false, 16>, ctas, STEM_THREADS, STEM_SMEM, st, tmap_w, p);
  } else {
    if (planes) rz_launch_pdl(rz_stem_tc_kernel<true, 20>, ctas, STEM_THREADS, STEM_SMEM, st, tmap_w, p);
    else        rz_launch_pdl(rz_stem_tc_kernel<false, 20>, ctas, STEM_THREADS, STEM_SMEM, st, tmap_w, p);
  }
  RZ_LAUNCH_CHECK("rz_net_stem_tc");
  return 0;
}

extern "C" int rz_net_stem_tc(const rz_game_desc* g, const uint32_t* rows, const int32_t* meta,
                              const void* weight, const float* bias, void* act_out, int n_boards,
                              int relu, int n_ctas, void* stream) {
  if (rz_check_game(g)) return -1;
  RZ_REQUIRE(rows && meta && weight && bias && act_out, "rz_net_stem_tc: null argument");
  RZ_REQUIRE(g->board_size <= 19 && g->width <= 19, "rz_net_stem_tc: the padded tile layouts hold boards up to 19x19");
  RZ_REQUIRE(n_boards >= 0, "rz_net_stem_tc: n_boards %d", n_boards);
  return stem_launch(g, rows, meta, nullptr, weight, bias, act_out, n_boards, relu, n_ctas, stream);
}

extern "C" int rz_net_stem_tc_planes(const rz_game_desc* g, const float* planes, const void* weight,
                                     const float* bias, void* act_out, int n_boards, int relu, int n_ctas,
                                     void* stream) {
  if (rz_check_game(g)) return -1;
  RZ_REQUIRE(planes && weight && bias && act_out, "rz_net_stem_tc_planes: null argument");
  RZ_REQUIRE(g->board_size <= 19 && g->width <= 19, "rz_net_stem_tc_planes: the padded tile layouts hold boards up to 19x19");
  RZ_REQUIRE(n_boards >= 0, "rz_net_stem_tc_planes: n_boards %d", n_boards);
  return stem_launch(g, nullptr, nullptr, planes, weight, bias, act_out, n_boards, relu, n_ctas, stream);
}

static int go_stem_launch(const rz_game_desc* g, const uint32_t* rows, const uint32_t* hist, const int32_t* meta,
                          const float* planes, const void* weight, const float* bias, void* act_out, int n_boards,
                          int relu, int n_ctas, void* stream) {
  if (n_boards == 0) return 0;
  const int H = g->board_size, W = H;
  const int S = H <= 15 ? 16 : 20;   // the Go stem keeps at most two boards per tile in its plane staging: 16 / 20 only
  RZ_REQUIRE(g->row_stride == 0 || g->row_stride == S, "rz_net_stem_go_tc: row_stride %d (a %dx%d Go board uses %d)",
             g->row_stride, H, W, S);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rz_stem_go_tc_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GO_STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_go_tc_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GO_STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_go_tc_kernel<false, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, GO_STEM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rz_stem_go_tc_kernel<true, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, GO_STEM_SMEM);
    if (e != cudaSuccess) { rz_set_error("rz_net_stem_go_tc: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    attr_set = true;
  }
  const long long rows_alloc = ((long long)n_boards * S * S + 255) / 256 * 256;
  CUtensorMap tmap_w, tmap_out;
  if (rz::make_tmap_2d(&tmap_w, weight, 128, 192, 128)) return -1;
  if (rz::make_tmap_2d(&tmap_out, act_out, (uint64_t)rows_alloc, 128, 128)) return -1;
  GoStemParams p;
  p.rows = rows; p.hist = hist; p.meta = meta; p.planes = planes; p.bias = bias;
  p.n_tiles = (int)(rows_alloc / 128); p.n_boards = n_boards; p.H = H; p.W = W; p.relu = relu;
  int ctas = n_ctas > 0 ? n_ctas : 148 * 2;
  if (ctas > p.n_tiles) ctas = p.n_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 16) {
    if (planes) rz_launch_pdl(rz_stem_go_tc_kernel<true, 16>, ctas, STEM_THREADS, GO_STEM_SMEM, st, tmap_w, tmap_out, p);
    else        rz_launch_pdl(rz_stem_go_tc_kernel<false, 16>, ctas, STEM_THREADS, GO_STEM_SMEM, st, tmap_w, tmap_out, p);
  } else {
    if (planes) rz_launch_pdl(rz_stem_go_tc_kernel<true, 20>, ctas, STEM_THREADS, GO_STEM_SMEM, st, tmap_w, tmap_out, p);
    else        rz_launch_pdl(rz_stem_go_tc_kernel<false, 20>, ctas, STEM_THREADS, GO_STEM_SMEM, st, tmap_w, tmap_out, p);
  }
  RZ_LAUNCH_CHECK("rz_net_stem_go_tc");
  return 0;
}

extern "C" int rz_net_stem_go_tc(const rz_game_desc* g, const uint32_t* rows, const uint32_t* hist,
                                 const int32_t* meta, const void* weight, const float* bias, void* act_out,
                                 int n_boards, int relu, int n_ctas, void* stream) {
  if (rz_check_game(g)) return -1;
  RZ_REQUIRE(g->game_type == RZ_GAME_GO, "rz_net_stem_go_tc: game_type %d is not RZ_GAME_GO", g->game_type);
  RZ_REQUIRE(rows && hist && meta && weight && bias && act_out, "rz_net_stem_go_tc: null argument");
  RZ_REQUIRE(n_boards >= 0, "rz_net_stem_go_tc: n_boards %d", n_boards);
  return go_stem_launch(g, rows, hist, meta, nullptr, weight, bias, act_out, n_boards, relu, n_ctas, stream);
}

extern "C" int rz_net_stem_go_tc_planes(const rz_game_desc* g, const float* planes, const void* weight,
                                        const float* bias, void* act_out, int n_boards, int relu, int n_ctas,
                                        void* stream) {
  if (rz_check_game(g)) return -1;
  RZ_REQUIRE(g->game_type == RZ_GAME_GO, "rz_net_stem_go_tc_planes: game_type %d is not RZ_GAME_GO", g->game_type);
  RZ_REQUIRE(planes && weight && bias && act_out, "rz_net_stem_go_tc_planes: null argument");
  RZ_REQUIRE(n_boards >= 0, "rz_net_stem_go_tc_planes: n_boards %d", n_boards);
  return go_stem_launch(g, nullptr, nullptr, nullptr, planes, weight, bias, act_out, n_boards, relu, n_ctas, stream);
}
