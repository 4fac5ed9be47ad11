// rlzero_b200 -- 3x3 convolution of the policy-value trunk as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Reference op: nn.Conv2d(C, 128, kernel_size=3, padding=1) (+ folded BatchNorm, residual add,
// ReLU) as used by the trunk of rlzero/games/gomoku/policy_value_net.py:14-16,36-38 and the
// ResNet-N trunk SURVEY.md section 7 defines for the benchmark configs.
//
// Data layout (bf16, channels last, one 16x16 tile of positions per board):
//   act[b][p][c], p = y*16 + x; squares with x >= H or y >= H hold ZERO in every activation
//   tensor.  With row stride 16 > H the zero column x = 15 is at once the right halo of row y
//   and the left halo of row y+1, and the zero row y = 15 is the bottom halo of board b and the
//   top halo of board b+1, so the input of tap (dy,dx) for 128 consecutive output positions is
//   simply the 128 consecutive rows starting dy*16+dx further on: ONE 2-D TMA box per tap and
//   k-block, no im2col buffer, no boundary code (rows before the tensor start are zero-filled
//   by TMA).
//   w[tap][cout][cin] (K-major B operand), tap = kh*3 + kw.
//
// Kernel: persistent, one CTA per SM, 128(M) x 128(N) output tile = half a board.
//   warp 0      TMA producer  (A tap tile 128x64 + B tap tile 128x64 per stage, 128B swizzle)
//   warp 1      MMA issuer    (one elected lane: 4 x tcgen05.mma M128 N128 K16 per stage)
//   warps 2..5  epilogue      (tcgen05.ld -> +bias (+residual) -> ReLU -> zero the pad squares
//                              -> bf16 -> 16-byte global stores), double-buffered TMEM so the
//                              epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_bf16.h>

#include "rz_common.cuh"
#include "rz_tc.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int TILE_N = 128;
constexpr int KBLK = 64;  // bf16 channels per k-block = one 128-byte swizzle row
constexpr int STAGES = 6;
constexpr int A_BYTES = TILE_M * KBLK * 2;
constexpr int B_BYTES = TILE_N * KBLK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers, bias*/;

struct ConvParams {
  const float* bias;               // [128] (BatchNorm folded in)
  const __nv_bfloat16* residual;   // [rows][128] or null
  __nv_bfloat16* out;              // [rows][128]
  int n_tiles;                     // rows / 128
  int board;                       // H
  int kblocks;                     // Cin / 64
  int relu;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
rz_conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_act,
                     const __grid_constant__ CUtensorMap tmap_w, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (rz::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - rz::smem_u32(smem_raw));
  const uint32_t ctrl = smem_base + STAGES * STAGE_BYTES;
  uint8_t* ctrl_ptr = smem_al + STAGES * STAGE_BYTES;
  // control block: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr, bias[128]
  const uint32_t bar_full = ctrl, bar_empty = ctrl + 8 * STAGES;
  const uint32_t bar_tfull = ctrl + 16 * STAGES, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(ctrl_ptr + 16 * STAGES + 32);
  float* s_bias = reinterpret_cast<float*>(ctrl_ptr + 16 * STAGES + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_iters = 9 * p.kblocks;

  if (warp == 0 && lane == 0) {
    rz::tma_prefetch_desc(&tmap_act);
    rz::tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      rz::mbar_init(bar_full + 8 * s, 1);
      rz::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      rz::mbar_init(bar_tfull + 8 * b, 1);
      rz::mbar_init(bar_tempty + 8 * b, 4);
    }
    rz::fence_barrier_init();
  }
  if (warp == 1) {
    rz::tmem_alloc(rz::smem_u32(tmem_holder), TMEM_COLS);
    rz::tmem_relinquish();
  }
  if (threadIdx.x >= 64) s_bias[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  rz::tc_fence_before();
  __syncthreads();
  rz::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int row0 = tile * TILE_M;
        for (int tap = 0; tap < 9; ++tap) {
          const int off = (tap / 3 - 1) * 16 + (tap % 3 - 1);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            rz::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            const uint32_t a_dst = smem_base + s * STAGE_BYTES;
            rz::mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
            rz::tma_load_2d(a_dst, &tmap_act, bar_full + 8 * s, kb * KBLK, row0 + off);
            rz::tma_load_2d(a_dst + A_BYTES, &tmap_w, bar_full + 8 * s, kb * KBLK, tap * TILE_N);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = rz::umma_idesc_bf16(TILE_M, TILE_N);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        rz::mbar_wait(bar_tempty + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u);
        rz::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * TILE_N;
        for (int k = 0; k < k_iters; ++k) {
          rz::mbar_wait(bar_full + 8 * s, ph);
          rz::tc_fence_after();
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          const uint64_t adesc = rz::umma_desc_sw128(a_addr);
          const uint64_t bdesc = rz::umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int kk = 0; kk < KBLK / 16; ++kk) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the >>4 field
            rz::umma_bf16(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc,
                          (k > 0 || kk > 0) ? 1u : 0u);
          }
          rz::umma_commit(bar_empty + 8 * s);  // frees the stage when these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        rz::umma_commit(bar_tfull + 8 * buf);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      rz::mbar_wait(bar_tfull + 8 * buf, (uint32_t)(it >> 1) & 1u);
      rz::tc_fence_after();
      const size_t row = (size_t)tile * TILE_M + q * 32 + lane;
      const int pos = (int)(row & 255);
      const bool valid = ((pos & 15) < p.board) && ((pos >> 4) < p.board);
      __nv_bfloat16* orow = p.out + row * TILE_N;
      const __nv_bfloat16* rrow = p.residual ? p.residual + row * TILE_N : nullptr;
#pragma unroll 1
      for (int ch = 0; ch < TILE_N / 32; ++ch) {
        uint32_t acc[32];
        rz::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TILE_N + ch * 32), acc);
        rz::tmem_ld_wait();
        uint4 res[4];
        if (rrow && valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) res[j] = *reinterpret_cast<const uint4*>(rrow + ch * 32 + j * 8);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t packed[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = j * 8 + e * 2;
            float v0 = __uint_as_float(acc[c]) + s_bias[ch * 32 + c];
            float v1 = __uint_as_float(acc[c + 1]) + s_bias[ch * 32 + c + 1];
            if (rrow && valid) {
              const uint32_t rw = (&res[j].x)[e];
              const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&rw);
              v0 += __low2float(r2);
              v1 += __high2float(r2);
            }
            if (p.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
            if (!valid) { v0 = 0.0f; v1 = 0.0f; }
            const __nv_bfloat162 o2 = __floats2bfloat162_rn(v0, v1);
            packed[e] = *reinterpret_cast<const uint32_t*>(&o2);
          }
          *reinterpret_cast<uint4*>(orow + ch * 32 + j * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      }
      rz::tc_fence_before();
      __syncwarp();
      if (lane == 0) rz::mbar_arrive(bar_tempty + 8 * buf);
    }
  }

  rz::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    rz::tc_fence_after();
    rz::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

extern "C" int rz_net_conv3x3_tc(const void* act_in, const void* weight, const float* bias,
                                 const void* residual, void* act_out, int n_boards, int board_size,
                                 int c_in, int relu, int n_ctas, void* stream) {
  RZ_REQUIRE(act_in && weight && bias && act_out, "rz_net_conv3x3_tc: null argument");
  RZ_REQUIRE(n_boards >= 0, "rz_net_conv3x3_tc: n_boards %d", n_boards);
  RZ_REQUIRE(board_size >= 1 && board_size <= 15, "rz_net_conv3x3_tc: board_size %d not in [1,15]", board_size);
  RZ_REQUIRE(c_in == 64 || c_in == 128, "rz_net_conv3x3_tc: c_in %d (64 or 128)", c_in);
  RZ_REQUIRE(act_in != act_out, "rz_net_conv3x3_tc: in-place convolution is not supported");
  if (n_boards == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rz_conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { rz_set_error("rz_net_conv3x3_tc: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    attr_set = true;
  }
  CUtensorMap tmap_act, tmap_w;
  if (rz::make_tmap_2d(&tmap_act, act_in, (uint64_t)n_boards * 256, (uint64_t)c_in, TILE_M)) return -1;
  if (rz::make_tmap_2d(&tmap_w, weight, (uint64_t)9 * 128, (uint64_t)c_in, TILE_N)) return -1;
  ConvParams p;
  p.bias = bias;
  p.residual = (const __nv_bfloat16*)residual;
  p.out = (__nv_bfloat16*)act_out;
  p.n_tiles = n_boards * 2;
  p.board = board_size;
  p.kblocks = c_in / KBLK;
  p.relu = relu;
  int ctas = n_ctas > 0 ? n_ctas : 148;
  if (ctas > p.n_tiles) ctas = p.n_tiles;
  rz_conv3x3_tc_kernel<<<ctas, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmap_act, tmap_w, p);
  RZ_LAUNCH_CHECK("rz_net_conv3x3_tc");
  return 0;
}
