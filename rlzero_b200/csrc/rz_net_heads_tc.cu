// rlzero_b200 -- the fully connected layers of the policy / value heads on the tensor cores.
//
// Reference ops (rlzero/games/gomoku/policy_value_net.py): act_fc1 (4HW -> A) + log_softmax (:42-44),
// val_fc1 (2HW -> 64) + ReLU, val_fc2 (64 -> 1), tanh (:48-51), applied to the 6 head features per square
// that the last trunk layer's epilogue (rz_net_conv3x3_tc2_head / _tc3_head) or rz_net_head_features wrote.
//
// The CUDA-core heads kernel (rz_net_heads.cu) spends 211 us on 8192 boards of 15x15 (3.8 GFLOP of fp32 FMAs
// through shared memory) and 721 us at 19x19.  Here the two FCs are one GEMM per 128 boards:
//   M = 128 boards, K = the 6*P features of a board in the PADDED square order k' = f*P + y*S + x -- exactly the
//       row feat[b][.] the trunk wrote, so building the A operand is a contiguous copy, and the weight matrix
//       carries zeros at the padding squares;
//   N = AS policy outputs for k' < 4P (filters 0..3), 64 value-hidden outputs for 4P <= k' < 6P (filters 4, 5).
// float32 accuracy on the bf16 tensor pipe: features and weights are split x = hi + lo (two bf16, lo = the
// rounding error of hi), and every k-step issues hi*hi + lo*hi + hi*lo (the dropped lo*lo term is 2^-18
// relative), so logits agree with the fp32 FC to ~1e-6 -- far inside the 1e-3 budget of the bf16 trunk.
// Accumulators live in TMEM: columns [0, AS) policy logits, [AS, AS+64) value hidden units; the epilogue
// (one thread per board row) adds the biases, does log_softmax / ReLU-FC2-tanh straight from TMEM.
#include <cuda_bf16.h>

#include "rz_common.cuh"
#include "rz_tc.cuh"

namespace {

constexpr int HT_THREADS = 256;
constexpr int HT_A_BYTES = 128 * 128;            // [128 boards][64 k] bf16, SW128
constexpr int HT_CTRL_BYTES = 4096;

struct HeadsTcParams {
  const float* feat;      // [n][6][P]
  const float* bp;        // [AS]
  const float* bv1;       // [64]
  const float* wv2;       // [64]
  const float* bv2;       // [1]
  float* logp;            // [n][AS]
  float* value;           // [n]
  int n_boards, A, AS, P;
  int n_stages;           // 1 or 2 shared-memory stages
  int w_box;              // rows per TMA box of the policy weights (divides AS)
  int tmem_cols;          // power of two >= AS + 64
};

// stage layout: A_hi | A_lo | W_hi [WR rows] | W_lo [WR rows], WR = max(AS, 64): the value chunks load 64 rows
__host__ __device__ __forceinline__ uint32_t ht_w_rows(int AS) { return (uint32_t)(AS > 64 ? AS : 64); }
__host__ __device__ __forceinline__ uint32_t ht_stage_bytes(int AS) { return 2u * HT_A_BYTES + 2u * ht_w_rows(AS) * 128u; }

__global__ void __launch_bounds__(HT_THREADS, 1)
rz_heads_tc_kernel(const __grid_constant__ CUtensorMap tmap_whi, const __grid_constant__ CUtensorMap tmap_wlo,
                   const __grid_constant__ CUtensorMap tmap_vhi, const __grid_constant__ CUtensorMap tmap_vlo,
                   const HeadsTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (rz::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - rz::smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int AS = p.AS, P = p.P;
  const uint32_t stage_bytes = ht_stage_bytes(AS);
  const uint32_t ctrl = base + (uint32_t)p.n_stages * stage_bytes;
  uint8_t* ctrl_ptr = al + (size_t)p.n_stages * stage_bytes;
  const uint32_t bar_w = ctrl, bar_free = ctrl + 16, bar_done = ctrl + 32;   // [2], [2], [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(ctrl_ptr + 40);
  float* s_bp = reinterpret_cast<float*>(ctrl_ptr + 64);        // [AS <= 384]
  float* s_bv1 = s_bp + 384;                                    // [64]
  float* s_wv2 = s_bv1 + 64;                                    // [64]

  if (tid == 0) {
    rz::tma_prefetch_desc(&tmap_whi);
    rz::tma_prefetch_desc(&tmap_wlo);
    rz::tma_prefetch_desc(&tmap_vhi);
    rz::tma_prefetch_desc(&tmap_vlo);
    for (int s = 0; s < 2; ++s) { rz::mbar_init(bar_w + 8 * s, 1); rz::mbar_init(bar_free + 8 * s, 1); }
    rz::mbar_init(bar_done, 1);
    rz::fence_barrier_init();
  }
  if (warp == 0) { rz::tmem_alloc(rz::smem_u32(tmem_holder), (uint32_t)p.tmem_cols); rz::tmem_relinquish(); }
  for (int i = tid; i < AS; i += HT_THREADS) s_bp[i] = i < p.A ? p.bp[i] : 0.0f;
  if (tid < 64) { s_bv1[tid] = p.bv1[tid]; s_wv2[tid] = p.wv2[tid]; }
  rz::tc_fence_before();
  __syncthreads();
  rz::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int b0 = blockIdx.x * 128;
  const int r = tid & 127, khalf = tid >> 7;          // this thread builds row r, k in [32*khalf, 32*khalf + 32)
  const bool row_live = b0 + r < p.n_boards;
  const int K6 = 6 * P, K4 = 4 * P;
  const int n_chunks = (K6 + 63) >> 6, n_pol = K4 >> 6;   // 4P is a multiple of 64 for S in {8, 16, 20}
  const float* frow = p.feat + (size_t)(b0 + r) * K6 + 32 * khalf;

  float4 pre[8];
  auto prefetch = [&](int c) {
    const int k0 = c * 64 + 32 * khalf;
    const bool live = row_live && c < n_chunks && k0 < K6;   // K6 is a multiple of 32
    const float4* src = reinterpret_cast<const float4*>(frow + (size_t)c * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) pre[j] = live ? __ldg(src + j) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  };
  prefetch(0);

  const uint32_t idesc_v = rz::umma_idesc_bf16(128, 64);
  for (int c = 0; c < n_chunks; ++c) {
    const int s = c % p.n_stages, use = c / p.n_stages;
    const uint32_t st = base + (uint32_t)s * stage_bytes;
    const uint32_t a_hi = st, a_lo = st + HT_A_BYTES, w_hi = st + 2 * HT_A_BYTES, w_lo = w_hi + ht_w_rows(AS) * 128u;
    const bool policy = c < n_pol;
    // the MMAs of the chunk that used this stage before have read it
    if (use > 0) rz::mbar_wait(bar_free + 8 * s, (uint32_t)(use - 1) & 1u);
    if (tid == 0) {
      if (policy) {
        rz::mbar_expect_tx(bar_w + 8 * s, 2u * (uint32_t)AS * 128u);
        for (int j = 0; j < AS; j += p.w_box) {
          rz::tma_load_2d(w_hi + (uint32_t)j * 128u, &tmap_whi, bar_w + 8 * s, c * 64, j);
          rz::tma_load_2d(w_lo + (uint32_t)j * 128u, &tmap_wlo, bar_w + 8 * s, c * 64, j);
        }
      } else {
        rz::mbar_expect_tx(bar_w + 8 * s, 2u * 64u * 128u);
        rz::tma_load_2d(w_hi, &tmap_vhi, bar_w + 8 * s, c * 64, 0);
        rz::tma_load_2d(w_lo, &tmap_vlo, bar_w + 8 * s, c * 64, 0);
      }
    }
    // this thread's 32 features -> hi / lo bf16, 4 + 4 swizzled 16-byte chunks of row r
    {
      const uint32_t row_off = (uint32_t)r * 128u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x[8] = {pre[2 * q].x, pre[2 * q].y, pre[2 * q].z, pre[2 * q].w,
                            pre[2 * q + 1].x, pre[2 * q + 1].y, pre[2 * q + 1].z, pre[2 * q + 1].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * e]), h1 = __float2bfloat16_rn(x[2 * e + 1]);
          hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[e] = rz::pack_bf16x2(x[2 * e] - __bfloat162float(h0), x[2 * e + 1] - __bfloat162float(h1));
        }
        const uint32_t chunk = (uint32_t)(khalf * 4 + q);
        const uint32_t off = row_off + ((chunk ^ ((uint32_t)r & 7u)) << 4);
        rz::st_shared_v4(a_hi + off, hi[0], hi[1], hi[2], hi[3]);
        rz::st_shared_v4(a_lo + off, lo[0], lo[1], lo[2], lo[3]);
      }
    }
    prefetch(c + 1);                 // the next chunk's loads fly while this chunk's MMAs are issued
    rz::fence_proxy_async();
    rz::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      rz::tc_fence_after();
      rz::mbar_wait(bar_w + 8 * s, (uint32_t)use & 1u);
      const uint64_t d_ahi = rz::umma_desc_sw128(a_hi), d_alo = rz::umma_desc_sw128(a_lo);
      if (policy) {
        // N = AS in pieces of at most 256 columns (AS = 384 for 19x19 Go: 256 + 128)
        for (int n0 = 0; n0 < AS; n0 += 256) {
          const int nn = min(256, AS - n0);
          const uint32_t idesc = rz::umma_idesc_bf16(128, nn);
          const uint64_t d_whi = rz::umma_desc_sw128(w_hi + (uint32_t)n0 * 128u);
          const uint64_t d_wlo = rz::umma_desc_sw128(w_lo + (uint32_t)n0 * 128u);
          const uint32_t d = tmem_base + (uint32_t)n0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            rz::umma_bf16(d, d_ahi + (uint64_t)(2 * kk), d_whi + (uint64_t)(2 * kk), idesc, (c > 0 || kk > 0) ? 1u : 0u);
            rz::umma_bf16(d, d_alo + (uint64_t)(2 * kk), d_whi + (uint64_t)(2 * kk), idesc, 1u);
            rz::umma_bf16(d, d_ahi + (uint64_t)(2 * kk), d_wlo + (uint64_t)(2 * kk), idesc, 1u);
          }
        }
      } else {
        const uint64_t d_whi = rz::umma_desc_sw128(w_hi), d_wlo = rz::umma_desc_sw128(w_lo);
        const uint32_t d = tmem_base + (uint32_t)AS;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          rz::umma_bf16(d, d_ahi + (uint64_t)(2 * kk), d_whi + (uint64_t)(2 * kk), idesc_v, (c > n_pol || kk > 0) ? 1u : 0u);
          rz::umma_bf16(d, d_alo + (uint64_t)(2 * kk), d_whi + (uint64_t)(2 * kk), idesc_v, 1u);
          rz::umma_bf16(d, d_ahi + (uint64_t)(2 * kk), d_wlo + (uint64_t)(2 * kk), idesc_v, 1u);
        }
      }
      rz::umma_commit(bar_free + 8 * s);
      if (c == n_chunks - 1) rz::umma_commit(bar_done);
    }
  }
  rz::mbar_wait(bar_done, 0);
  rz::tc_fence_after();

  // ---- epilogue: warps 0..3 = log_softmax of board row (warp*32 + lane); warps 4..7 = the value of the same rows
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const bool live = b0 + row < p.n_boards;
  const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
  if (warp < 4) {
    // act_fc1 bias + log_softmax over the A real outputs (:43-44); three sweeps over the TMEM row
    float mx = -3.0e38f;
    for (int c0 = 0; c0 < AS; c0 += 32) {
      uint32_t acc[32];
      rz::tmem_ld_32x32(t_row + (uint32_t)c0, acc);
      rz::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < p.A) mx = fmaxf(mx, __uint_as_float(acc[j]) + s_bp[c0 + j]);
    }
    float sum = 0.0f;
    for (int c0 = 0; c0 < AS; c0 += 32) {
      uint32_t acc[32];
      rz::tmem_ld_32x32(t_row + (uint32_t)c0, acc);
      rz::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < p.A) sum += expf((__uint_as_float(acc[j]) + s_bp[c0 + j]) - mx);
    }
    const float lse = mx + logf(sum);
    float* lp = p.logp + (size_t)(b0 + row) * AS;
    for (int c0 = 0; c0 < AS; c0 += 32) {
      uint32_t acc[32];
      rz::tmem_ld_32x32(t_row + (uint32_t)c0, acc);
      rz::tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = j8 * 8 + e;
            o[e] = c0 + j < p.A ? __float_as_uint((__uint_as_float(acc[j]) + s_bp[c0 + j]) - lse) : 0u;
          }
          rz::st_global_v8(lp + c0 + j8 * 8, o);
        }
      }
    }
  } else {
    // val_fc1 bias + ReLU, val_fc2, tanh (:49-51)
    float hv = 0.0f;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t acc[32];
      rz::tmem_ld_32x32(t_row + (uint32_t)(AS + c0), acc);
      rz::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        hv = fmaf(fmaxf(__uint_as_float(acc[j]) + s_bv1[c0 + j], 0.0f), s_wv2[c0 + j], hv);
    }
    if (live) p.value[b0 + row] = tanhf(hv + p.bv2[0]);
  }
  rz::tc_fence_before();
  __syncthreads();
  if (warp == 0) { rz::tc_fence_after(); rz::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

}  // namespace

extern "C" int rz_net_heads_tc(const rz_heads_desc* h, const float* feat, float* logp, float* value,
                               int n_boards, void* stream) {
  RZ_REQUIRE(h && feat && logp && value, "rz_net_heads_tc: null argument");
  RZ_REQUIRE(h->wtc_hi && h->wtc_lo && h->bp && h->bv1 && h->wv2 && h->bv2, "rz_net_heads_tc: null weight pointer");
  RZ_REQUIRE(h->board_size >= 1 && h->board_size <= RZ_MAX_BOARD, "rz_net_heads_tc: board_size %d", h->board_size);
  const int W = h->width > 0 ? h->width : h->board_size;
  RZ_REQUIRE(W >= 1 && W <= RZ_MAX_BOARD, "rz_net_heads_tc: width %d", W);
  const int A = h->n_actions > 0 ? h->n_actions : h->board_size * W;
  const int AS = h->action_stride;
  RZ_REQUIRE(AS >= A && (AS & 31) == 0 && AS <= 384, "rz_net_heads_tc: action_stride %d", AS);
  const int S = rz_row_stride(h->board_size, W, h->row_stride);
  RZ_REQUIRE(S != 0, "rz_net_heads_tc: row_stride %d does not hold a %dx%d board", h->row_stride, h->board_size, W);
  RZ_REQUIRE(n_boards >= 0, "rz_net_heads_tc: n_boards %d", n_boards);
  if (n_boards == 0) return 0;
  const int P = S * S;
  const int KP = (6 * P + 63) / 64 * 64;
  HeadsTcParams p;
  p.feat = feat; p.bp = h->bp; p.bv1 = h->bv1; p.wv2 = h->wv2; p.bv2 = h->bv2; p.logp = logp; p.value = value;
  p.n_boards = n_boards; p.A = A; p.AS = AS; p.P = P;
  p.w_box = (AS % 128 == 0) ? 128 : ((AS % 64 == 0) ? 64 : 32);
  p.tmem_cols = 32;
  while (p.tmem_cols < AS + 64) p.tmem_cols *= 2;
  const size_t stage = ht_stage_bytes(AS);
  p.n_stages = (2 * stage + HT_CTRL_BYTES + 1024 <= 232448) ? 2 : 1;
  const size_t smem = (size_t)p.n_stages * stage + HT_CTRL_BYTES + 1024;
  // weights: bf16 [AS + 64][KP] K-major, rows 0..AS-1 the policy outputs, AS..AS+63 the value hidden units
  const __nv_bfloat16* whi = reinterpret_cast<const __nv_bfloat16*>(h->wtc_hi);
  const __nv_bfloat16* wlo = reinterpret_cast<const __nv_bfloat16*>(h->wtc_lo);
  CUtensorMap t_whi, t_wlo, t_vhi, t_vlo;
  if (rz::make_tmap_2d(&t_whi, whi, (uint64_t)AS, (uint64_t)KP, (uint32_t)p.w_box)) return -1;
  if (rz::make_tmap_2d(&t_wlo, wlo, (uint64_t)AS, (uint64_t)KP, (uint32_t)p.w_box)) return -1;
  if (rz::make_tmap_2d(&t_vhi, whi + (size_t)AS * KP, 64, (uint64_t)KP, 64)) return -1;
  if (rz::make_tmap_2d(&t_vlo, wlo + (size_t)AS * KP, 64, (uint64_t)KP, 64)) return -1;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(rz_heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rz_set_error("rz_net_heads_tc: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    attr_smem = smem;
  }
  const int grid = (n_boards + 127) / 128;
  rz_heads_tc_kernel<<<grid, HT_THREADS, smem, (cudaStream_t)stream>>>(t_whi, t_wlo, t_vhi, t_vlo, p);
  RZ_LAUNCH_CHECK("rz_net_heads_tc");
  return 0;
}
