// rlzero_b200 -- the training step of the reference's own network, written by hand: forward with saved
// activations, analytic backward pass, loss / entropy, Adam.  No autograd, no cuDNN, no cuBLAS.
//
// Reference: AlphaZeroAgent.learn (rlzero/games/gomoku/alphazero_agent.py:59-86) on PolicyValueNet
// (rlzero/games/gomoku/policy_value_net.py:6-52):
//     loss    = mse(v, z) - mean_b sum_a pi[b,a] * log p[b,a]          (:70-75; L2 lives in Adam's weight_decay)
//     entropy = -mean_b sum_a p * log p                                (:83-85)
//     Adam(lr, betas = (0.9, 0.999), eps = 1e-8, weight_decay = 1e-4)  (:22-24; torch.optim.Adam semantics)
// Oracle: oracle/train_oracle.py (numpy float64, pinned to autograd and to the live reference agent).
//
// Everything here is float32 on CUDA cores and DETERMINISTIC: every reduction over the batch runs in a fixed order
// (partials per block, then a second pass), no atomics.  Activations are channels-last [n][HW][C] like the fp32
// inference path (rz_net_conv3x3_f32 is the forward convolution AND, with flipped/transposed weights, the data
// gradient); gradients of the parameters are written in the PyTorch layouts of the state_dict, so that the flat
// parameter / gradient / moment buffers the host keeps are the module's own tensors.
#include "rz_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// C[m][n] = alpha * sum_k A(m,k) * B(k,n) (+ C[m][n] if accumulate), A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
// 64x64 tile per block, 16x16 threads, 4x4 outputs per thread, k ascending: deterministic.
// ---------------------------------------------------------------------------------------------------------
constexpr int SG_KT = 32;
__global__ void __launch_bounds__(256)
rz_sgemm_kernel(int M, int N, int K, const float* __restrict__ A, long long sam, long long sak,
                const float* __restrict__ B, long long sbk, long long sbn, float* __restrict__ Cm, long long ldc,
                float alpha, int accumulate) {
  __shared__ float As[SG_KT][64 + 1];
  __shared__ float Bs[SG_KT][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  // consecutive threads read along whichever index is contiguous in memory (coalesced either way); the odd row
  // length of the shared tiles keeps both write patterns conflict-free
  const bool a_k = sak == 1, b_k = sbk == 1;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_KT) {
    for (int i = threadIdx.x; i < SG_KT * 64; i += 256) {
      int kk = a_k ? (i % SG_KT) : (i >> 6), mm = a_k ? (i / SG_KT) : (i & 63);
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[(long long)m * sam + (long long)k * sak] : 0.0f;
      kk = b_k ? (i % SG_KT) : (i >> 6); mm = b_k ? (i / SG_KT) : (i & 63);
      const int n = n0 + mm;
      k = k0 + kk;
      Bs[kk][mm] = (n < N && k < K) ? B[(long long)k * sbk + (long long)n * sbn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_KT; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty + 16 * i]; b[i] = Bs[kk][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < M && n < N) {
        float* c = Cm + (long long)m * ldc + n;
        *c = alpha * acc[i][j] + (accumulate ? *c : 0.0f);
      }
    }
}

// column sums of a [rows][cols] matrix (row stride ld), two deterministic passes:
//   part[s][c] = sum over the rows of slice s; out[c] = sum_s part[s][c]
__global__ void __launch_bounds__(256)
rz_colsum_part_kernel(const float* __restrict__ in, long long rows, int cols, long long ld, float* __restrict__ part,
                      int n_slices) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5, s = blockIdx.y;
  const long long per = (rows + n_slices - 1) / n_slices;
  const long long r0 = s * per, r1 = min(rows, r0 + per);
  float acc = 0.0f;
  if (c < cols)
    for (long long r = r0 + ry; r < r1; r += 8) acc += in[r * ld + c];
  red[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    part[(long long)s * cols + c] = t;
  }
}
__global__ void rz_colsum_final_kernel(const float* __restrict__ part, int n_slices, int cols, float* __restrict__ out,
                                       float alpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float t = 0.0f;
  for (int s = 0; s < n_slices; ++s) t += part[(long long)s * cols + c];
  out[c] = alpha * t;
}

// ---------------------------------------------------------------------------------------------------------
// convolution weights: PyTorch [cout][cin][3][3] -> forward [tap][cin][cout] and data-gradient
// [8 - tap][cout][cin] (the transposed convolution is a convolution with the taps mirrored and the channel roles
// swapped: dx[p][ci] = sum_tap sum_co dy[p - d(tap)][co] * w[co][ci][tap])
// ---------------------------------------------------------------------------------------------------------
__global__ void rz_pack_conv_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wb,
                                    int cin, int cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 9) return;
  const int tap = i % 9, ci = (i / 9) % cin, co = i / (9 * cin);
  const float v = w[i];
  if (wf) wf[((size_t)tap * cin + ci) * cout + co] = v;
  if (wb) wb[((size_t)(8 - tap) * cout + co) * cin + ci] = v;
}

// g *= (a > 0): the ReLU mask
__global__ void rz_relu_bwd_kernel(const float* __restrict__ a, float* __restrict__ g, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(a[i] > 0.0f)) g[i] = 0.0f;
}

// [n][C][HW] -> [n][HW][C] (observation planes to channels-last)
__global__ void rz_nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int C, int HW) {
  const long long total = (long long)n * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int pos = (int)(r % HW), b = (int)(r / HW);
    out[i] = in[((long long)b * C + c) * HW + pos];
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient of the 3x3 convolution: part[s][tap][ci][co] = sum over the (board, square) rows of slice s of
// x[board][square + d(tap)][ci] * dz[board][square][co]   (zero outside the board)
// ---------------------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 16;
__global__ void __launch_bounds__(256)
rz_conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ part, int n_boards,
                     int H, int cin, int cout, int n_slices) {
  extern __shared__ float sm[];
  float* xs = sm;                       // [WG_ROWS][cin]
  float* ds = sm + WG_ROWS * cin;       // [WG_ROWS][cout]
  const int tap = blockIdx.x, s = blockIdx.y;
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  const int HW = H * H;
  const long long rows = (long long)n_boards * HW;
  const long long per = ((rows + n_slices - 1) / n_slices + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  const long long r0 = s * per, r1 = min(rows, r0 + per);
  const int ci0 = threadIdx.x >> 5, co0 = threadIdx.x & 31;     // thread tile: ci = ci0 + 8 i, co = co0 + 32 j
  float acc[8][4] = {};
  for (long long rb = r0; rb < r1; rb += WG_ROWS) {
    for (int i = threadIdx.x; i < WG_ROWS * cin; i += 256) {
      const int rr = i / cin, c = i - rr * cin;
      const long long r = rb + rr;
      float v = 0.0f;
      if (r < r1) {
        const int b = (int)(r / HW), pos = (int)(r - (long long)b * HW);
        const int yy = pos / H + dy, xx = pos % H + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < H) v = x[((long long)b * HW + yy * H + xx) * cin + c];
      }
      xs[i] = v;
    }
    for (int i = threadIdx.x; i < WG_ROWS * cout; i += 256) {
      const int rr = i / cout;
      const long long r = rb + rr;
      ds[i] = r < r1 ? dz[r * cout + (i - rr * cout)] : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < WG_ROWS; ++rr) {
      float xv[8], dv[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = (ci0 + 8 * i < cin) ? xs[rr * cin + ci0 + 8 * i] : 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) dv[j] = (co0 + 32 * j < cout) ? ds[rr * cout + co0 + 32 * j] : 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], dv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + ((size_t)s * 9 + tap) * cin * cout;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + 8 * i, co = co0 + 32 * j;
      if (ci < cin && co < cout) out[(size_t)ci * cout + co] = acc[i][j];
    }
}
// dw[co][ci][tap] = sum_s part[s][tap][ci][co]
__global__ void rz_conv_wgrad_final_kernel(const float* __restrict__ part, float* __restrict__ dw, int n_slices, int cin,
                                           int cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 9) return;
  const int tap = i % 9, ci = (i / 9) % cin, co = i / (9 * cin);
  float t = 0.0f;
  for (int s = 0; s < n_slices; ++s) t += part[(((size_t)s * 9 + tap) * cin + ci) * cout + co];
  dw[i] = t;
}

// ---------------------------------------------------------------------------------------------------------
// heads.  feat[b][f][pos] = relu(sum_c a3[b][pos][c] * w1x1[f][c] + b1x1[f]), f < 4: act_conv1, f >= 4: val_conv1
// (policy_value_net.py:41,47); the FC inputs x.view(-1, 4*H*W) / (-1, 2*H*W) are feat[b][0:4] / feat[b][4:6] as they lie
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rz_head_feat_fwd_kernel(const float* __restrict__ a3, const float* __restrict__ w1x1, const float* __restrict__ b1x1,
                        float* __restrict__ feat, int n_boards, int HW) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)n_boards * HW) return;
  const float4 a = reinterpret_cast<const float4*>(a3 + row * 128)[lane];
  float s[6];
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const float4 w = reinterpret_cast<const float4*>(w1x1 + f * 128)[lane];
    s[f] = fmaf(a.w, w.w, fmaf(a.z, w.z, fmaf(a.y, w.y, a.x * w.x)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int f = 0; f < 6; ++f) s[f] += __shfl_xor_sync(RZ_FULL, s[f], o);
  if (lane < 6) {
    const int b = (int)(row / HW), pos = (int)(row - (long long)b * HW);
    float v = 0.0f;
#pragma unroll
    for (int f = 0; f < 6; ++f) if (lane == f) v = s[f];
    feat[((size_t)b * 6 + lane) * HW + pos] = fmaxf(v + b1x1[lane], 0.0f);
  }
}

// logp[b][a] = log_softmax(logits[b][:] + bp)[a]  (in place on `logits`, padding columns set to 0)
__global__ void __launch_bounds__(128)
rz_logsoftmax_kernel(float* __restrict__ logits, const float* __restrict__ bp, int A, int AS) {
  __shared__ float red[4];
  float* row = logits + (size_t)blockIdx.x * AS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -3.0e38f;
  for (int a = threadIdx.x; a < A; a += 128) { row[a] += bp[a]; mx = fmaxf(mx, row[a]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(RZ_FULL, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.0f;
  for (int a = threadIdx.x; a < A; a += 128) sum += expf(row[a] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(RZ_FULL, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float lse = mx + logf((red[0] + red[1]) + (red[2] + red[3]));
  for (int a = threadIdx.x; a < AS; a += 128) row[a] = a < A ? row[a] - lse : 0.0f;
}

// h[b][j] = relu(hpre[b][j] + bv1[j]) (in place), v[b] = tanh(sum_j h[b][j] * wv2[j] + bv2)   (policy_value_net.py:49-51)
__global__ void __launch_bounds__(64)
rz_value_fwd_kernel(float* __restrict__ h, const float* __restrict__ bv1, const float* __restrict__ wv2,
                    const float* __restrict__ bv2, float* __restrict__ v) {
  __shared__ float red[2];
  const int b = blockIdx.x, j = threadIdx.x;
  const float hv = fmaxf(h[(size_t)b * 64 + j] + bv1[j], 0.0f);
  h[(size_t)b * 64 + j] = hv;
  float s = hv * wv2[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(RZ_FULL, s, o);
  if ((j & 31) == 0) red[j >> 5] = s;
  __syncthreads();
  if (j == 0) v[b] = tanhf(red[0] + red[1] + bv2[0]);
}

// per sample: the loss terms and the gradients at the outputs (train_oracle.loss_and_grads):
//   dlogits = (softmax * sum_a pi - pi) / B,   dpre2 = 2 (v - z) / B * (1 - v^2),   dh = dpre2 * wv2 * (h > 0)
//   terms[b] = ((v - z)^2, -sum pi log p, -sum p log p)
__global__ void __launch_bounds__(128)
rz_loss_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ pi, int pi_stride,
                   const float* __restrict__ v, const float* __restrict__ z, const float* __restrict__ h,
                   const float* __restrict__ wv2, float* __restrict__ dlogits, float* __restrict__ dpre2,
                   float* __restrict__ dh, float* __restrict__ terms, int B, int A, int AS) {
  __shared__ float red[3][4];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* lp = logp + (size_t)b * AS;
  const float* pr = pi + (size_t)b * pi_stride;
  float s_pi = 0.0f, s_pl = 0.0f, s_ent = 0.0f;
  for (int a = threadIdx.x; a < A; a += 128) {
    const float l = lp[a], p = expf(l), t = pr[a];
    s_pi += t; s_pl += t * l; s_ent += p * l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_pi += __shfl_xor_sync(RZ_FULL, s_pi, o);
    s_pl += __shfl_xor_sync(RZ_FULL, s_pl, o);
    s_ent += __shfl_xor_sync(RZ_FULL, s_ent, o);
  }
  if (lane == 0) { red[0][warp] = s_pi; red[1][warp] = s_pl; red[2][warp] = s_ent; }
  __syncthreads();
  s_pi = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
  s_pl = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  s_ent = (red[2][0] + red[2][1]) + (red[2][2] + red[2][3]);
  const float invB = 1.0f / (float)B;
  for (int a = threadIdx.x; a < AS; a += 128)
    dlogits[(size_t)b * AS + a] = a < A ? (expf(lp[a]) * s_pi - pr[a]) * invB : 0.0f;
  const float vv = v[b], d = vv - z[b];
  const float dp = 2.0f * d * invB * (1.0f - vv * vv);
  if (threadIdx.x < 64) dh[(size_t)b * 64 + threadIdx.x] = h[(size_t)b * 64 + threadIdx.x] > 0.0f ? dp * wv2[threadIdx.x] : 0.0f;
  if (threadIdx.x == 0) {
    dpre2[b] = dp;
    terms[(size_t)b * 3 + 0] = d * d;
    terms[(size_t)b * 3 + 1] = -s_pl;
    terms[(size_t)b * 3 + 2] = -s_ent;
  }
}

// da3[b][pos][c] = sum_f dfeat[b][f][pos] * w1x1[f][c]  (dfeat already masked by feat > 0); per block of 64 rows the
// partial sums of dW1x1[f][c] = sum_rows dfeat * a3 and db1x1[f] = sum_rows dfeat: part[blk][6][128 + 1]
__global__ void __launch_bounds__(128)
rz_head_feat_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ a3, const float* __restrict__ w1x1,
                        float* __restrict__ da3, float* __restrict__ part, int n_boards, int HW) {
  const int c = threadIdx.x;
  const long long rows = (long long)n_boards * HW;
  const long long r0 = (long long)blockIdx.x * 64, r1 = min(rows, r0 + 64);
  float w[6], gw[6] = {}, gb[6] = {};
#pragma unroll
  for (int f = 0; f < 6; ++f) w[f] = w1x1[f * 128 + c];
  for (long long r = r0; r < r1; ++r) {
    const int b = (int)(r / HW), pos = (int)(r - (long long)b * HW);
    const float a = a3[r * 128 + c];
    float d = 0.0f;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const float g = dfeat[((size_t)b * 6 + f) * HW + pos];
      d = fmaf(g, w[f], d);
      gw[f] = fmaf(g, a, gw[f]);
      gb[f] += g;
    }
    da3[r * 128 + c] = d;
  }
  float* out = part + (size_t)blockIdx.x * (6 * 129);
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    out[f * 129 + c] = gw[f];
    if (c == 0) out[f * 129 + 128] = gb[f];
  }
}

// torch.optim.Adam, one flat buffer (single-tensor semantics): g += wd p; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g g;
// p -= step_size * m / (sqrt(v) / bc2_sqrt + eps)
__global__ void rz_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                               float* __restrict__ v, long long n, float b1, float b2, float eps, float wd,
                               float step_size, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pw = p[i];
    const float grad = fmaf(wd, pw, g[i]);
    const float mm = fmaf(1.0f - b1, grad - m[i], m[i]);               // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = fmaf((1.0f - b2) * grad, grad, b2 * v[i]);        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mm; v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = pw - step_size * (mm / denom);
  }
}

inline int grid_for(long long n, int threads = 256, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// =========================================================================================================
// C ABI (include/rlzero_b200.h, "training step")
// =========================================================================================================
extern "C" int rz_learn_sgemm(int M, int N, int K, const float* A, long long sam, long long sak, const float* B,
                              long long sbk, long long sbn, float* Cm, long long ldc, float alpha, int accumulate,
                              void* stream) {
  RZ_REQUIRE(A && B && Cm && M >= 0 && N >= 0 && K >= 0, "rz_learn_sgemm: bad arguments");
  if (M == 0 || N == 0) return 0;
  rz_sgemm_kernel<<<dim3((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64)), 256, 0, (cudaStream_t)stream>>>(
      M, N, K, A, sam, sak, B, sbk, sbn, Cm, ldc, alpha, accumulate);
  RZ_LAUNCH_CHECK("rz_learn_sgemm");
  return 0;
}

extern "C" int rz_learn_colsum(const float* in, long long rows, int cols, long long ld, float* out, float alpha,
                               float* scratch, int n_slices, void* stream) {
  RZ_REQUIRE(in && out && scratch && rows >= 0 && cols >= 1 && n_slices >= 1, "rz_learn_colsum: bad arguments");
  rz_colsum_part_kernel<<<dim3((unsigned)((cols + 31) / 32), (unsigned)n_slices), 256, 0, (cudaStream_t)stream>>>(
      in, rows, cols, ld, scratch, n_slices);
  rz_colsum_final_kernel<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scratch, n_slices, cols, out, alpha);
  RZ_LAUNCH_CHECK("rz_learn_colsum");
  return 0;
}

extern "C" int rz_learn_pack_conv(const float* w_oihw, float* w_fwd, float* w_bwd, int c_in, int c_out, void* stream) {
  RZ_REQUIRE(w_oihw && (w_fwd || w_bwd) && c_in >= 1 && c_out >= 1, "rz_learn_pack_conv: bad arguments");
  rz_pack_conv_kernel<<<(c_out * c_in * 9 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_oihw, w_fwd, w_bwd, c_in, c_out);
  RZ_LAUNCH_CHECK("rz_learn_pack_conv");
  return 0;
}

extern "C" int rz_learn_relu_bwd(const float* act, float* grad, long long n, void* stream) {
  RZ_REQUIRE(act && grad && n >= 0, "rz_learn_relu_bwd: bad arguments");
  if (n == 0) return 0;
  rz_relu_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(act, grad, n);
  RZ_LAUNCH_CHECK("rz_learn_relu_bwd");
  return 0;
}

extern "C" int rz_learn_nchw_to_nhwc(const float* in, float* out, int n, int channels, int hw, void* stream) {
  RZ_REQUIRE(in && out && n >= 0 && channels >= 1 && hw >= 1, "rz_learn_nchw_to_nhwc: bad arguments");
  if (n == 0) return 0;
  rz_nchw_to_nhwc_kernel<<<grid_for((long long)n * channels * hw), 256, 0, (cudaStream_t)stream>>>(in, out, n, channels, hw);
  RZ_LAUNCH_CHECK("rz_learn_nchw_to_nhwc");
  return 0;
}

extern "C" int rz_learn_conv_wgrad(const float* x, const float* dz, float* dw_oihw, float* db, float* scratch,
                                   long long scratch_floats, int n_boards, int board_size, int c_in, int c_out,
                                   void* stream) {
  RZ_REQUIRE(x && dz && dw_oihw && db && scratch, "rz_learn_conv_wgrad: null argument");
  RZ_REQUIRE(c_in >= 1 && c_in <= 64 && c_out >= 1 && c_out <= 128, "rz_learn_conv_wgrad: channels %d -> %d (<= 64 -> <= 128)",
             c_in, c_out);
  RZ_REQUIRE(board_size >= 1 && board_size <= RZ_MAX_BOARD && n_boards >= 1, "rz_learn_conv_wgrad: bad sizes");
  const long long rows = (long long)n_boards * board_size * board_size;
  int slices = (int)((rows + 511) / 512);
  if (slices > 96) slices = 96;
  const long long need = (long long)slices * 9 * c_in * c_out;
  RZ_REQUIRE(scratch_floats >= need && scratch_floats >= (long long)slices * c_out,
             "rz_learn_conv_wgrad: scratch holds %lld floats, %lld needed", scratch_floats, need);
  const size_t smem = sizeof(float) * WG_ROWS * (size_t)(c_in + c_out);
  cudaStream_t st = (cudaStream_t)stream;
  rz_conv_wgrad_kernel<<<dim3(9, (unsigned)slices), 256, smem, st>>>(x, dz, scratch, n_boards, board_size, c_in, c_out, slices);
  rz_conv_wgrad_final_kernel<<<(c_out * c_in * 9 + 255) / 256, 256, 0, st>>>(scratch, dw_oihw, slices, c_in, c_out);
  // bias gradient: column sums of dz (re-uses the scratch after the weight partials have been folded)
  rz_colsum_part_kernel<<<dim3((unsigned)((c_out + 31) / 32), (unsigned)slices), 256, 0, st>>>(dz, rows, c_out, c_out, scratch, slices);
  rz_colsum_final_kernel<<<(c_out + 127) / 128, 128, 0, st>>>(scratch, slices, c_out, db, 1.0f);
  RZ_LAUNCH_CHECK("rz_learn_conv_wgrad");
  return 0;
}

extern "C" int rz_learn_head_feat_fwd(const float* a3, const float* w1x1, const float* b1x1, float* feat, int n_boards,
                                      int hw, void* stream) {
  RZ_REQUIRE(a3 && w1x1 && b1x1 && feat && n_boards >= 0 && hw >= 1, "rz_learn_head_feat_fwd: bad arguments");
  if (n_boards == 0) return 0;
  const long long rows = (long long)n_boards * hw;
  rz_head_feat_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a3, w1x1, b1x1, feat, n_boards, hw);
  RZ_LAUNCH_CHECK("rz_learn_head_feat_fwd");
  return 0;
}

extern "C" int rz_learn_logsoftmax(float* logits, const float* bias, int n, int n_actions, int action_stride,
                                   void* stream) {
  RZ_REQUIRE(logits && bias && n >= 0 && n_actions >= 1 && action_stride >= n_actions, "rz_learn_logsoftmax: bad arguments");
  if (n == 0) return 0;
  rz_logsoftmax_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(logits, bias, n_actions, action_stride);
  RZ_LAUNCH_CHECK("rz_learn_logsoftmax");
  return 0;
}

extern "C" int rz_learn_value_fwd(float* h, const float* bv1, const float* wv2, const float* bv2, float* v, int n,
                                  void* stream) {
  RZ_REQUIRE(h && bv1 && wv2 && bv2 && v && n >= 0, "rz_learn_value_fwd: bad arguments");
  if (n == 0) return 0;
  rz_value_fwd_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(h, bv1, wv2, bv2, v);
  RZ_LAUNCH_CHECK("rz_learn_value_fwd");
  return 0;
}

extern "C" int rz_learn_loss_bwd(const float* logp, const float* pi, int pi_stride, const float* v, const float* z,
                                 const float* h, const float* wv2, float* dlogits, float* dpre2, float* dh,
                                 float* terms, float* loss3, float* scratch, int n, int n_actions, int action_stride,
                                 void* stream) {
  RZ_REQUIRE(logp && pi && v && z && h && wv2 && dlogits && dpre2 && dh && terms && loss3 && scratch,
             "rz_learn_loss_bwd: null argument");
  RZ_REQUIRE(n >= 1 && n_actions >= 1 && action_stride >= n_actions && pi_stride >= n_actions, "rz_learn_loss_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  rz_loss_bwd_kernel<<<n, 128, 0, st>>>(logp, pi, pi_stride, v, z, h, wv2, dlogits, dpre2, dh, terms, n, n_actions,
                                        action_stride);
  // loss3 = (value_loss, policy_loss, entropy) = column means of terms
  rz_colsum_part_kernel<<<dim3(1, 8), 256, 0, st>>>(terms, n, 3, 3, scratch, 8);
  rz_colsum_final_kernel<<<1, 128, 0, st>>>(scratch, 8, 3, loss3, 1.0f / (float)n);
  RZ_LAUNCH_CHECK("rz_learn_loss_bwd");
  return 0;
}

extern "C" int rz_learn_head_feat_bwd(const float* dfeat, const float* a3, const float* w1x1, float* da3, float* dw1x1,
                                      float* db1x1, float* scratch, long long scratch_floats, int n_boards, int hw,
                                      void* stream) {
  RZ_REQUIRE(dfeat && a3 && w1x1 && da3 && dw1x1 && db1x1 && scratch && n_boards >= 1 && hw >= 1,
             "rz_learn_head_feat_bwd: bad arguments");
  const long long rows = (long long)n_boards * hw;
  const int blocks = (int)((rows + 63) / 64);
  const int slices = blocks < 64 ? blocks : 64;
  RZ_REQUIRE(scratch_floats >= (long long)blocks * 6 * 129 + (long long)slices * 6 * 129,
             "rz_learn_head_feat_bwd: scratch holds %lld floats", scratch_floats);
  cudaStream_t st = (cudaStream_t)stream;
  rz_head_feat_bwd_kernel<<<blocks, 128, 0, st>>>(dfeat, a3, w1x1, da3, scratch, n_boards, hw);
  // fold the per-block partials [blocks][6*129] -> [6*129], then split into dW1x1 [6][128] and db1x1 [6]
  float* part2 = scratch + (size_t)blocks * 6 * 129;
  rz_colsum_part_kernel<<<dim3((6 * 129 + 31) / 32, (unsigned)slices), 256, 0, st>>>(scratch, blocks, 6 * 129, 6 * 129, part2, slices);
  rz_colsum_final_kernel<<<(6 * 129 + 127) / 128, 128, 0, st>>>(part2, slices, 6 * 129, scratch, 1.0f);
  for (int f = 0; f < 6; ++f) {
    cudaMemcpyAsync(dw1x1 + f * 128, scratch + f * 129, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(db1x1 + f, scratch + f * 129 + 128, sizeof(float), cudaMemcpyDeviceToDevice, st);
  }
  RZ_LAUNCH_CHECK("rz_learn_head_feat_bwd");
  return 0;
}

extern "C" int rz_learn_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  RZ_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "rz_learn_adam: bad arguments");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  rz_adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                              weight_decay, (float)((double)lr / bc1), (float)sqrt(bc2));
  RZ_LAUNCH_CHECK("rz_learn_adam");
  return 0;
}
