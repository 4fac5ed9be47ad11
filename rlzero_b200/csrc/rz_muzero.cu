// rlzero_b200 -- MuZero search in latent space (BASELINE.json config 5).
//
// The reference contains NO MuZero code, so this follows the pseudocode published
// with the MuZero paper (Schrittwieser et al., "Mastering Atari, Go, chess and shogi by planning with
// a learned model", 2020, supplementary `pseudocode.py`): run_mcts / select_child / ucb_score /
// expand_node / backpropagate / add_exploration_noise / MinMaxStats, with the two-player convention of
// the public re-implementations for the value sign (a child's value enters its parent's score negated,
// and MinMaxStats is fed reward + discount * -value); rewards are 0 (board games), PARITY UNPINNED.
// oracle/muzero_oracle.py restates the same rules; tests compare both bit-exactly on replayed network
// outputs.
//
// Layout: like the AlphaZero trees (rz_tree.cu) one warp owns one tree and a node's children live in
// its edge block (N int32, W float64 = the child's value_sum from the child's own to_play view,
// P float32, child int32).  Every simulation expands exactly one node (there are no terminal states in
// latent space), so simulation i of EVERY tree creates node i+1: hidden states are stored node-major,
// pool[node][tree][position][channel], and the dynamics network writes slot i+1 of all trees as one
// contiguous tensor.  Its input is gathered from the parents' slots (rz_mz_gather), with the last
// channel replaced by the one-hot plane of the action.
#include <cuda_bf16.h>
#include <math.h>

#include "rz_common.cuh"

#define RZ_MZ_WARPS 4
#define RZ_MZ_THREADS (RZ_MZ_WARPS * 32)
#define RZ_MZ_MAX_ITERS 12

__device__ __forceinline__ size_t rz_mz_base(const rz_mz_desc& t, int g, int node) {
  return ((size_t)g * t.max_nodes + node) * (size_t)t.action_stride;
}

// MinMaxStats.normalize
__device__ __forceinline__ double rz_mz_normalize(double v, double mn, double mx) {
  return mx > mn ? __ddiv_rn(__dsub_rn(v, mn), __dsub_rn(mx, mn)) : v;
}

// ---------------------------------------------------------------------------
// expand_node(root, legal_actions, initial_inference) + add_exploration_noise
// ---------------------------------------------------------------------------
__device__ float rz_mz_gamma(float alpha, unsigned long long seed, uint32_t c0, uint32_t c1, uint32_t c2) {
  const float d = alpha + 1.0f - 1.0f / 3.0f;
  const float c = rsqrtf(9.0f * d);
  for (uint32_t it = 0; it < 64; ++it) {
    uint32_t r[4];
    rz_philox4(c0, c1, c2, it, seed, r);
    const float u1 = rz_u01_24(r[0]), u2 = rz_u01_24(r[1]);
    const float x = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    float v = 1.0f + c * x;
    if (v <= 0.0f) continue;
    v = v * v * v;
    const float u = rz_u01_24(r[2]);
    if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) return d * v * powf(rz_u01_24(r[3]), 1.0f / alpha);
  }
  return alpha;
}

__global__ void __launch_bounds__(RZ_MZ_THREADS)
rz_mz_root_kernel(rz_mz_desc t, const float* __restrict__ logp, const uint8_t* __restrict__ legal,
                  float noise_eps, float noise_alpha, unsigned long long seed, uint32_t move_id,
                  const int32_t* __restrict__ move_ids) {
  const int g = blockIdx.x * RZ_MZ_WARPS + (threadIdx.x >> 5);
  if (g >= t.n_trees) return;
  const int lane = rz_lane(), A = t.n_actions, AS = t.action_stride;
  const size_t nb = rz_mz_base(t, g, 0);
  if (move_ids) move_id = (uint32_t)move_ids[g];
  float p[RZ_MZ_MAX_ITERS], nz[RZ_MZ_MAX_ITERS];
  float sum = 0.0f, nsum = 0.0f;
#pragma unroll
  for (int i = 0; i < RZ_MZ_MAX_ITERS; ++i) {
    const int s = lane + 32 * i;
    p[i] = -1.0f; nz[i] = 0.0f;
    if (s < AS) {
      const bool ok = s < A && (!legal || legal[(size_t)g * A + s]);
      if (ok) {
        p[i] = expf(logp[(size_t)g * AS + s]);
        sum += p[i];
        if (noise_eps > 0.0f) {
          nz[i] = rz_mz_gamma(noise_alpha, seed, (uint32_t)(t.global_offset + g), move_id, (uint32_t)s);
          nsum += nz[i];
        }
      }
    }
  }
  sum = rz_warp_sum_f32(sum);
  nsum = rz_warp_sum_f32(nsum);
  const float inv = sum > 0.0f ? 1.0f / sum : 0.0f, ninv = nsum > 0.0f ? 1.0f / nsum : 0.0f;
#pragma unroll
  for (int i = 0; i < RZ_MZ_MAX_ITERS; ++i) {
    const int s = lane + 32 * i;
    if (s < AS) {
      const bool ok = p[i] >= 0.0f;
      float pr = ok ? p[i] * inv : 0.0f;                         // p / policy_sum over the legal actions
      if (ok && noise_eps > 0.0f) pr = pr * (1.0f - noise_eps) + nz[i] * ninv * noise_eps;
      t.edge_N[nb + s] = ok ? 0 : -1;
      t.edge_W[nb + s] = 0.0;
      t.edge_P[nb + s] = pr;
      t.edge_child[nb + s] = -1;
    }
  }
  if (lane == 0) {
    t.n_nodes[g] = 1;
    t.root_N[g] = 0;
    t.root_W[g] = 0.0;
    t.mm_min[g] = t.known_min;       // MinMaxStats(known_bounds): +inf / -inf when unknown
    t.mm_max[g] = t.known_max;
    t.depth[g] = 0;
  }
}

// ---------------------------------------------------------------------------
// select_child down to the first unexpanded child (run_mcts inner loop)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(RZ_MZ_THREADS) rz_mz_select_kernel(rz_mz_desc t) {
  rz::grid_dep_wait();     // programmatic dependent launch: see rz_common.cuh
  rz::grid_dep_launch();
  const int g = blockIdx.x * RZ_MZ_WARPS + (threadIdx.x >> 5);
  if (g >= t.n_trees) return;
  const int lane = rz_lane(), AS = t.action_stride, iters = AS >> 5;
  int32_t* pnode = t.path_node + (size_t)g * t.max_depth;
  int32_t* pact = t.path_action + (size_t)g * t.max_depth;
  const double mn = t.mm_min[g], mx = t.mm_max[g];
  int node = 0, depth = 0, Np = t.root_N[g], fault = 0;
  for (;;) {
    if (depth >= t.max_depth) { fault = 1; break; }
    const size_t base = rz_mz_base(t, g, node);
    int NpC = Np;
    if (NpC >= t.pbc_table_len) { fault = 1; NpC = t.pbc_table_len - 1; }
    const double pbc0 = t.pbc_table[NpC];                        // log((Np + base + 1) / base) + init
    const double sq = __dsqrt_rn((double)Np);
    double best_s = 0.0;
    int best_slot = -1, best_n = 0;
#pragma unroll
    for (int i = 0; i < RZ_MZ_MAX_ITERS; ++i) {
      if (i >= iters) break;
      const int slot = lane + 32 * i;
      const int n = t.edge_N[base + slot];
      if (n < 0) continue;
      const double pbc = __dmul_rn(pbc0, __ddiv_rn(sq, (double)(n + 1)));
      double s = __dmul_rn(pbc, (double)t.edge_P[base + slot]);
      if (n > 0) {
        // child.reward + discount * -child.value()   (reward 0; two-player sign)
        const double q = __dmul_rn(t.discount, -__ddiv_rn(t.edge_W[base + slot], (double)n));
        s = __dadd_rn(s, rz_mz_normalize(q, mn, mx));
      }
      // max((score, action, child)): ties go to the HIGHEST action
      if (best_slot < 0 || s > best_s || (s == best_s && slot > best_slot)) { best_s = s; best_slot = slot; best_n = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double s2 = __shfl_xor_sync(RZ_FULL, best_s, o);
      const int slot2 = __shfl_xor_sync(RZ_FULL, best_slot, o);
      const int n2 = __shfl_xor_sync(RZ_FULL, best_n, o);
      if (slot2 >= 0 && (best_slot < 0 || s2 > best_s || (s2 == best_s && slot2 > best_slot))) {
        best_s = s2; best_slot = slot2; best_n = n2;
      }
    }
    if (best_slot < 0) { fault = 1; break; }
    if (lane == 0) { pnode[depth] = node; pact[depth] = best_slot; }
    depth += 1;
    const int child = t.edge_child[base + best_slot];
    if (child < 0) break;                                          // not node.expanded()
    node = child;
    Np = best_n;
  }
  if (lane == 0) {
    t.depth[g] = fault ? -1 : depth;
    t.leaf_parent[g] = fault ? 0 : pnode[depth - 1];
    t.leaf_action[g] = fault ? 0 : pact[depth - 1];
    if (fault) t.fault[g] |= 1;
  }
}

// ---------------------------------------------------------------------------
// expand_node(leaf, recurrent_inference) + backpropagate
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(RZ_MZ_THREADS)
rz_mz_expand_backup_kernel(rz_mz_desc t, const float* __restrict__ logp, const float* __restrict__ value) {
  rz::grid_dep_wait();     // programmatic dependent launch: see rz_common.cuh
  rz::grid_dep_launch();
  const int g = blockIdx.x * RZ_MZ_WARPS + (threadIdx.x >> 5);
  if (g >= t.n_trees) return;
  const int depth = t.depth[g];
  if (depth <= 0) return;
  const int lane = rz_lane(), A = t.n_actions, AS = t.action_stride;
  const int32_t* pnode = t.path_node + (size_t)g * t.max_depth;
  const int32_t* pact = t.path_action + (size_t)g * t.max_depth;
  const int nn = t.n_nodes[g];
  if (nn >= t.max_nodes) { if (lane == 0) t.fault[g] |= 2; return; }
  const size_t nb = rz_mz_base(t, g, nn);
  for (int s = lane; s < AS; s += 32) {
    const bool ok = s < A;                                        // the whole action space below the root
    t.edge_N[nb + s] = ok ? 0 : -1;
    t.edge_W[nb + s] = 0.0;
    t.edge_P[nb + s] = ok ? expf(logp[(size_t)g * AS + s]) : 0.0f;
    t.edge_child[nb + s] = -1;
  }
  if (lane == 0) {
    t.edge_child[rz_mz_base(t, g, pnode[depth - 1]) + pact[depth - 1]] = nn;
    t.n_nodes[g] = nn + 1;
    // backpropagate: the leaf adds +value (its own to_play), its parent -value, ...
    double val = (double)value[g];
    double mn = t.mm_min[g], mx = t.mm_max[g];
    for (int d = depth; d >= 0; --d) {
      const double x = ((depth - d) & 1) ? -val : val;
      double w;
      int n;
      if (d == 0) {
        w = __dadd_rn(t.root_W[g], x); n = t.root_N[g] + 1;
        t.root_W[g] = w; t.root_N[g] = n;
      } else {
        const size_t e = rz_mz_base(t, g, pnode[d - 1]) + pact[d - 1];
        w = __dadd_rn(t.edge_W[e], x); n = t.edge_N[e] + 1;
        t.edge_W[e] = w; t.edge_N[e] = n;
      }
      const double u = __dmul_rn(t.discount, -__ddiv_rn(w, (double)n));   // reward + discount * -value()
      mn = fmin(mn, u); mx = fmax(mx, u);
      val = __dmul_rn(t.discount, val);                                   // reward (0) + discount * value
    }
    t.mm_min[g] = mn; t.mm_max[g] = mx;
  }
}

// ---------------------------------------------------------------------------
// dynamics input: stage[g] = pool[parent[g]][g] with channel 127 := one-hot plane of action[g]
// pool / stage rows are bf16 [.][128]; a board owns P = S*S rows (row = y*S + x).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rz_mz_gather_kernel(const uint4* __restrict__ pool, const int32_t* __restrict__ parent,
                    const int32_t* __restrict__ action, uint4* __restrict__ stage, int n_trees, int P, int W,
                    int S, int cells, size_t slot_stride_u4) {
  const int g = blockIdx.y;
  const int par = parent[g];
  const int a = action[g];
  const int arow = a < cells ? (a / W) * S + (a % W) : -1;        // a >= cells (Go's pass): empty plane
  const uint4* src = pool + (size_t)par * slot_stride_u4 + (size_t)g * P * 16;
  uint4* dst = stage + (size_t)g * P * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P * 16; i += gridDim.x * blockDim.x) {
    uint4 v = src[i];
    if ((i & 15) == 15) {                                          // channels 120..127 of row i >> 4
      const uint32_t one = (i >> 4) == arow ? 0x3F80u : 0u;
      v.w = (v.w & 0x0000ffffu) | (one << 16);
    }
    dst[i] = v;
  }
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
static int rz_check_mz(const rz_mz_desc* t, const char* who) {
  RZ_REQUIRE(t, "%s: null desc", who);
  RZ_REQUIRE(t->n_trees >= 0 && t->n_actions >= 1 && t->max_nodes >= 2 && t->max_depth >= 1, "%s: sizes", who);
  RZ_REQUIRE(t->action_stride >= t->n_actions && !(t->action_stride & 31) &&
                 t->action_stride <= 32 * RZ_MZ_MAX_ITERS, "%s: action_stride %d", who, t->action_stride);
  RZ_REQUIRE(t->edge_N && t->edge_W && t->edge_P && t->edge_child && t->n_nodes && t->root_N && t->root_W &&
                 t->mm_min && t->mm_max && t->path_node && t->path_action && t->depth && t->leaf_parent &&
                 t->leaf_action && t->fault && t->pbc_table && t->pbc_table_len >= 2, "%s: null array", who);
  return 0;
}
static inline dim3 rz_mz_grid(int n) { return dim3((unsigned)((n + RZ_MZ_WARPS - 1) / RZ_MZ_WARPS)); }

extern "C" int rz_sizeof_mz_desc(void) { return (int)sizeof(rz_mz_desc); }

extern "C" int rz_mz_root(const rz_mz_desc* t, const float* logp, const uint8_t* legal, float noise_eps,
                          float noise_alpha, unsigned long long seed, unsigned int move_id,
                          const int32_t* move_ids, void* stream) {
  if (rz_check_mz(t, "rz_mz_root")) return -1;
  RZ_REQUIRE(logp, "rz_mz_root: null logp");
  RZ_REQUIRE(noise_eps >= 0.0f && noise_eps <= 1.0f && (noise_eps == 0.0f || noise_alpha > 0.0f), "rz_mz_root: noise");
  if (t->n_trees == 0) return 0;
  rz_mz_root_kernel<<<rz_mz_grid(t->n_trees), RZ_MZ_THREADS, 0, (cudaStream_t)stream>>>(*t, logp, legal, noise_eps,
                                                                                         noise_alpha, seed, move_id,
                                                                                         move_ids);
  RZ_LAUNCH_CHECK("rz_mz_root");
  return 0;
}

extern "C" int rz_mz_select(const rz_mz_desc* t, void* stream) {
  if (rz_check_mz(t, "rz_mz_select")) return -1;
  if (t->n_trees == 0) return 0;
  rz_launch_pdl(rz_mz_select_kernel, rz_mz_grid(t->n_trees), RZ_MZ_THREADS, 0, (cudaStream_t)stream, *t);
  RZ_LAUNCH_CHECK("rz_mz_select");
  return 0;
}

extern "C" int rz_mz_expand_backup(const rz_mz_desc* t, const float* logp, const float* value, void* stream) {
  if (rz_check_mz(t, "rz_mz_expand_backup")) return -1;
  RZ_REQUIRE(logp && value, "rz_mz_expand_backup: null network output");
  if (t->n_trees == 0) return 0;
  rz_launch_pdl(rz_mz_expand_backup_kernel, rz_mz_grid(t->n_trees), RZ_MZ_THREADS, 0, (cudaStream_t)stream, *t, logp, value);
  RZ_LAUNCH_CHECK("rz_mz_expand_backup");
  return 0;
}

extern "C" int rz_mz_gather(const void* pool, const int32_t* parent, const int32_t* action, void* stage,
                            int n_trees, int board_rows, int board_cols, int row_stride, long long slot_rows,
                            void* stream) {
  RZ_REQUIRE(pool && parent && action && stage, "rz_mz_gather: null argument");
  RZ_REQUIRE(n_trees >= 0 && board_rows >= 1 && board_cols >= 1 && row_stride >= board_cols &&
                 row_stride >= board_rows, "rz_mz_gather: geometry");
  RZ_REQUIRE(slot_rows >= (long long)n_trees * row_stride * row_stride, "rz_mz_gather: slot_rows %lld", slot_rows);
  if (n_trees == 0) return 0;
  const int P = row_stride * row_stride;
  dim3 grid((unsigned)((P * 16 + 255) / 256), (unsigned)n_trees);
  rz_mz_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)pool, parent, action, (uint4*)stage, n_trees, P, board_cols, row_stride, board_rows * board_cols,
      (size_t)slot_rows * 16);
  RZ_LAUNCH_CHECK("rz_mz_gather");
  return 0;
}
