// rlzero_b200 -- training-side data kernels that consume the self-play trajectory records
// without leaving HBM.
//
// Reference: TrainPipeline.get_equi_data (tools/train_alphazero.py:59-79): every recorded ply
// (state planes [4][H][W], pi [H*W], z) is expanded into its 8 board symmetries, in this order:
//   for i in 1..4:  rot90(state, i), then fliplr of that
// The policy vector goes through flipud o rot90^i o flipud (and fliplr in between for the flipped
// copies) -- literally what the reference does, including the fact that this is NOT the same
// geometric map as the one applied to the planes (the reference kept upstream's flipud although
// its current_state is not flipped).  Parity is bit-for-bit data movement: no arithmetic.
#include "rz_common.cuh"

namespace {

// source coordinates of rot90 applied `i` times: out[r][c] = in[R1^i(r, c)], R1(r, c) = (c, N-1-r)
__device__ __forceinline__ void rot_src(int i, int N, int& r, int& c) {
  for (int t = 0; t < i; ++t) {
    const int nr = c, nc = N - 1 - r;
    r = nr; c = nc;
  }
}

__global__ void __launch_bounds__(256)
rz_augment_equi_kernel(rz_game_desc gd, const uint32_t* __restrict__ rows, const int32_t* __restrict__ info,
                       int info_stride, const float* __restrict__ pi, float* __restrict__ out_planes,
                       float* __restrict__ out_pi, float* __restrict__ out_z, int n) {
  const int H = gd.board_size, A = gd.n_actions, AS = gd.action_stride;
  const long long total = (long long)n * 8 * A;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(idx % A);
    const long long rec8 = idx / A;
    const int sym = (int)(rec8 & 7);          // 2*(i-1) + flipped
    const int rec = (int)(rec8 >> 3);
    const int i = (sym >> 1) + 1, flipped = sym & 1;
    const int r = cell / H, c = cell - r * H;
    // planes: out[r][c] = state[R1^i(r, flipped ? N-1-c : c)]
    int sr = r, sc = flipped ? H - 1 - c : c;
    rot_src(i, H, sr, sc);
    const int32_t* inf = info + (size_t)rec * info_stride;
    const int mover = inf[0], last = inf[1];
    const uint32_t* rw = rows + (size_t)rec * 2 * H;
    const uint32_t mine = rw[mover * H + sr], theirs = rw[(mover ^ 1) * H + sr];
    int stones = 0;
    for (int y = 0; y < H; ++y) stones += __popc(rw[y]) + __popc(rw[H + y]);
    float* op = out_planes + ((size_t)rec8 * 4) * A + cell;
    op[0] = (float)((mine >> sc) & 1u);                                   // gomoku_env.py:100-106
    op[A] = (float)((theirs >> sc) & 1u);
    op[2 * A] = (stones > 0 && last == sr * H + sc) ? 1.0f : 0.0f;        // :108
    op[3 * A] = (stones & 1) ? 0.0f : 1.0f;                               // :110-111
    // pi: out[r][c] = P[N-1-a][b], (a, b) = R1^i(N-1-r, flipped ? N-1-c : c)
    int a = H - 1 - r, b = flipped ? H - 1 - c : c;
    rot_src(i, H, a, b);
    out_pi[(size_t)rec8 * A + cell] = pi[(size_t)rec * AS + (H - 1 - a) * H + b];
    if (cell == 0) out_z[rec8] = (float)inf[2];
  }
}

// gather rows of a [cap][width] float table: dst[i] = src[index[i]]  (replay-buffer mini-batches)
__global__ void __launch_bounds__(256)
rz_gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ index, float* __restrict__ dst,
                      int n, int width) {
  const long long total = (long long)n * width;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / width), col = (int)(idx - (long long)row * width);
    dst[idx] = src[index[row] * width + col];
  }
}

}  // namespace

extern "C" int rz_augment_equi(const rz_game_desc* g, const uint32_t* rows, const int32_t* info, int info_stride,
                               const float* pi, float* out_planes, float* out_pi, float* out_z, int n,
                               void* stream) {
  if (rz_check_game(g)) return -1;
  RZ_REQUIRE(rows && info && pi && out_planes && out_pi && out_z, "rz_augment_equi: null argument");
  RZ_REQUIRE(info_stride >= 3 && n >= 0, "rz_augment_equi: bad sizes");
  if (n == 0) return 0;
  const long long total = (long long)n * 8 * g->n_actions;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rz_augment_equi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*g, rows, info, info_stride, pi, out_planes,
                                                                  out_pi, out_z, n);
  RZ_LAUNCH_CHECK("rz_augment_equi");
  return 0;
}

extern "C" int rz_gather_rows(const float* src, const long long* index, float* dst, int n, int width,
                              void* stream) {
  RZ_REQUIRE(src && index && dst && n >= 0 && width >= 1, "rz_gather_rows: bad arguments");
  if (n == 0) return 0;
  const long long total = (long long)n * width;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rz_gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, index, dst, n, width);
  RZ_LAUNCH_CHECK("rz_gather_rows");
  return 0;
}
