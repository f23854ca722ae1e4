// nnk_delta.cu -- delta feature computation (the inverse-side stencil of MLPG), batched.
//
// Replaces preprocessing.delta_features (nnmnkwii/preprocessing/generic.py:229-288):
//   for each window:  y[:, w*D + d] = np.correlate(x[:, d], window, mode="same")
// np.correlate(..., "same") centres the window at index len(window) // 2 and treats frames outside
// [0, T) as zero; it ignores the (l, u) of a bandmat-style triple and uses only the coefficients.
// One thread per (frame, static dim), lanes along d (coalesced rows), every utterance of a flat
// (sum_T, D) batch in one launch (deltas never cross an utterance boundary).  Arithmetic in float64
// (the reference's float64 window coefficients promote the correlation to float64), rounded to the
// dtype of x on store.  HBM bound: D values in, nw * D values out per frame.
#include "nnk_common.cuh"

namespace nnk {

struct DeltaParams {
  const void* x;
  void* out;
  int64_t x_ld, out_ld;
  const int64_t* utt_off;
  const int32_t* utt_len;
  int D;
  nnk_windows_t win;
};

template <typename T>
__global__ void __launch_bounds__(256) delta_kernel(const __grid_constant__ DeltaParams p) {
  const int utt = blockIdx.z;
  const int64_t row0 = p.utt_off[utt];
  const int Tn = p.utt_len ? p.utt_len[utt] : (int)(p.utt_off[utt + 1] - row0);
  const int t = blockIdx.x * 8 + threadIdx.y;
  const int d = blockIdx.y * 32 + threadIdx.x;
  if (t >= Tn || d >= p.D) return;
  const T* x = reinterpret_cast<const T*>(p.x) + row0 * p.x_ld + d;
  T* o = reinterpret_cast<T*>(p.out) + (row0 + t) * p.out_ld + d;
#pragma unroll 1
  for (int w = 0; w < p.win.nw; ++w) {
    const int M = p.win.l[w] + p.win.u[w] + 1;
    const int c = M >> 1;
    double acc = 0.0;
    for (int m = 0; m < M; ++m) {
      const int s = t + m - c;
      if (s >= 0 && s < Tn) acc = __dadd_rn(acc, __dmul_rn(p.win.coef[w][m], (double)x[(int64_t)s * p.x_ld]));
    }
    o[(int64_t)w * p.D] = (T)acc;
  }
}

}  // namespace nnk

using namespace nnk;

extern "C" int nnk_delta_features(const void* x, int32_t dtype, int32_t D, int64_t x_ld, const int64_t* utt_off,
                                  const int32_t* utt_len, int32_t n_utt, int32_t max_T, const nnk_windows_t* win,
                                  void* out, int64_t out_ld, void* stream) {
  NNK_REQUIRE(x && out && utt_off && win, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(dtype == NNK_F32 || dtype == NNK_F64, NNK_ERR_ARG, "bad dtype");
  NNK_REQUIRE(win->nw >= 1 && win->nw <= NNK_MAX_WIN, NNK_ERR_UNSUPPORTED, "unsupported number of windows");
  NNK_REQUIRE(D >= 0 && n_utt >= 0 && max_T >= 0, NNK_ERR_ARG, "bad size");
  if (D == 0 || n_utt == 0 || max_T == 0) return NNK_OK;
  NNK_REQUIRE(n_utt <= 65535 && (D + 31) / 32 <= 65535, NNK_ERR_ARG, "too many utterances / dims for one launch");
  DeviceGuard guard(x);
  DeltaParams p;
  p.x = x; p.out = out; p.x_ld = x_ld; p.out_ld = out_ld; p.utt_off = utt_off; p.utt_len = utt_len; p.D = D; p.win = *win;
  dim3 grid((unsigned)((max_T + 7) / 8), (unsigned)((D + 31) / 32), (unsigned)n_utt), block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NNK_F32) delta_kernel<float><<<grid, block, 0, st>>>(p);
  else delta_kernel<double><<<grid, block, 0, st>>>(p);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}
