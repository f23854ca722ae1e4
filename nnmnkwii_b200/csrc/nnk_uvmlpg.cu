// nnk_uvmlpg.cu -- UnitVarianceMLPG forward / backward as a banded stencil sweep (sm_100a).
//
// Replaces autograd.UnitVarianceMLPG (autograd/_impl/mlpg.py:107-172):
//     forward   out  = torch.matmul(R, reshaped_means)        R: (T, nw*T) dense float32
//     backward  grad = torch.matmul(R.transpose(0, 1), grad_output)
// R = (sum_w Wt_w^T W_w)^-1 [Wt_0^T .. Wt_{nw-1}^T] is dense in memory but numerically banded around
// the diagonal of each T x T window block (it decays geometrically; 95 % of the reference's GEMM
// multiplies zeros).  The band is extracted ONCE per R into two small tables
//     Rb [t][w][j] = R[t, w*T + t + j - K]        (forward : y[t]      = sum_w sum_j Rb  * mu_w[t+j-K])
//     RbT[s][w][j] = R[s + j - K, w*T + s]        (backward: g_w[s]    = sum_j     RbT * o[s+j-K])
// with the half-width K chosen on the host from the measured decay profile so that everything
// dropped is below a tolerance relative to max|R| (K = T-1 keeps R exactly: always correct, never a
// fallback).  The sweep reads every input element once from HBM, keeps the tables in L1/L2 and is
// HBM/FP32-FMA bound; no tensor cores (a 40-tap banded FIR is not a dense contraction).
#include "nnk_common.cuh"

namespace nnk {

// profile[dist] = max_{w, |t-s| = dist} |R[t, w*T + s]|   (non-negative floats order like uints)
template <typename T>
__global__ void uv_profile_kernel(const T* __restrict__ R, int Tn, int nw, float* __restrict__ profile) {
  const int t = blockIdx.x;
  const int64_t ld = (int64_t)nw * Tn;
  for (int c = threadIdx.x; c < nw * Tn; c += blockDim.x) {
    const int s = c % Tn;
    const float v = fabsf((float)R[(int64_t)t * ld + c]);
    const int dist = s > t ? s - t : t - s;
    if (v > 0.f) atomicMax(reinterpret_cast<unsigned int*>(profile) + dist, __float_as_uint(v));
  }
}

template <typename T>
__global__ void uv_extract_kernel(const T* __restrict__ R, int Tn, int nw, int K, T* __restrict__ Rb, T* __restrict__ RbT) {
  const int t = blockIdx.x;
  const int W = 2 * K + 1;
  const int64_t ld = (int64_t)nw * Tn;
  for (int c = threadIdx.x; c < nw * W; c += blockDim.x) {
    const int w = c / W, j = c % W;
    const int s = t + j - K;
    const bool ok = (s >= 0 && s < Tn);
    Rb[((int64_t)t * nw + w) * W + j] = ok ? R[(int64_t)t * ld + (int64_t)w * Tn + s] : T(0);
    RbT[((int64_t)t * nw + w) * W + j] = ok ? R[(int64_t)s * ld + (int64_t)w * Tn + t] : T(0);
  }
}

// x element (b, s, w, d):  natural layout  (B, T, nw*sd): x[b][s][w*sd + d]
//                          reshaped layout (B, nw*T, sd): x[b][w*T + s][d]   (mlpg.py:124-136)
template <typename T, int BB>
__global__ void __launch_bounds__(128) uv_fwd_kernel(const T* __restrict__ Rb, const T* __restrict__ x,
                                                     T* __restrict__ y, int B, int Tn, int sd, int nw, int K,
                                                     int reshaped) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * 4 + warp;
  const int d = blockIdx.y * 32 + lane;
  const int b0 = blockIdx.z * BB;
  if (t >= Tn) return;
  const bool on = d < sd;
  const int W = 2 * K + 1;
  const int j_lo = max(0, K - t), j_hi = min(2 * K, Tn - 1 - t + K);
  const int64_t bstride = (int64_t)Tn * nw * sd;
  T acc[BB];
#pragma unroll
  for (int i = 0; i < BB; ++i) acc[i] = T(0);
  for (int w = 0; w < nw; ++w) {
    const T* coef = Rb + ((int64_t)t * nw + w) * W;
    for (int j = j_lo; j <= j_hi; ++j) {
      const T c = __ldg(coef + j);
      const int s = t + j - K;
      const int64_t off = reshaped ? ((int64_t)w * Tn + s) * sd + d : ((int64_t)s * nw + w) * sd + d;
#pragma unroll
      for (int i = 0; i < BB; ++i)
        if (on && b0 + i < B) acc[i] = fma(c, x[(int64_t)(b0 + i) * bstride + off], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < BB; ++i)
    if (on && b0 + i < B) y[((int64_t)(b0 + i) * Tn + t) * sd + d] = acc[i];
}

template <typename T, int BB>
__global__ void __launch_bounds__(128) uv_bwd_kernel(const T* __restrict__ RbT, const T* __restrict__ go,
                                                     T* __restrict__ gx, int B, int Tn, int sd, int nw, int K,
                                                     int reshaped) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * 4 + warp;
  const int d = blockIdx.y * 32 + lane;
  const int b0 = blockIdx.z * BB;
  if (s >= Tn) return;
  const bool on = d < sd;
  const int W = 2 * K + 1;
  const int j_lo = max(0, K - s), j_hi = min(2 * K, Tn - 1 - s + K);
  const int64_t bstride = (int64_t)Tn * nw * sd;
  for (int w = 0; w < nw; ++w) {
    T acc[BB];
#pragma unroll
    for (int i = 0; i < BB; ++i) acc[i] = T(0);
    const T* coef = RbT + ((int64_t)s * nw + w) * W;
    for (int j = j_lo; j <= j_hi; ++j) {
      const T c = __ldg(coef + j);
      const int t = s + j - K;
#pragma unroll
      for (int i = 0; i < BB; ++i)
        if (on && b0 + i < B) acc[i] = fma(c, go[((int64_t)(b0 + i) * Tn + t) * sd + d], acc[i]);
    }
    const int64_t off = reshaped ? ((int64_t)w * Tn + s) * sd + d : ((int64_t)s * nw + w) * sd + d;
#pragma unroll
    for (int i = 0; i < BB; ++i)
      if (on && b0 + i < B) gx[(int64_t)(b0 + i) * bstride + off] = acc[i];
  }
}

template <typename T>
static int uv_apply(const void* tab, const void* x, void* y, int B, int Tn, int sd, int nw, int K, int backward,
                    int reshaped, cudaStream_t st) {
  constexpr int BB = 4;
  dim3 grid((Tn + 3) / 4, (sd + 31) / 32, (B + BB - 1) / BB);
  if (grid.y > 65535 || grid.z > 65535) { set_error("static_dim or batch too large for one launch"); return NNK_ERR_ARG; }
  if (backward)
    uv_bwd_kernel<T, BB><<<grid, 128, 0, st>>>((const T*)tab, (const T*)x, (T*)y, B, Tn, sd, nw, K, reshaped);
  else
    uv_fwd_kernel<T, BB><<<grid, 128, 0, st>>>((const T*)tab, (const T*)x, (T*)y, B, Tn, sd, nw, K, reshaped);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

}  // namespace nnk

using namespace nnk;

extern "C" int nnk_uv_band_profile(const void* R, int32_t dtype, int32_t T, int32_t nw, float* profile, void* stream) {
  NNK_REQUIRE(R && profile && T > 0 && nw > 0, NNK_ERR_ARG, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  NNK_CUDA_CHECK(cudaMemsetAsync(profile, 0, sizeof(float) * (size_t)T, st));
  if (dtype == NNK_F32) uv_profile_kernel<float><<<T, 256, 0, st>>>((const float*)R, T, nw, profile);
  else uv_profile_kernel<double><<<T, 256, 0, st>>>((const double*)R, T, nw, profile);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

extern "C" int nnk_uv_band_extract(const void* R, int32_t dtype, int32_t T, int32_t nw, int32_t K, void* Rb, void* RbT,
                                   void* stream) {
  NNK_REQUIRE(R && Rb && RbT && T > 0 && nw > 0 && K >= 0, NNK_ERR_ARG, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NNK_F32) uv_extract_kernel<float><<<T, 128, 0, st>>>((const float*)R, T, nw, K, (float*)Rb, (float*)RbT);
  else uv_extract_kernel<double><<<T, 128, 0, st>>>((const double*)R, T, nw, K, (double*)Rb, (double*)RbT);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

extern "C" int nnk_uv_apply(const void* table, const void* x, void* y, int32_t dtype, int32_t B, int32_t T, int32_t sd,
                            int32_t nw, int32_t K, int32_t backward, int32_t reshaped, void* stream) {
  NNK_REQUIRE(table && x && y, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(B >= 0 && T > 0 && sd >= 0 && nw > 0 && K >= 0, NNK_ERR_ARG, "bad size");
  if (B == 0 || sd == 0) return NNK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == NNK_F32 ? uv_apply<float>(table, x, y, B, T, sd, nw, K, backward, reshaped, st)
                          : uv_apply<double>(table, x, y, B, T, sd, nw, K, backward, reshaped, st);
}
