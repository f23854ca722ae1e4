// nnk_gmm.cu -- frame-wise GMM mapping in front of MLPG on sm_100a, float64.
//
// Replaces, for all frames of an utterance (or a batch of utterances) at once, the per-frame Python of
//   baseline.gmm.MLPGBase._transform_frame   baseline/gmm.py:97-121   (Eq. 9, 11, 13: posterior mean)
//   baseline.gmm.MLPG.transform              baseline/gmm.py:207-247  (Eq. 37: arg-max mixture, Eq. 22 means
//                                                                       E, Eq. 23 diagonal variances D)
// The reference calls sklearn's predict_proba and one np.linalg.solve per frame and mixture; here
//   gmm_logprob_kernel  : lp[t][m] = log w_m + log N(x_t | mu_m, Sxx_m) = c_m - 1/2 |(x_t - mu_m) U_m|^2
//                         (U_m = sklearn's precisions_cholesky_, c_m folded on the host);
//   gmm_select_kernel   : m*(t) = first arg-max of lp[t][:], E[t] = nu_{m*} + A_{m*} (x_t - mu_{m*}),
//                         D[t] = Dm[m*]  -- written straight in the (T, D) layout paramgen.mlpg consumes;
//   gmm_posterior_kernel: E[t] = sum_m softmax(lp[t])_m (nu_m + A_m (x_t - mu_m)).
// A_m = Syx_m Sxx_m^-1 is formed once on the host (the reference re-solves per frame).  A block owns a tile
// of frames; the D x D matrix of the current mixture is staged in shared memory once per tile, lanes run
// along the output dimension (conflict-free shared-memory rows, coalesced global rows), every thread keeps
// FPW frames x EPL outputs in registers.  Narrow float64 matrix-vector work: FP64 pipe, no tensor cores.
#include <math_constants.h>

#include "nnk_common.cuh"

namespace nnk {

constexpr int GMM_FT = 16;   // frames per block
constexpr int GMM_FPW = 4;   // frames per warp (4 warps)
constexpr int GMM_EPL = 3;   // output dims per lane: D <= 96

struct GmmParams {
  nnk_gmm_t g;
  const double* x;
  int64_t x_ld;
  int T;
  double* lp;      // (T, M)
  double* E;       // (T, D)
  double* Dv;      // (T, D) or NULL
  int32_t* mix;    // (T) or NULL
};

// stage a D x D row-major matrix and the tile's (x - mu) rows in shared memory
__device__ __forceinline__ void gmm_stage(const double* __restrict__ mat, const double* __restrict__ mu, const GmmParams& p, int t0,
                                          double* sm_mat, double* sm_diff) {
  const int D = p.g.D;
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) sm_mat[e] = mat[e];
  for (int e = threadIdx.x; e < GMM_FT * D; e += blockDim.x) {
    const int f = e / D, d = e - f * D;
    const int t = t0 + f;
    sm_diff[e] = (t < p.T) ? p.x[(int64_t)t * p.x_ld + d] - mu[d] : 0.0;
  }
}

// acc[f][k] = sum_d diff[f][d] * mat[d][lane + 32 k]   for the warp's GMM_FPW frames
__device__ __forceinline__ void gmm_matvec(const double* sm_mat, const double* sm_diff, int D, int warp, int lane,
                                           double (&acc)[GMM_FPW][GMM_EPL]) {
#pragma unroll
  for (int f = 0; f < GMM_FPW; ++f)
#pragma unroll
    for (int k = 0; k < GMM_EPL; ++k) acc[f][k] = 0.0;
  const double* df = sm_diff + (size_t)(warp * GMM_FPW) * D;
  for (int d = 0; d < D; ++d) {
    double m[GMM_EPL];
#pragma unroll
    for (int k = 0; k < GMM_EPL; ++k) {
      const int e = lane + 32 * k;
      m[k] = (e < D) ? sm_mat[(size_t)d * D + e] : 0.0;
    }
#pragma unroll
    for (int f = 0; f < GMM_FPW; ++f) {
      const double v = df[(size_t)f * D + d];
#pragma unroll
      for (int k = 0; k < GMM_EPL; ++k) acc[f][k] = fma(v, m[k], acc[f][k]);
    }
  }
}

__global__ void __launch_bounds__(128) gmm_logprob_kernel(const GmmParams p) {
  extern __shared__ __align__(16) double gsm[];
  const int D = p.g.D, m = blockIdx.y, t0 = blockIdx.x * GMM_FT;
  double* sm_mat = gsm;
  double* sm_diff = gsm + (size_t)D * D;
  gmm_stage(p.g.prec_chol + (size_t)m * D * D, p.g.src_means + (size_t)m * D, p, t0, sm_mat, sm_diff);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[GMM_FPW][GMM_EPL];
  gmm_matvec(sm_mat, sm_diff, D, warp, lane, acc);
#pragma unroll
  for (int f = 0; f < GMM_FPW; ++f) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < GMM_EPL; ++k) s = fma(acc[f][k], acc[f][k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int t = t0 + warp * GMM_FPW + f;
    if (lane == 0 && t < p.T) p.lp[(int64_t)t * p.g.M + m] = p.g.log_const[m] - 0.5 * s;
  }
}

// one warp per frame: arg-max mixture (first maximum, like numpy / torch argmax), then the affine map of
// that mixture.  A_t is stored transposed ([m][j][i]) so that lanes read consecutive i.
__global__ void __launch_bounds__(128) gmm_select_kernel(const GmmParams p) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (t >= p.T) return;
  const int D = p.g.D, M = p.g.M;
  double best = -CUDART_INF;
  int arg = 0x7fffffff;
  for (int m = lane; m < M; m += 32) {
    const double v = p.lp[(int64_t)t * M + m];
    if (v > best) { best = v; arg = m; }  // ascending m per lane: keeps the lane's first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (arg == 0x7fffffff) arg = 0;  // all NaN / -inf: numpy's argmax returns 0
  const double* At = p.g.A_t + (size_t)arg * D * D;
  const double* mu = p.g.src_means + (size_t)arg * D;
  double acc[GMM_EPL];
#pragma unroll
  for (int k = 0; k < GMM_EPL; ++k) acc[k] = 0.0;
  for (int j = 0; j < D; ++j) {
    const double dj = p.x[(int64_t)t * p.x_ld + j] - mu[j];
#pragma unroll
    for (int k = 0; k < GMM_EPL; ++k) {
      const int i = lane + 32 * k;
      if (i < D) acc[k] = fma(At[(size_t)j * D + i], dj, acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < GMM_EPL; ++k) {
    const int i = lane + 32 * k;
    if (i < D) {
      p.E[(int64_t)t * D + i] = p.g.tgt_means[(size_t)arg * D + i] + acc[k];
      if (p.Dv) p.Dv[(int64_t)t * D + i] = p.g.Dm[(size_t)arg * D + i];
    }
  }
  if (p.mix && lane == 0) p.mix[t] = arg;
}

// E[t] = sum_m post[t][m] (nu_m + A_m (x_t - mu_m)),  post = softmax(lp[t][:])   (Eq. 9, 11, 13)
__global__ void __launch_bounds__(128) gmm_posterior_kernel(const GmmParams p) {
  extern __shared__ __align__(16) double gsm[];
  const int D = p.g.D, M = p.g.M, t0 = blockIdx.x * GMM_FT;
  double* sm_mat = gsm;
  double* sm_diff = gsm + (size_t)D * D;
  double* sm_max = sm_diff + (size_t)GMM_FT * D;  // [FT] max_m lp, then [FT] 1 / sum exp
  double* sm_inv = sm_max + GMM_FT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int f = warp; f < GMM_FT; f += 4) {  // softmax statistics of the tile's frames
    const int t = t0 + f;
    double mx = -CUDART_INF;
    if (t < p.T)
      for (int m = lane; m < M; m += 32) mx = fmax(mx, p.lp[(int64_t)t * M + m]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double s = 0.0;
    if (t < p.T)
      for (int m = lane; m < M; m += 32) s += exp(p.lp[(int64_t)t * M + m] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { sm_max[f] = mx; sm_inv[f] = (t < p.T) ? 1.0 / s : 0.0; }
  }
  double out[GMM_FPW][GMM_EPL];
#pragma unroll
  for (int f = 0; f < GMM_FPW; ++f)
#pragma unroll
    for (int k = 0; k < GMM_EPL; ++k) out[f][k] = 0.0;
  for (int m = 0; m < M; ++m) {
    __syncthreads();
    gmm_stage(p.g.A_t + (size_t)m * D * D, p.g.src_means + (size_t)m * D, p, t0, sm_mat, sm_diff);
    __syncthreads();
    double acc[GMM_FPW][GMM_EPL];
    gmm_matvec(sm_mat, sm_diff, D, warp, lane, acc);
#pragma unroll
    for (int f = 0; f < GMM_FPW; ++f) {
      const int fi = warp * GMM_FPW + f, t = t0 + fi;
      const double post = (t < p.T) ? exp(p.lp[(int64_t)t * M + m] - sm_max[fi]) * sm_inv[fi] : 0.0;
#pragma unroll
      for (int k = 0; k < GMM_EPL; ++k) {
        const int i = lane + 32 * k;
        if (i < D) out[f][k] = fma(post, p.g.tgt_means[(size_t)m * D + i] + acc[f][k], out[f][k]);
      }
    }
  }
#pragma unroll
  for (int f = 0; f < GMM_FPW; ++f) {
    const int t = t0 + warp * GMM_FPW + f;
#pragma unroll
    for (int k = 0; k < GMM_EPL; ++k) {
      const int i = lane + 32 * k;
      if (t < p.T && i < D) p.E[(int64_t)t * D + i] = out[f][k];
    }
  }
}

}  // namespace nnk

using namespace nnk;

static int gmm_check(const nnk_gmm_t* g, const double* x, int32_t T) {
  NNK_REQUIRE(g && x, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(g->src_means && g->prec_chol && g->log_const && g->tgt_means && g->A_t, NNK_ERR_ARG, "NULL GMM table");
  NNK_REQUIRE(T >= 0 && g->M >= 1 && g->D >= 1, NNK_ERR_ARG, "bad size");
  NNK_REQUIRE(g->D <= 32 * GMM_EPL, NNK_ERR_UNSUPPORTED, "feature dimension > 96 is not supported by the GMM kernels");
  NNK_REQUIRE(g->M <= 65535, NNK_ERR_UNSUPPORTED, "more than 65535 mixtures");
  return NNK_OK;
}

extern "C" int nnk_gmm_logprob(const nnk_gmm_t* g, const double* x, int64_t x_ld, int32_t T, double* lp, void* stream) {
  int rc = gmm_check(g, x, T);
  if (rc) return rc;
  NNK_REQUIRE(lp != nullptr, NNK_ERR_ARG, "NULL output");
  if (T == 0) return NNK_OK;
  DeviceGuard guard(x);
  GmmParams p{};
  p.g = *g; p.x = x; p.x_ld = x_ld; p.T = T; p.lp = lp;
  const size_t smem = sizeof(double) * ((size_t)g->D * g->D + (size_t)GMM_FT * g->D);
  NNK_CUDA_CHECK(cudaFuncSetAttribute(gmm_logprob_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((T + GMM_FT - 1) / GMM_FT), (unsigned)g->M);
  gmm_logprob_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(p);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

extern "C" int nnk_gmm_map(const nnk_gmm_t* g, const double* x, int64_t x_ld, int32_t T, const double* lp, int32_t mode,
                           double* E, double* Dv, int32_t* mix, void* stream) {
  int rc = gmm_check(g, x, T);
  if (rc) return rc;
  NNK_REQUIRE(lp && E, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(mode == 0 || mode == 1, NNK_ERR_ARG, "mode must be 0 (arg-max mixture) or 1 (posterior mean)");
  NNK_REQUIRE(mode == 1 || Dv == nullptr || g->Dm != nullptr, NNK_ERR_ARG, "Dm table missing");
  if (T == 0) return NNK_OK;
  DeviceGuard guard(x);
  GmmParams p{};
  p.g = *g; p.x = x; p.x_ld = x_ld; p.T = T; p.lp = const_cast<double*>(lp); p.E = E; p.Dv = Dv; p.mix = mix;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    gmm_select_kernel<<<(unsigned)((T + 3) / 4), 128, 0, st>>>(p);
  } else {
    const size_t smem = sizeof(double) * ((size_t)g->D * g->D + (size_t)GMM_FT * g->D + 2 * GMM_FT);
    NNK_CUDA_CHECK(cudaFuncSetAttribute(gmm_posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_posterior_kernel<<<(unsigned)((T + GMM_FT - 1) / GMM_FT), 128, smem, st>>>(p);
  }
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}
