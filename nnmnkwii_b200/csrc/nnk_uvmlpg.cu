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
                                                     int reshaped, int r_lo, int r_hi, int skip_lo, int skip_hi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = r_lo + blockIdx.x * 4 + warp;
  if (t >= skip_lo) t += skip_hi - skip_lo;  // rows [skip_lo, skip_hi) belong to the Toeplitz kernel
  if (t >= r_hi) return;
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
                                                     int reshaped, int r_lo, int r_hi, int skip_lo, int skip_hi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int s = r_lo + blockIdx.x * 4 + warp;
  if (s >= skip_lo) s += skip_hi - skip_lo;
  if (s >= r_hi) return;
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

// ---- Toeplitz interior fast path ---------------------------------------------------------------------
// Away from the two ends R is shift-invariant (rows of each window block are copies of one FIR
// filter to float32 rounding): rows t in [t_lo, t_hi) share the coefficient table `h`.  The table
// lives in the kernel parameter (constant bank), every coefficient index is a compile-time
// constant after unrolling, so the inner loop is FFMA with a constant operand -- no coefficient
// loads at all.  One thread produces TTU consecutive frames of one (batch item, static dim) and
// slides over the TTU + 2K input frames it needs, lanes along d (coalesced row segments).
template <int KK, int NWT>
struct UvTaps {
  float h[NWT][2 * KK + 1];
};

template <int KK, int NWT, int TTU>
__global__ void __launch_bounds__(128, 3) uv_fwd_toeplitz_kernel(const __grid_constant__ UvTaps<KK, NWT> taps,
                                                                 const float* __restrict__ x, float* __restrict__ y, int B,
                                                                 int Tn, int sd, int t_lo, int t_hi, int reshaped) {
  const int d = blockIdx.y * 32 + (threadIdx.x & 31);
  const int t0 = t_lo + (blockIdx.x * 4 + (threadIdx.x >> 5)) * TTU;  // warp-uniform
  const int b = blockIdx.z;
  if (t0 >= t_hi || d >= sd) return;
  float acc[TTU];
#pragma unroll
  for (int i = 0; i < TTU; ++i) acc[i] = 0.f;
  const int64_t fstep = reshaped ? (int64_t)sd : (int64_t)NWT * sd;  // elements between consecutive frames
  const int s0 = t0 - KK;
#pragma unroll
  for (int w = 0; w < NWT; ++w) {
    const float* px = x + (int64_t)b * Tn * NWT * sd + (reshaped ? ((int64_t)w * Tn + s0) * sd + d : ((int64_t)s0 * NWT + w) * sd + d);
#pragma unroll
    for (int r = 0; r < TTU + 2 * KK; ++r) {  // frame s0 + r feeds output i with tap j = r - i
      const int s = s0 + r;
      const int sc = min(max(s, 0), Tn - 1);  // frames outside [0, T) contribute zero
      const float ld = __ldg(px + (int64_t)(sc - s0) * fstep);
      const float v = (s == sc) ? ld : 0.f;
#pragma unroll
      for (int i = 0; i < TTU; ++i) {
        const int j = r - i;
        if (j >= 0 && j <= 2 * KK) acc[i] = fmaf(taps.h[w][j], v, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TTU; ++i)
    if (t0 + i < t_hi) y[((int64_t)b * Tn + t0 + i) * sd + d] = acc[i];
}

// backward: gx[b, s, w, d] = sum_j hT[w][j] * go[b, s + j - K, d]; one input stream, NWT outputs
template <int KK, int NWT, int TTU>
__global__ void __launch_bounds__(128, 3) uv_bwd_toeplitz_kernel(const __grid_constant__ UvTaps<KK, NWT> taps,
                                                                 const float* __restrict__ go, float* __restrict__ gx, int B,
                                                                 int Tn, int sd, int t_lo, int t_hi, int reshaped) {
  const int d = blockIdx.y * 32 + (threadIdx.x & 31);
  const int s0 = t_lo + (blockIdx.x * 4 + (threadIdx.x >> 5)) * TTU;
  const int b = blockIdx.z;
  if (s0 >= t_hi || d >= sd) return;
  const int r0 = s0 - KK;
  const float* pg = go + ((int64_t)b * Tn + r0) * sd + d;
  const int64_t bbase = (int64_t)b * Tn * NWT * sd;
  float acc[NWT][TTU];
#pragma unroll
  for (int w = 0; w < NWT; ++w)
#pragma unroll
    for (int i = 0; i < TTU; ++i) acc[w][i] = 0.f;
#pragma unroll
  for (int r = 0; r < TTU + 2 * KK; ++r) {
    const int t = r0 + r;
    const int tc = min(max(t, 0), Tn - 1);
    const float ld = __ldg(pg + (int64_t)(tc - r0) * sd);
    const float v = (t == tc) ? ld : 0.f;
#pragma unroll
    for (int i = 0; i < TTU; ++i) {
      const int j = r - i;
      if (j >= 0 && j <= 2 * KK) {
#pragma unroll
        for (int w = 0; w < NWT; ++w) acc[w][i] = fmaf(taps.h[w][j], v, acc[w][i]);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < NWT; ++w)
#pragma unroll
    for (int i = 0; i < TTU; ++i) {
      const int s = s0 + i;
      if (s < t_hi) {
        const int64_t off = reshaped ? ((int64_t)w * Tn + s) * sd + d : ((int64_t)s * NWT + w) * sd + d;
        gx[bbase + off] = acc[w][i];
      }
    }
}

// ---- packed variants: two adjacent static dims per thread --------------------------------------------
// Blackwell issues FP32 FMAs two at a time from 64-bit register pairs (fma.rn.f32x2, SASS FFMA2):
// a thread owns dims (2*dl, 2*dl+1), loads its frames as float2 and keeps the filter taps of one
// window as (h, h) pairs in registers, so the inner loop is one FFMA2 per two multiply-adds and one
// 8-byte load per two inputs.  Needs an even static_dim and 8-byte aligned tensors.
template <int KK, int NWT, int TTU, bool BWD>
__global__ void __launch_bounds__(128, 3) uv_toeplitz2_kernel(const __grid_constant__ UvTaps<KK, NWT> taps,
                                                              const float* __restrict__ x, float* __restrict__ y, int B,
                                                              int Tn, int sd, int t_lo, int t_hi, int reshaped) {
  const int dl = blockIdx.y * 32 + (threadIdx.x & 31);
  const int d = 2 * dl;
  const int t0 = t_lo + (blockIdx.x * 4 + (threadIdx.x >> 5)) * TTU;  // warp-uniform
  const int b = blockIdx.z;
  if (t0 >= t_hi || d >= sd) return;
  const int s0 = t0 - KK;
  const bool interior = (s0 >= 0) && (t0 + TTU + KK <= Tn);
  const int64_t nwsd = (int64_t)NWT * sd;
  // forward: NWT input streams (window side), one output; backward: one input stream, NWT outputs
  float2 acc[BWD ? NWT : 1][TTU];
#pragma unroll
  for (int w = 0; w < (BWD ? NWT : 1); ++w)
#pragma unroll
    for (int i = 0; i < TTU; ++i) acc[w][i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < NWT; ++w) {
    float2 hp[2 * KK + 1];
#pragma unroll
    for (int j = 0; j <= 2 * KK; ++j) hp[j] = make_float2(taps.h[w][j], taps.h[w][j]);
    if (BWD && w > 0) {
      // the input stream is re-read per window: it is L1 resident, the taps are what fills the registers
    }
    const float* px;
    int64_t fstep;
    if (!BWD) {
      px = x + (int64_t)b * Tn * nwsd + (reshaped ? ((int64_t)w * Tn + s0) * sd + d : ((int64_t)s0 * NWT + w) * sd + d);
      fstep = reshaped ? (int64_t)sd : nwsd;
    } else {
      px = x + ((int64_t)b * Tn + s0) * sd + d;
      fstep = sd;
    }
    float2(&a)[TTU] = acc[BWD ? w : 0];
    if (interior) {
#pragma unroll
      for (int r = 0; r < TTU + 2 * KK; ++r) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(px + r * fstep));
#pragma unroll
        for (int i = 0; i < TTU; ++i) {
          const int j = r - i;
          if (j >= 0 && j <= 2 * KK) a[i] = __ffma2_rn(hp[j], v, a[i]);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < TTU + 2 * KK; ++r) {
        const int s = s0 + r;
        const int sc = min(max(s, 0), Tn - 1);
        float2 v = __ldg(reinterpret_cast<const float2*>(px + (int64_t)(sc - s0) * fstep));
        if (s != sc) v = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < TTU; ++i) {
          const int j = r - i;
          if (j >= 0 && j <= 2 * KK) a[i] = __ffma2_rn(hp[j], v, a[i]);
        }
      }
    }
    if (BWD) {
#pragma unroll
      for (int i = 0; i < TTU; ++i) {
        const int s = t0 + i;
        if (s < t_hi) {
          const int64_t off = reshaped ? ((int64_t)w * Tn + s) * sd + d : ((int64_t)s * NWT + w) * sd + d;
          *reinterpret_cast<float2*>(y + (int64_t)b * Tn * nwsd + off) = a[i];
        }
      }
    }
  }
  if (!BWD) {
#pragma unroll
    for (int i = 0; i < TTU; ++i)
      if (t0 + i < t_hi) *reinterpret_cast<float2*>(y + ((int64_t)b * Tn + t0 + i) * sd + d) = acc[0][i];
  }
}

template <int KK, int NWT>
static int uv_toeplitz_launch(const float* taps_host, const float* x, float* y, int B, int Tn, int sd, int t_lo, int t_hi,
                              int backward, int reshaped, cudaStream_t st) {
  constexpr int TTU = 8;
  UvTaps<KK, NWT> taps;
  for (int w = 0; w < NWT; ++w)
    for (int j = 0; j <= 2 * KK; ++j) taps.h[w][j] = taps_host[w * (2 * KK + 1) + j];
  const int rows = t_hi - t_lo;
  const bool packed = (sd % 2 == 0) && (((uintptr_t)x | (uintptr_t)y) % 8 == 0) && KK <= 32;
  if (packed) {
    dim3 grid2((rows + 4 * TTU - 1) / (4 * TTU), (sd / 2 + 31) / 32, B);
    if (backward) uv_toeplitz2_kernel<KK, NWT, TTU, true><<<grid2, 128, 0, st>>>(taps, x, y, B, Tn, sd, t_lo, t_hi, reshaped);
    else uv_toeplitz2_kernel<KK, NWT, TTU, false><<<grid2, 128, 0, st>>>(taps, x, y, B, Tn, sd, t_lo, t_hi, reshaped);
    count_launch();
    NNK_CUDA_CHECK(cudaGetLastError());
    return NNK_OK;
  }
  dim3 grid((rows + 4 * TTU - 1) / (4 * TTU), (sd + 31) / 32, B);
  if (backward) uv_bwd_toeplitz_kernel<KK, NWT, TTU><<<grid, 128, 0, st>>>(taps, x, y, B, Tn, sd, t_lo, t_hi, reshaped);
  else uv_fwd_toeplitz_kernel<KK, NWT, TTU><<<grid, 128, 0, st>>>(taps, x, y, B, Tn, sd, t_lo, t_hi, reshaped);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

template <int KK, int NWT>
static int uv_toeplitz_launch_pad(const float* padded, const float* x, float* y, int B, int Tn, int sd, int t_lo, int t_hi,
                                  int backward, int reshaped, cudaStream_t st) {
  return uv_toeplitz_launch<KK, NWT>(padded, x, y, B, Tn, sd, t_lo, t_hi, backward, reshaped, st);
}

template <typename T, int BB>
static int uv_apply_bb(const void* tab, const void* x, void* y, int B, int Tn, int sd, int nw, int K, int backward,
                       int reshaped, int skip_lo, int skip_hi, cudaStream_t st) {
  const int rows = Tn - (skip_hi - skip_lo);
  if (rows <= 0) return NNK_OK;
  dim3 grid((rows + 3) / 4, (sd + 31) / 32, (B + BB - 1) / BB);
  if (grid.y > 65535 || grid.z > 65535) { set_error("static_dim or batch too large for one launch"); return NNK_ERR_ARG; }
  if (backward)
    uv_bwd_kernel<T, BB><<<grid, 128, 0, st>>>((const T*)tab, (const T*)x, (T*)y, B, Tn, sd, nw, K, reshaped, 0, Tn, skip_lo, skip_hi);
  else
    uv_fwd_kernel<T, BB><<<grid, 128, 0, st>>>((const T*)tab, (const T*)x, (T*)y, B, Tn, sd, nw, K, reshaped, 0, Tn, skip_lo, skip_hi);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

// few rows (the edges next to a Toeplitz interior): one batch item per thread for more parallelism;
// the whole matrix: four batch items per thread so that a coefficient load feeds four FMAs
template <typename T>
static int uv_apply(const void* tab, const void* x, void* y, int B, int Tn, int sd, int nw, int K, int backward,
                    int reshaped, int skip_lo, int skip_hi, cudaStream_t st) {
  return (skip_hi > skip_lo) ? uv_apply_bb<T, 1>(tab, x, y, B, Tn, sd, nw, K, backward, reshaped, skip_lo, skip_hi, st)
                             : uv_apply_bb<T, 4>(tab, x, y, B, Tn, sd, nw, K, backward, reshaped, skip_lo, skip_hi, st);
}

// ---- factored Toeplitz path ------------------------------------------------------------------------------
// R = P^-1 [Wt_0^T .. Wt_{nw-1}^T]: away from the ends every window block of R is the SAME long filter
// h_0 (a row of P^-1, 2K+1 taps) convolved with a short window stencil c_w (2KC+1 taps, KC <= 2):
//     h_w = h_0 * c_w.
// The host recovers c_w from R alone (least squares on the band rows, residual checked against the
// float32 resolution of R; nnk_uv_apply_toeplitz stays the path when the fit fails), and the sweeps become
//     forward :  b[s] = sum_w sum_k c_w[k] mu_w[s + k - KC]         (short stencils, ~7 FMAs)
//                y[t] = sum_j h_0[j] b[t + j - K]                    (ONE long filter)
//     backward:  a[r] = sum_j hT_0[j] o[r + j - K]                   (ONE long filter)
//                g_w[s] = sum_k cT_w[k] a[s + k - KC]                (short stencils)
// i.e. (2K+1) + ~7 multiply-adds per output instead of nw * (2K+1): 54 instead of 147 at T = 1000.
// One CTA = (batch item, 32 dim PAIRS, time tile).  The intermediate (b or a) of the tile plus its halo
// lives in shared memory [position][lane] as float2; a thread owns two adjacent static dims, all
// multiply-adds are packed FFMA2 whose filter operand is a (h, h) pair in a uniform register
// (constant bank) -- the kernel needs 32-40 registers.  Odd static_dim: the pair (d, d+1) straddles
// the row end for the last lane only; loads and stores of the second element are predicated
// (VEC = false: two 4-byte accesses, the multiply-adds stay packed).
template <int KK, int KC, int NWT>
struct UvFactTaps {
  float2 h0[2 * KK + 1];          // (h, h) pairs of the long filter
  float2 c[NWT][2 * KC + 1];      // (c, c) pairs of the short stencils
};

template <bool VEC>
__device__ __forceinline__ float2 uv_ld2(const float* p, bool ok0, bool ok1) {
  if (VEC) return ok0 ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f);
  float2 v;
  v.x = ok0 ? __ldg(p) : 0.f;
  v.y = ok1 ? __ldg(p + 1) : 0.f;
  return v;
}
template <bool VEC>
__device__ __forceinline__ void uv_st2(float* p, float2 v, bool ok0, bool ok1) {
  if (VEC) {
    if (ok0) *reinterpret_cast<float2*>(p) = v;
  } else {
    if (ok0) p[0] = v.x;
    if (ok1) p[1] = v.y;
  }
}

constexpr int UVF_TTU = 8;  // outputs of the long filter per register block

// rows outside the shift-invariant interval (the ~K rows next to each end) use the per-row band table
// `tab` ([T][nw][2K+1]); they are handled by extra CTAs of the SAME launch (blockIdx.x >= n_tiles), one
// row per warp, so that a sweep is a single kernel.  forward: y[t] = sum_w sum_j tab[t][w][j] x_w[t+j-K];
// backward: g_w[s] = sum_j tab[s][w][j] o[s+j-K].
template <bool BWD, bool VEC>
__device__ __forceinline__ void uv_edge_rows(const float* __restrict__ tab, const float* __restrict__ x, float* __restrict__ y,
                                             int Tn, int sd, int nw, int K, int t_lo, int t_hi, int reshaped, int eblk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = eblk * 4 + warp;          // index among the edge rows: [0, t_lo) then [t_hi, Tn)
  if (t >= t_lo) t += t_hi - t_lo;
  if (t >= Tn) return;
  const int d = 2 * (blockIdx.y * 32 + lane);
  const bool ok0 = d < sd, ok1 = d + 1 < sd;
  const int b = blockIdx.z;
  const int W = 2 * K + 1;
  const int j_lo = max(0, K - t), j_hi = min(2 * K, Tn - 1 - t + K);
  const int64_t nwsd = (int64_t)nw * sd;
  if (!BWD) {
    const float* xb = x + (int64_t)b * Tn * nwsd + d;
    float2 acc = make_float2(0.f, 0.f);
    for (int w = 0; w < nw; ++w) {
      const float* coef = tab + ((int64_t)t * nw + w) * W;
      for (int j = j_lo; j <= j_hi; ++j) {
        const float c = __ldg(coef + j);
        const int s = t + j - K;
        const int64_t off = reshaped ? ((int64_t)w * Tn + s) * sd : (int64_t)s * nwsd + (int64_t)w * sd;
        const float2 v = uv_ld2<VEC>(xb + off, ok0, ok1);
        acc.x = fmaf(c, v.x, acc.x);
        acc.y = fmaf(c, v.y, acc.y);
      }
    }
    uv_st2<VEC>(y + ((int64_t)b * Tn + t) * sd + d, acc, ok0, ok1);
  } else {
    const float* gb = x + (int64_t)b * Tn * sd + d;
    float* ob = y + (int64_t)b * Tn * nwsd + d;
    for (int w = 0; w < nw; ++w) {
      const float* coef = tab + ((int64_t)t * nw + w) * W;
      float2 acc = make_float2(0.f, 0.f);
      for (int j = j_lo; j <= j_hi; ++j) {
        const float c = __ldg(coef + j);
        const float2 v = uv_ld2<VEC>(gb + (int64_t)(t + j - K) * sd, ok0, ok1);
        acc.x = fmaf(c, v.x, acc.x);
        acc.y = fmaf(c, v.y, acc.y);
      }
      const int64_t off = reshaped ? ((int64_t)w * Tn + t) * sd : (int64_t)t * nwsd + (int64_t)w * sd;
      uv_st2<VEC>(ob + off, acc, ok0, ok1);
    }
  }
}

// forward: tile of TC output frames; stage 1 fills sb[p] = b[t0 - KK + p], p < TC + 2 KK
template <int KK, int KC, int NWT, int TC, bool VEC>
__global__ void __launch_bounds__(128, 6) uv_fact_fwd_kernel(const __grid_constant__ UvFactTaps<KK, KC, NWT> taps,
                                                          const float* __restrict__ tab, const float* __restrict__ x,
                                                          float* __restrict__ y, int Tn, int sd, int nw, int K, int t_lo, int t_hi,
                                                          int reshaped, int n_tiles) {
  constexpr int NP = TC + 2 * KK;      // intermediate positions of the tile
  constexpr int PW = NP / 4;           // ... per warp
  constexpr int NBLK = 4;              // stage-1 register blocks per warp (unrolled: the loads of a block overlap
  constexpr int CH = PW / NBLK;        //   the multiply-adds of the previous one)
  static_assert(NP % (4 * NBLK) == 0 && TC % (4 * UVF_TTU) == 0, "tile geometry");
  extern __shared__ __align__(16) float2 sb[];  // [NP][32]
  if ((int)blockIdx.x >= n_tiles) {
    uv_edge_rows<false, VEC>(tab, x, y, Tn, sd, nw, K, t_lo, t_hi, reshaped, (int)blockIdx.x - n_tiles);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = 2 * (blockIdx.y * 32 + lane);
  const bool ok0 = d < sd, ok1 = d + 1 < sd;
  const int b = blockIdx.z;
  const int t0 = t_lo + blockIdx.x * TC;
  const int64_t nwsd = (int64_t)nw * sd;
  const float* xb = x + (int64_t)b * Tn * nwsd + d;
  const int64_t fstep = reshaped ? (int64_t)sd : nwsd;  // elements between consecutive frames of one window
  // ---- stage 1: short stencils -> b ----
#pragma unroll
  for (int half = 0; half < NBLK; ++half) {
    const int p0 = warp * PW + half * CH;
    const int f0 = t0 - KK + p0 - KC;  // first input frame of the block
    const bool interior = (f0 >= 0) && (f0 + CH + 2 * KC <= Tn);
    float2 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < NWT; ++w) {
      if (w < nw) {
        const float* pw = xb + (reshaped ? (int64_t)w * Tn * sd : (int64_t)w * sd);
        float2 v[CH + 2 * KC];
        if (interior) {
#pragma unroll
          for (int r = 0; r < CH + 2 * KC; ++r) v[r] = uv_ld2<VEC>(pw + (int64_t)(f0 + r) * fstep, ok0, ok1);
        } else {
#pragma unroll
          for (int r = 0; r < CH + 2 * KC; ++r) {
            const int f = f0 + r;
            const bool in = (f >= 0 && f < Tn);
            v[r] = uv_ld2<VEC>(pw + (int64_t)(in ? f : 0) * fstep, ok0 && in, ok1 && in);
          }
        }
#pragma unroll
        for (int r = 0; r < CH + 2 * KC; ++r) {  // frame f0 + r feeds position i with tap k = r - i
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int k = r - i;
            if (k >= 0 && k <= 2 * KC) acc[i] = __ffma2_rn(taps.c[w][k], v[r], acc[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) sb[(p0 + i) * 32 + lane] = acc[i];
  }
  __syncthreads();
  // ---- stage 2: the long filter ----
  constexpr int OW = TC / 4;  // outputs per warp
#pragma unroll 1
  for (int blk = 0; blk < OW / UVF_TTU; ++blk) {
    const int o0 = warp * OW + blk * UVF_TTU;  // first output of the block, relative to t0
    if (t0 + o0 >= t_hi) break;
    float2 acc[UVF_TTU];
#pragma unroll
    for (int i = 0; i < UVF_TTU; ++i) acc[i] = make_float2(0.f, 0.f);
    const float2* src = sb + o0 * 32 + lane;  // b[t0 + o0 - KK + r] = sb[o0 + r]
#pragma unroll
    for (int r = 0; r < UVF_TTU + 2 * KK; ++r) {
      const float2 v = src[r * 32];
#pragma unroll
      for (int i = 0; i < UVF_TTU; ++i) {
        const int j = r - i;
        if (j >= 0 && j <= 2 * KK) acc[i] = __ffma2_rn(taps.h0[j], v, acc[i]);
      }
    }
    float* yo = y + ((int64_t)b * Tn + t0 + o0) * sd + d;
#pragma unroll
    for (int i = 0; i < UVF_TTU; ++i)
      if (t0 + o0 + i < t_hi) uv_st2<VEC>(yo + (int64_t)i * sd, acc[i], ok0, ok1);
  }
}

// backward: tile of TC = 64 - 2 KC rows; stage 1 fills sa[p] = a[s0 - KC + p], p < 64 (the long filter,
// inputs from global / L1); stage 2 applies the short stencils and writes the nw gradient streams
template <int KK, int KC, int NWT, bool VEC>
__global__ void __launch_bounds__(128) uv_fact_bwd_kernel(const __grid_constant__ UvFactTaps<KK, KC, NWT> taps,
                                                          const float* __restrict__ tab, const float* __restrict__ go,
                                                          float* __restrict__ gx, int Tn, int sd, int nw, int K, int t_lo, int t_hi,
                                                          int reshaped, int n_tiles) {
  constexpr int NP = 64, TC = NP - 2 * KC, PW = NP / 4;
  __shared__ __align__(16) float2 sa[NP * 32];
  if ((int)blockIdx.x >= n_tiles) {
    uv_edge_rows<true, VEC>(tab, go, gx, Tn, sd, nw, K, t_lo, t_hi, reshaped, (int)blockIdx.x - n_tiles);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = 2 * (blockIdx.y * 32 + lane);
  const bool ok0 = d < sd, ok1 = d + 1 < sd;
  const int b = blockIdx.z;
  const int s0 = t_lo + blockIdx.x * TC;
  const float* gb = go + (int64_t)b * Tn * sd + d;
  // ---- stage 1: a[r], r = s0 - KC + p ----
#pragma unroll 1
  for (int blk = 0; blk < PW / UVF_TTU; ++blk) {
    const int p0 = warp * PW + blk * UVF_TTU;
    const int r0 = s0 - KC + p0 - KK;  // first input frame of the block
    float2 acc[UVF_TTU];
#pragma unroll
    for (int i = 0; i < UVF_TTU; ++i) acc[i] = make_float2(0.f, 0.f);
    const bool interior = (r0 >= 0) && (r0 + UVF_TTU + 2 * KK <= Tn);
    if (interior) {
#pragma unroll
      for (int r = 0; r < UVF_TTU + 2 * KK; ++r) {
        const float2 v = uv_ld2<VEC>(gb + (int64_t)(r0 + r) * sd, ok0, ok1);
#pragma unroll
        for (int i = 0; i < UVF_TTU; ++i) {
          const int j = r - i;
          if (j >= 0 && j <= 2 * KK) acc[i] = __ffma2_rn(taps.h0[j], v, acc[i]);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < UVF_TTU + 2 * KK; ++r) {
        const int t = r0 + r;
        const bool in = (t >= 0 && t < Tn);
        const float2 v = uv_ld2<VEC>(gb + (int64_t)(in ? t : 0) * sd, ok0 && in, ok1 && in);
#pragma unroll
        for (int i = 0; i < UVF_TTU; ++i) {
          const int j = r - i;
          if (j >= 0 && j <= 2 * KK) acc[i] = __ffma2_rn(taps.h0[j], v, acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < UVF_TTU; ++i) sa[(p0 + i) * 32 + lane] = acc[i];
  }
  __syncthreads();
  // ---- stage 2: g_w[s] = sum_k c_w[k] a[s + k - KC], s = s0 + q: a[s + k - KC] = sa[q + k] ----
  const int64_t nwsd = (int64_t)nw * sd;
  float* ob = gx + (int64_t)b * Tn * nwsd + d;
  for (int q = warp; q < TC; q += 4) {
    const int s = s0 + q;
    if (s >= t_hi) break;
    float2 v[2 * KC + 1];
#pragma unroll
    for (int k = 0; k <= 2 * KC; ++k) v[k] = sa[(q + k) * 32 + lane];
#pragma unroll
    for (int w = 0; w < NWT; ++w) {
      if (w < nw) {
        float2 g = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k <= 2 * KC; ++k) g = __ffma2_rn(taps.c[w][k], v[k], g);
        const int64_t off = reshaped ? ((int64_t)w * Tn + s) * sd : (int64_t)s * nwsd + (int64_t)w * sd;
        uv_st2<VEC>(ob + off, g, ok0, ok1);
      }
    }
  }
}

template <int KK, int KC>
static int uv_fact_launch(const float* tab, const float* h0, const float* c, const float* x, float* y, int B, int Tn, int sd, int nw,
                          int K, int kc, int t_lo, int t_hi, int backward, int reshaped, cudaStream_t st) {
  constexpr int NWT = 3;
  UvFactTaps<KK, KC, NWT> taps;
  for (int j = 0; j <= 2 * KK; ++j) {  // pad the long filter to the instantiated half-width
    const int jj = j - (KK - K);
    const float v = (jj >= 0 && jj <= 2 * K) ? h0[jj] : 0.f;
    taps.h0[j] = make_float2(v, v);
  }
  for (int w = 0; w < NWT; ++w)
    for (int k = 0; k <= 2 * KC; ++k) {
      const int kk = k - (KC - kc);
      const float v = (w < nw && kk >= 0 && kk <= 2 * kc) ? c[w * (2 * kc + 1) + kk] : 0.f;
      taps.c[w][k] = make_float2(v, v);
    }
  const int rows = t_hi - t_lo;
  const int n_edge_blocks = (Tn - rows + 3) / 4;  // edge rows ride along as extra CTAs, one row per warp
  const bool vec = (sd % 2 == 0) && (((uintptr_t)x | (uintptr_t)y) % 8 == 0);
  const unsigned gy = (unsigned)(((sd + 1) / 2 + 31) / 32);
  if (!backward) {
    constexpr int TC = 64;
    constexpr size_t smem = (size_t)(TC + 2 * KK) * 32 * sizeof(float2);
    const int n_tiles = (rows + TC - 1) / TC;
    dim3 grid((unsigned)(n_tiles + n_edge_blocks), gy, (unsigned)B);
    static bool attr_done[64][2];  // per template instance and device: raise the dynamic shared-memory limit once
    int dev_id = 0;
    NNK_CUDA_CHECK(cudaGetDevice(&dev_id));
    bool* attr_set = attr_done[dev_id & 63];
    if (vec) {
      if (!attr_set[1]) {
        NNK_CUDA_CHECK(cudaFuncSetAttribute(uv_fact_fwd_kernel<KK, KC, NWT, TC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[1] = true;
      }
      uv_fact_fwd_kernel<KK, KC, NWT, TC, true><<<grid, 128, smem, st>>>(taps, tab, x, y, Tn, sd, nw, K, t_lo, t_hi, reshaped, n_tiles);
    } else {
      if (!attr_set[0]) {
        NNK_CUDA_CHECK(cudaFuncSetAttribute(uv_fact_fwd_kernel<KK, KC, NWT, TC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[0] = true;
      }
      uv_fact_fwd_kernel<KK, KC, NWT, TC, false><<<grid, 128, smem, st>>>(taps, tab, x, y, Tn, sd, nw, K, t_lo, t_hi, reshaped, n_tiles);
    }
  } else {
    constexpr int TC = 64 - 2 * KC;
    const int n_tiles = (rows + TC - 1) / TC;
    dim3 grid((unsigned)(n_tiles + n_edge_blocks), gy, (unsigned)B);
    if (vec) uv_fact_bwd_kernel<KK, KC, NWT, true><<<grid, 128, 0, st>>>(taps, tab, x, y, Tn, sd, nw, K, t_lo, t_hi, reshaped, n_tiles);
    else uv_fact_bwd_kernel<KK, KC, NWT, false><<<grid, 128, 0, st>>>(taps, tab, x, y, Tn, sd, nw, K, t_lo, t_hi, reshaped, n_tiles);
  }
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

}  // namespace nnk

using namespace nnk;

extern "C" int nnk_uv_band_profile(const void* R, int32_t dtype, int32_t T, int32_t nw, float* profile, void* stream) {
  NNK_REQUIRE(R && profile && T > 0 && nw > 0, NNK_ERR_ARG, "bad argument");
  DeviceGuard guard(R);
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
  DeviceGuard guard(R);
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
  DeviceGuard guard(x);
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == NNK_F32 ? uv_apply<float>(table, x, y, B, T, sd, nw, K, backward, reshaped, 0, 0, st)
                          : uv_apply<double>(table, x, y, B, T, sd, nw, K, backward, reshaped, 0, 0, st);
}

// As nnk_uv_apply (float32 only), but rows [t_lo, t_hi) are computed with the shift-invariant filter
// `taps` (host pointer, nw x (2K+1) floats: the band row of R those rows share); the remaining edge
// rows use the per-row table.  Supported: nw <= 3, K <= 64 (otherwise NNK_ERR_UNSUPPORTED: call
// nnk_uv_apply).
extern "C" int nnk_uv_apply_toeplitz(const void* table, const float* taps, const void* x, void* y, int32_t B, int32_t T,
                                     int32_t sd, int32_t nw, int32_t K, int32_t t_lo, int32_t t_hi, int32_t backward,
                                     int32_t reshaped, void* stream) {
  NNK_REQUIRE(table && taps && x && y, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(B >= 0 && T > 0 && sd >= 0 && nw > 0 && K >= 0 && t_lo >= 0 && t_hi <= T && t_lo <= t_hi, NNK_ERR_ARG, "bad size");
  if (B == 0 || sd == 0) return NNK_OK;
  NNK_REQUIRE(B <= 65535 && (sd + 31) / 32 <= 65535, NNK_ERR_ARG, "batch or static_dim too large for one launch");
  DeviceGuard guard(x);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = uv_apply<float>(table, x, y, B, T, sd, nw, K, backward, reshaped, t_lo, t_hi, st);
  if (rc || t_hi == t_lo) return rc;
  // pad the filter to the next instantiated half-width (extra taps are zero)
  const int KS[] = {8, 16, 24, 32, 48, 64};
  int KK = -1;
  for (int k : KS) if (K <= k) { KK = k; break; }
  NNK_REQUIRE(KK > 0 && nw <= 3, NNK_ERR_UNSUPPORTED, "Toeplitz path supports nw <= 3 and K <= 64");
  float padded[3 * 129];
  for (int w = 0; w < 3; ++w)
    for (int j = 0; j <= 2 * KK; ++j) {
      const int jj = j - (KK - K);
      padded[w * (2 * KK + 1) + j] = (w < nw && jj >= 0 && jj <= 2 * K) ? taps[w * (2 * K + 1) + jj] : 0.f;
    }
  const float* xf = (const float*)x;
  float* yf = (float*)y;
#define NNK_TOEP(KV)                                                                                              \
  case KV:                                                                                                        \
    return nw == 1 ? uv_toeplitz_launch<KV, 1>(padded, xf, yf, B, T, sd, t_lo, t_hi, backward, reshaped, st)      \
         : nw == 2 ? uv_toeplitz_launch_pad<KV, 2>(padded, xf, yf, B, T, sd, t_lo, t_hi, backward, reshaped, st)  \
                   : uv_toeplitz_launch_pad<KV, 3>(padded, xf, yf, B, T, sd, t_lo, t_hi, backward, reshaped, st);
  switch (KK) {
    NNK_TOEP(8) NNK_TOEP(16) NNK_TOEP(24) NNK_TOEP(32) NNK_TOEP(48) NNK_TOEP(64)
  }
#undef NNK_TOEP
  return NNK_ERR_UNSUPPORTED;
}

// Factored variant of nnk_uv_apply_toeplitz (float32): rows [t_lo, t_hi) are computed as ONE long filter
// `h0` (HOST, 2K+1 floats: the band row of the static-window block) combined with short per-window
// stencils `c` (HOST, nw x (2*KC+1) floats) such that the band row of window w equals h0 * c[w]
// (the host fits c from R and checks the residual); the remaining edge rows use the per-row table.
// Supported: nw <= 3, K <= 32, KC <= 2 (otherwise NNK_ERR_UNSUPPORTED: call nnk_uv_apply_toeplitz).
extern "C" int nnk_uv_apply_factored(const void* table, const float* h0, const float* c, const void* x, void* y, int32_t B,
                                     int32_t T, int32_t sd, int32_t nw, int32_t K, int32_t KC, int32_t t_lo, int32_t t_hi,
                                     int32_t backward, int32_t reshaped, void* stream) {
  NNK_REQUIRE(table && h0 && c && x && y, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(B >= 0 && T > 0 && sd >= 0 && nw > 0 && K >= 0 && KC >= 0 && t_lo >= 0 && t_hi <= T && t_lo <= t_hi, NNK_ERR_ARG, "bad size");
  if (B == 0 || sd == 0) return NNK_OK;
  NNK_REQUIRE(nw <= 3 && K <= 32 && KC <= 2, NNK_ERR_UNSUPPORTED, "factored path supports nw <= 3, K <= 32, KC <= 2");
  NNK_REQUIRE(B <= 65535 && (sd + 63) / 64 <= 65535, NNK_ERR_ARG, "batch or static_dim too large for one launch");
  DeviceGuard guard(x);
  cudaStream_t st = (cudaStream_t)stream;
  if (t_hi == t_lo) return uv_apply<float>(table, x, y, B, T, sd, nw, K, backward, reshaped, 0, 0, st);
  const float* xf = (const float*)x;
  float* yf = (float*)y;
  const float* tf = (const float*)table;
#define NNK_FACT(KV, CV) return uv_fact_launch<KV, CV>(tf, h0, c, xf, yf, B, T, sd, nw, K, KC, t_lo, t_hi, backward, reshaped, st)
  if (KC <= 1) {
    if (K <= 16) NNK_FACT(16, 1);
    if (K <= 24) NNK_FACT(24, 1);
    NNK_FACT(32, 1);
  } else {
    if (K <= 16) NNK_FACT(16, 2);
    if (K <= 24) NNK_FACT(24, 2);
    NNK_FACT(32, 2);
  }
#undef NNK_FACT
}
