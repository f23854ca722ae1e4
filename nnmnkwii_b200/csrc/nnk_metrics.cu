// nnk_metrics.cu -- length-masked objective metrics as device reductions.
//
// Replaces the per-utterance Python loops of nnmnkwii/metrics/__init__.py:
//   melcd                  (:27-71)    sum_b sum_{t < len_b} sqrt(sum_d (x - y)^2)      , count = sum_b len_b
//   mean_squared_error     (:74-110)   sum_b sum_{t < len_b} sum_d (x - y)^2
//   lf0_mean_squared_error (:113-165)  sum over frames with src_vuv + tgt_vuv >= 2 of (x - y)^2, count = voiced
//   vuv_error              (:168-190)  sum over frames of (src_vuv != tgt_vuv)
// The kernels return the SUM and the COUNT; the scalar finish (mean, sqrt, 10/ln10*sqrt2) stays in
// the host stub, as in the reference.
//
// Layout: padded (B, T, D) row-major batches, `lengths` (B) on the device (NULL = every frame valid).
// A group of G = min(32, pow2 >= D) lanes owns a frame, so a warp reads 32/G frames = one contiguous
// span per load instruction; 4 frames per group are in flight.  Accumulation in float64.  The result
// is deterministic: every block writes one partial, the last block to finish (ticket) adds the
// partials in index order and resets the ticket, so a workspace that was zero before its first use
// stays usable without a memset per call (one launch per metric).
// HBM bound: 2 * sizeof(T) * D bytes per valid frame (4 scalars per frame for the F0 metrics).
#include "nnk_common.cuh"

namespace nnk {

constexpr int MT_BLOCK = 256;
constexpr int64_t MT_MAX_BLOCKS = 1 << 20;
constexpr int MT_MIN_TILE_FRAMES = 46;  // narrow-frame tiles hold >= 46 frames (D < 128, float64)
constexpr int MT_UNROLL = 4;

struct MetricParams {
  const void* x;
  const void* y;
  const void* xv;  // F0 metrics: src_vuv
  const void* yv;  // F0 metrics: tgt_vuv
  const int32_t* lengths;
  int64_t item_stride;  // elements between batch items
  int64_t frame_stride; // elements between frames
  int B, T, D;
  int chunks;           // blocks per batch item
  int64_t max_blk;      // partial slots in the workspace
  int kind;
  double* partial_sum;
  long long* partial_cnt;
  unsigned int* ticket;
  double* out_sum;
  long long* out_cnt;
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// one partial per block, then the last block folds them in index order (deterministic)
__device__ __forceinline__ void finish(const MetricParams& p, double acc, long long cnt) {
  __shared__ double sh[MT_BLOCK / 32];
  __shared__ bool last;
  const double s = block_sum(acc, sh);
  const double c = block_sum((double)cnt, sh);  // counts < 2^53: exact in float64
  if (threadIdx.x == 0) {
    p.partial_sum[blockIdx.x] = s;
    p.partial_cnt[blockIdx.x] = (long long)c;
    __threadfence();
    const unsigned int done = atomicAdd(p.ticket, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double ts = 0.0, tc = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    ts += __ldcg(p.partial_sum + i);
    tc += (double)__ldcg(p.partial_cnt + i);
  }
  ts = block_sum(ts, sh);
  tc = block_sum(tc, sh);
  if (threadIdx.x == 0) {
    *p.out_sum = ts;
    *p.out_cnt = (long long)tc;
    *p.ticket = 0u;  // ready for the next call on this workspace
  }
}

// kind 0: sum of per-frame Euclidean norms (melcd), kind 1: sum of squared differences
template <typename T, int G>
__global__ void __launch_bounds__(MT_BLOCK) frame_metric_kernel(const __grid_constant__ MetricParams p) {
  constexpr int FPW = 32 / G;                        // frames per warp per step
  constexpr int FPB = (MT_BLOCK / 32) * FPW;         // frames per block per step
  const int b = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int len = p.lengths ? min(max(p.lengths[b], 0), p.T) : p.T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / G, dl = lane % G;
  const T* X = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.item_stride;
  const T* Y = reinterpret_cast<const T*>(p.y) + (int64_t)b * p.item_stride;
  double acc = 0.0;
  long long cnt = 0;
  const int step = p.chunks * FPB * MT_UNROLL;
  for (int t0 = chunk * FPB * MT_UNROLL; t0 < len; t0 += step) {
    double s[MT_UNROLL];
    bool ok[MT_UNROLL];
    const T* xr[MT_UNROLL];
    const T* yr[MT_UNROLL];
#pragma unroll
    for (int u = 0; u < MT_UNROLL; ++u) {
      const int t = t0 + u * FPB + warp * FPW + sub;
      ok[u] = t < len;
      const int64_t o = (int64_t)(ok[u] ? t : 0) * p.frame_stride;
      xr[u] = X + o;
      yr[u] = Y + o;
      s[u] = 0.0;
    }
    // all 2 * MT_UNROLL loads of a pass over d are issued before the first use
    for (int d0 = 0; d0 < p.D; d0 += G) {
      const int d = d0 + dl;
      const bool in = d < p.D;
      T xv[MT_UNROLL], yv[MT_UNROLL];
#pragma unroll
      for (int u = 0; u < MT_UNROLL; ++u) {
        xv[u] = T(0); yv[u] = T(0);
        if (in && ok[u]) { xv[u] = ld_stream(xr[u] + d); yv[u] = ld_stream(yr[u] + d); }
      }
#pragma unroll
      for (int u = 0; u < MT_UNROLL; ++u) {
        const double z = (double)(xv[u] - yv[u]);  // difference in the input dtype, like z = X - Y (:55)
        s[u] = fma(z, z, s[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < MT_UNROLL; ++u) {
#pragma unroll
      for (int o = G / 2; o; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
      if (ok[u] && dl == 0) {
        acc += p.kind == 0 ? sqrt(s[u]) : s[u];
        cnt += 1;
      }
    }
  }
  finish(p, acc, cnt);
}

// Narrow frames (D < 128, frames contiguous): the warp-per-frame mapping above spends ~50 warp
// instructions per frame on shuffles and the square root.  Here a block streams a TILE of F whole
// frames (F * D contiguous elements of X and of Y, coalesced, 8 + 8 loads in flight per thread) into
// shared memory as squared differences, then thread f adds up frame f (float64) and takes its square
// root: ~7 warp instructions per frame.
constexpr int MT_TILE_ELEMS = 11776;  // float32 elements of one tile (46 KB: static + dynamic stay under 48 KB)
constexpr int MT_TBLOCK = 128;        // threads (= frames per tile): small blocks, many per SM

template <typename T> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };

__device__ __forceinline__ float4 sqdiff(float4 a, float4 b) {
  float4 z = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
  return make_float4(z.x * z.x, z.y * z.y, z.z * z.z, z.w * z.w);
}
__device__ __forceinline__ double2 sqdiff(double2 a, double2 b) {
  double2 z = make_double2(a.x - b.x, a.y - b.y);
  return make_double2(z.x * z.x, z.y * z.y);
}

template <typename T>
__global__ void __launch_bounds__(MT_TBLOCK) frame_metric_tile_kernel(const __grid_constant__ MetricParams p, const int F) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::N;
  extern __shared__ __align__(16) unsigned char tile_raw[];
  T* sq0 = reinterpret_cast<T*>(tile_raw);
  const int b = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int len = p.lengths ? min(max(p.lengths[b], 0), p.T) : p.T;
  const T* X = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.item_stride;
  const T* Y = reinterpret_cast<const T*>(p.y) + (int64_t)b * p.item_stride;
  const int D = p.D;
  double acc = 0.0;
  long long cnt = 0;
  constexpr int UL = 8;   // scalar loads in flight per array and thread
  constexpr int UV = 4;   // 16-byte loads in flight per array and thread
  for (int f0 = chunk * F; f0 < len; f0 += p.chunks * F) {
    const int nf = min(F, len - f0);
    const int n = nf * D;
    const T* xt = X + (int64_t)f0 * D;
    const T* yt = Y + (int64_t)f0 * D;
    // element e of the tile lives at sq[e] = sq0[mis + e]: shared memory mirrors the 16-byte phase of
    // the global rows, so the aligned middle part moves as 16-byte vectors on both sides
    const int mis = (int)((reinterpret_cast<uintptr_t>(xt) / sizeof(T)) % VN);
    const bool vec_ok = (int)((reinterpret_cast<uintptr_t>(yt) / sizeof(T)) % VN) == mis;
    T* sq = sq0 + mis;
    const int head = vec_ok ? min(n, (VN - mis) % VN) : n;
    const int nv = vec_ok ? (n - head) / VN : 0;
    const int tail0 = head + nv * VN;
    auto scalar_range = [&](int lo, int hi) {
      for (int e0 = lo + threadIdx.x; e0 < hi; e0 += MT_TBLOCK * UL) {
        T xv[UL], yv[UL];
#pragma unroll
        for (int u = 0; u < UL; ++u) {
          const int e = e0 + u * MT_TBLOCK;
          xv[u] = T(0); yv[u] = T(0);
          if (e < hi) { xv[u] = ld_stream(xt + e); yv[u] = ld_stream(yt + e); }
        }
#pragma unroll
        for (int u = 0; u < UL; ++u) {
          const int e = e0 + u * MT_TBLOCK;
          const T z = xv[u] - yv[u];  // difference and square in the input dtype, like z * z (:55-56)
          if (e < hi) sq[e] = z * z;
        }
      }
    };
    scalar_range(0, head);
    {
      const V* xv4 = reinterpret_cast<const V*>(xt + head);
      const V* yv4 = reinterpret_cast<const V*>(yt + head);
      V* sv4 = reinterpret_cast<V*>(sq + head);
      for (int i0 = threadIdx.x; i0 < nv; i0 += MT_TBLOCK * UV) {
        V a[UV], c[UV];
#pragma unroll
        for (int u = 0; u < UV; ++u) {
          const int i = i0 + u * MT_TBLOCK;
          if (i < nv) { a[u] = __ldcs(xv4 + i); c[u] = __ldcs(yv4 + i); }
        }
#pragma unroll
        for (int u = 0; u < UV; ++u) {
          const int i = i0 + u * MT_TBLOCK;
          if (i < nv) sv4[i] = sqdiff(a[u], c[u]);
        }
      }
    }
    scalar_range(tail0, n);
    __syncthreads();
    for (int f = threadIdx.x; f < nf; f += MT_TBLOCK) {
      const T* r = sq + f * D;
      double s = 0.0;
      for (int d = 0; d < D; ++d) s += (double)r[d];
      acc += p.kind == 0 ? sqrt(s) : s;
      cnt += 1;
    }
    __syncthreads();
  }
  finish(p, acc, cnt);
}

// kind 0: lf0 MSE (log domain), 1: lf0 MSE (linear domain: exp first), 2: vuv error
template <typename T>
__global__ void __launch_bounds__(MT_BLOCK) f0_metric_kernel(const __grid_constant__ MetricParams p) {
  const int b = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int len = p.lengths ? min(max(p.lengths[b], 0), p.T) : p.T;
  const int64_t base = (int64_t)b * p.item_stride;
  const T* X = reinterpret_cast<const T*>(p.x) + base;
  const T* Y = reinterpret_cast<const T*>(p.y) + base;
  const T* XV = reinterpret_cast<const T*>(p.xv) + base;
  const T* YV = reinterpret_cast<const T*>(p.yv) + base;
  double acc = 0.0;
  long long cnt = 0;
  const int stride = p.chunks * MT_BLOCK;
  for (int t0 = chunk * MT_BLOCK + threadIdx.x; t0 < len; t0 += stride * MT_UNROLL) {
    T xv[MT_UNROLL], yv[MT_UNROLL], x[MT_UNROLL], y[MT_UNROLL];
    bool ok[MT_UNROLL];
#pragma unroll
    for (int u = 0; u < MT_UNROLL; ++u) {
      const int64_t t = (int64_t)t0 + (int64_t)u * stride;
      ok[u] = t < len;
      const int64_t i = (ok[u] ? t : 0) * p.frame_stride;
      xv[u] = ld_stream(XV + i);
      yv[u] = ld_stream(YV + i);
      x[u] = T(0); y[u] = T(0);
      if (p.kind != 2) { x[u] = ld_stream(X + i); y[u] = ld_stream(Y + i); }
    }
#pragma unroll
    for (int u = 0; u < MT_UNROLL; ++u) {
      if (!ok[u]) continue;
      if (p.kind == 2) {
        acc += (xv[u] != yv[u]) ? 1.0 : 0.0;
        cnt += 1;
      } else if (xv[u] + yv[u] >= T(2)) {  // both voiced (metrics/__init__.py:143,155)
        T a = x[u], b2 = y[u];
        if (p.kind == 1) { a = exp(a); b2 = exp(b2); }
        const double z = (double)(a - b2);
        acc = fma(z, z, acc);
        cnt += 1;
      }
    }
  }
  finish(p, acc, cnt);
}

static int pick_chunks(int B, int T, int frames_per_block_step, int per_sm = 8) {
  const int64_t want = 148 * per_sm;  // blocks in flight per SM
  int64_t c = (want + B - 1) / B;
  const int64_t maxc = (T + frames_per_block_step - 1) / frames_per_block_step;
  if (c > maxc) c = maxc;
  if (c < 1) c = 1;
  return (int)c;
}

template <typename T, int G>
static void launch_frame(const MetricParams& p, cudaStream_t st) {
  frame_metric_kernel<T, G><<<(unsigned)(p.B * p.chunks), MT_BLOCK, 0, st>>>(p);
}

template <typename T>
static void dispatch_frame(MetricParams& p, cudaStream_t st) {
  if (p.D < 128 && p.frame_stride == p.D) {
    const int cap = MT_TILE_ELEMS * 4 / (int)sizeof(T) / p.D;  // frames per 48 KB tile
    const int F = cap < MT_TBLOCK ? cap : MT_TBLOCK;
    // a few tiles per block: many more blocks than SM slots, so the hardware block scheduler balances
    // ragged lengths; the partial of block i always covers the same tiles (deterministic)
    int64_t c = ((p.T + F - 1) / F + 3) / 4;  // ~4 tiles per block
    if ((int64_t)p.B * c > p.max_blk) c = p.max_blk / p.B;
    p.chunks = (int)(c < 1 ? 1 : c);
    frame_metric_tile_kernel<T><<<(unsigned)(p.B * p.chunks), MT_TBLOCK, (size_t)F * p.D * sizeof(T) + 16, st>>>(p, F);
    return;
  }
  int G = 1;
  while (G < 32 && G < p.D) G <<= 1;
  p.chunks = pick_chunks(p.B, p.T, (MT_BLOCK / 32) * (32 / G) * MT_UNROLL);
  switch (G) {
    case 1: launch_frame<T, 1>(p, st); break;
    case 2: launch_frame<T, 2>(p, st); break;
    case 4: launch_frame<T, 4>(p, st); break;
    case 8: launch_frame<T, 8>(p, st); break;
    case 16: launch_frame<T, 16>(p, st); break;
    default: launch_frame<T, 32>(p, st); break;
  }
}

static int carve(void* workspace, int64_t workspace_bytes, int64_t blocks, MetricParams& p) {
  const int64_t need = 64 + blocks * 16;
  NNK_REQUIRE(workspace && workspace_bytes >= need, NNK_ERR_WORKSPACE, "metric workspace too small");
  char* w = reinterpret_cast<char*>(workspace);
  p.max_blk = blocks;
  p.ticket = reinterpret_cast<unsigned int*>(w);
  p.partial_sum = reinterpret_cast<double*>(w + 64);
  p.partial_cnt = reinterpret_cast<long long*>(w + 64 + blocks * 8);
  return NNK_OK;
}

}  // namespace nnk

using namespace nnk;

static int64_t max_blocks(int64_t B, int64_t T) {
  if (B < 1) B = 1;
  int64_t tiles = B * ((T + MT_MIN_TILE_FRAMES - 1) / MT_MIN_TILE_FRAMES);
  if (tiles > MT_MAX_BLOCKS / 2) tiles = MT_MAX_BLOCKS / 2;
  const int64_t spread = B + 148 * 16;  // the other kernels: B * chunks <= B + 148*16
  return tiles > spread ? tiles : spread;
}

extern "C" int64_t nnk_metric_workspace_bytes(int32_t B, int32_t T) { return 64 + max_blocks(B, T) * 16; }

extern "C" int nnk_frame_metric(const void* X, const void* Y, int32_t dtype, int32_t B, int32_t T, int32_t D,
                                int64_t item_stride, int64_t frame_stride, const int32_t* lengths, int32_t kind,
                                double* sum_out, int64_t* count_out, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  NNK_REQUIRE(sum_out && count_out, NNK_ERR_ARG, "NULL output");
  DeviceGuard guard(sum_out);
  NNK_REQUIRE(dtype == NNK_F32 || dtype == NNK_F64, NNK_ERR_ARG, "bad dtype");
  NNK_REQUIRE(kind == 0 || kind == 1, NNK_ERR_ARG, "bad kind");
  NNK_REQUIRE(B >= 0 && T >= 0 && D >= 0 && T <= (1 << 30), NNK_ERR_ARG, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0 || T == 0 || D == 0) {
    NNK_CUDA_CHECK(cudaMemsetAsync(sum_out, 0, sizeof(double), st));
    NNK_CUDA_CHECK(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st));
    return NNK_OK;
  }
  NNK_REQUIRE(X && Y, NNK_ERR_ARG, "NULL input");
  NNK_REQUIRE((int64_t)B + 148 * 16 < MT_MAX_BLOCKS / 2, NNK_ERR_ARG, "batch too large for one launch");
  MetricParams p{};
  p.x = X; p.y = Y; p.lengths = lengths; p.item_stride = item_stride; p.frame_stride = frame_stride;
  p.B = B; p.T = T; p.D = D; p.kind = kind; p.out_sum = sum_out; p.out_cnt = reinterpret_cast<long long*>(count_out);
  const int rc = carve(workspace, workspace_bytes, max_blocks(B, T), p);
  if (rc) return rc;
  if (dtype == NNK_F32) dispatch_frame<float>(p, st);
  else dispatch_frame<double>(p, st);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

extern "C" int nnk_f0_metric(const void* src_f0, const void* src_vuv, const void* tgt_f0, const void* tgt_vuv,
                             int32_t dtype, int32_t B, int32_t T, int64_t item_stride, int64_t frame_stride,
                             const int32_t* lengths, int32_t kind, double* sum_out, int64_t* count_out,
                             void* workspace, int64_t workspace_bytes, void* stream) {
  NNK_REQUIRE(sum_out && count_out, NNK_ERR_ARG, "NULL output");
  DeviceGuard guard(sum_out);
  NNK_REQUIRE(dtype == NNK_F32 || dtype == NNK_F64, NNK_ERR_ARG, "bad dtype");
  NNK_REQUIRE(kind >= 0 && kind <= 2, NNK_ERR_ARG, "bad kind");
  NNK_REQUIRE(B >= 0 && T >= 0 && T <= (1 << 30), NNK_ERR_ARG, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0 || T == 0) {
    NNK_CUDA_CHECK(cudaMemsetAsync(sum_out, 0, sizeof(double), st));
    NNK_CUDA_CHECK(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st));
    return NNK_OK;
  }
  NNK_REQUIRE(src_vuv && tgt_vuv && (kind == 2 || (src_f0 && tgt_f0)), NNK_ERR_ARG, "NULL input");
  NNK_REQUIRE((int64_t)B + 148 * 16 < MT_MAX_BLOCKS / 2, NNK_ERR_ARG, "batch too large for one launch");
  MetricParams p{};
  p.x = src_f0; p.y = tgt_f0; p.xv = src_vuv; p.yv = tgt_vuv; p.lengths = lengths;
  p.item_stride = item_stride; p.frame_stride = frame_stride;
  p.B = B; p.T = T; p.D = 1; p.kind = kind; p.out_sum = sum_out; p.out_cnt = reinterpret_cast<long long*>(count_out);
  p.chunks = pick_chunks(B, T, MT_BLOCK * MT_UNROLL);
  const int rc = carve(workspace, workspace_bytes, max_blocks(B, T), p);
  if (rc) return rc;
  if (dtype == NNK_F32) f0_metric_kernel<float><<<(unsigned)(B * p.chunks), MT_BLOCK, 0, st>>>(p);
  else f0_metric_kernel<double><<<(unsigned)(B * p.chunks), MT_BLOCK, 0, st>>>(p);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}
