// nnk_dtw.cu -- DTW / FastDTW alignment on sm_100a: one thread block per utterance pair walks the
// anti-diagonals of the accumulated-cost matrix (wavefront), local cost computed in registers.
//
// Replaces the per-pair body of DTWAligner.transform / IterativeDTWAligner.transform
// (preprocessing/alignment.py:48-54, :136-143):
//       dist, path = fastdtw(x, y, radius=self.radius, dist=self.dist)
// i.e. the third-party `fastdtw` package (slaypni/fastdtw; unpinned dependency, setup.py:139) whose
// published algorithm is restated here: inputs as float64; recursion until a series is shorter
// than radius+2 (full DTW there); halve by averaging adjacent frames; project the coarse path,
// dilated by `radius`, onto the fine grid; DP restricted to that per-row window
//       D[i,j] = dt + first-min(D[i-1,j], D[i,j-1], D[i-1,j-1])      (order: up, left, diagonal)
// and backtrack.  radius < 0 selects the exact DTW (full window) -- the same kernel with the window
// test compiled out.  The whole recursion of one pair runs inside one CTA: coarse levels, window
// expansion, DP and backtrack never leave the SM except for the float64 level copies of the series
// (L2-resident scratch) and, in exact mode, the 1 byte/cell back-pointers.
//
// Local cost (cost_kind): 1 = metrics.melcd(x, y) on two frames = (10/ln10*sqrt2) * sqrt(sum((x-y)^2))
// (metrics/__init__.py:5,52-57), 0 = the default lambda x, y: norm(x - y) (alignment.py:35).  The
// squared differences are summed in float64 in the exact order of numpy's pairwise add-reduce so
// that costs, ties and therefore back-track indices are bit-identical to the CPU oracle.
#include <math_constants.h>

#include "nnk_common.cuh"

namespace nnk {

struct DtwParams {
  const void* X;
  const void* Y;
  int is_f64;
  int n_pairs;
  int64_t x_pair_stride, y_pair_stride;
  int x_ld, y_ld, D;
  const int32_t* len_x;
  const int32_t* len_y;
  const int32_t* order;
  int cost_kind, radius;
  int32_t* path_i;
  int32_t* path_j;
  int path_ld;
  int32_t* path_len;
  double* dist;
  long long* cells;
  int max_tx, max_ty;
  unsigned char* ws;
  size_t ws_pair_bytes, series_doubles, bp_bytes;
  int smem_bp_cap;  // bytes of back-pointer space available in shared memory (fast mode)
  double logdb;
};

// numpy DOUBLE_pairwise_sum order over a[k] = (x[k]-y[k])^2 without materialising a[]
__device__ __forceinline__ double sq(const double* __restrict__ x, const double* __restrict__ y, int k) {
  const double z = __dsub_rn(x[k], y[k]);
  return __dmul_rn(z, z);  // never contracted into an FMA: numpy multiplies, then adds
}
__device__ __forceinline__ double strided8(const double* x, const double* y, int j, int n8) {
  double r = sq(x, y, j);
  for (int i = 8; i < n8; i += 8) r = __dadd_rn(r, sq(x, y, i + j));
  return r;
}
__device__ double pairwise_block(const double* x, const double* y, int n) {  // n <= 128
  if (n < 8) {
    double res = -0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, sq(x, y, i));
    return res;
  }
  const int n8 = n - (n % 8);
  const double s01 = __dadd_rn(strided8(x, y, 0, n8), strided8(x, y, 1, n8));
  const double s23 = __dadd_rn(strided8(x, y, 2, n8), strided8(x, y, 3, n8));
  const double s0123 = __dadd_rn(s01, s23);
  const double s45 = __dadd_rn(strided8(x, y, 4, n8), strided8(x, y, 5, n8));
  const double s67 = __dadd_rn(strided8(x, y, 6, n8), strided8(x, y, 7, n8));
  double res = __dadd_rn(s0123, __dadd_rn(s45, s67));
  for (int i = n8; i < n; ++i) res = __dadd_rn(res, sq(x, y, i));
  return res;
}
__device__ double pairwise_sumsq(const double* x, const double* y, int n) {
  if (n <= 128) return pairwise_block(x, y, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(pairwise_sumsq(x, y, n2), pairwise_sumsq(x + n2, y + n2, n - n2));
}

__device__ __forceinline__ double local_cost(const double* x, const double* y, int D, int kind, double logdb) {
  const double r = sqrt(pairwise_sumsq(x, y, D));
  return kind == 1 ? __dmul_rn(logdb, r) : r;
}

template <int BLOCK>
__device__ __forceinline__ void block_sync() {
  if (BLOCK == 32) __syncwarp();
  else __syncthreads();
}

// One CTA per pair.  FULL = exact DTW (radius < 0): no window arrays, back-pointers in global.
template <int BLOCK, bool FULL>
__global__ void __launch_bounds__(BLOCK) dtw_kernel(const DtwParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int pair = p.order ? p.order[blockIdx.x] : blockIdx.x;
  const int Tx0 = p.len_x[pair], Ty0 = p.len_y[pair];
  const int D = p.D;
  if (Tx0 <= 0 || Ty0 <= 0) {
    if (tid == 0) { p.path_len[pair] = 0; p.dist[pair] = 0.0; if (p.cells) p.cells[pair] = 0; }
    return;
  }
  // ---- shared memory carve-up -------------------------------------------------------------------
  const int mtx = p.max_tx;
  double* Dbuf = reinterpret_cast<double*>(smem);  // [3][mtx]
  int* lo = reinterpret_cast<int*>(Dbuf + 3 * (size_t)mtx);
  int* hi = lo + (FULL ? 0 : mtx);
  int* off = hi + (FULL ? 0 : mtx);             // [mtx + 1] row offsets into bp (fast mode)
  int* jmn = off + (FULL ? 0 : mtx + 1);        // [mtx/2 + 1] coarse path extents per coarse row
  int* jmx = jmn + (FULL ? 0 : mtx / 2 + 1);
  unsigned char* bp_s = reinterpret_cast<unsigned char*>(jmx + (FULL ? 0 : mtx / 2 + 1));
  __shared__ int s_n;
  __shared__ long long s_cells;

  unsigned char* wsp = p.ws + (size_t)pair * p.ws_pair_bytes;
  double* xs = reinterpret_cast<double*>(wsp);
  double* ys = xs + p.series_doubles;
  unsigned char* bp_g = reinterpret_cast<unsigned char*>(ys + p.series_doubles);

  // ---- level 0 = the inputs widened to float64 (fastdtw: np.asanyarray(x, dtype='float')) --------
  {
    const int64_t xb = (int64_t)pair * p.x_pair_stride, yb = (int64_t)pair * p.y_pair_stride;
    for (int e = tid; e < Tx0 * D; e += BLOCK) {
      const int64_t src = xb + (int64_t)(e / D) * p.x_ld + (e % D);
      xs[e] = p.is_f64 ? reinterpret_cast<const double*>(p.X)[src] : (double)reinterpret_cast<const float*>(p.X)[src];
    }
    for (int e = tid; e < Ty0 * D; e += BLOCK) {
      const int64_t src = yb + (int64_t)(e / D) * p.y_ld + (e % D);
      ys[e] = p.is_f64 ? reinterpret_cast<const double*>(p.Y)[src] : (double)reinterpret_cast<const float*>(p.Y)[src];
    }
  }
  // ---- coarser levels: __reduce_by_half, until one side is shorter than radius + 2 ----------------
  int nlev = 1;
  if (!FULL) {
    const int min_time = p.radius + 2;
    int tx = Tx0, ty = Ty0;
    size_t xo = 0, yo = 0;
    block_sync<BLOCK>();
    while (tx >= min_time && ty >= min_time) {
      const int hx = tx / 2, hy = ty / 2;
      const double* xin = xs + xo; const double* yin = ys + yo;
      double* xout = xs + xo + (size_t)tx * D; double* yout = ys + yo + (size_t)ty * D;
      for (int e = tid; e < hx * D; e += BLOCK) {
        const int i = e / D, k = e % D;
        xout[e] = (xin[(size_t)(2 * i) * D + k] + xin[(size_t)(2 * i + 1) * D + k]) / 2;
      }
      for (int e = tid; e < hy * D; e += BLOCK) {
        const int i = e / D, k = e % D;
        yout[e] = (yin[(size_t)(2 * i) * D + k] + yin[(size_t)(2 * i + 1) * D + k]) / 2;
      }
      xo += (size_t)tx * D; yo += (size_t)ty * D;
      tx = hx; ty = hy;
      ++nlev;
      __threadfence_block();
      block_sync<BLOCK>();
    }
  }
  if (tid == 0) s_cells = 0;
  block_sync<BLOCK>();

  // ---- levels, coarsest first ---------------------------------------------------------------------
  for (int lev = nlev - 1; lev >= 0; --lev) {
    int Tx = Tx0, Ty = Ty0;
    size_t xo = 0, yo = 0;
    for (int l = 0; l < lev; ++l) { xo += (size_t)Tx * D; yo += (size_t)Ty * D; Tx /= 2; Ty /= 2; }
    const double* xl = xs + xo;
    const double* yl = ys + yo;
    bool bp_in_smem = false;
    if (!FULL) {
      // window: full rectangle at the coarsest level, else __expand_window of the coarser path
      if (lev == nlev - 1) {
        for (int i = tid; i < Tx; i += BLOCK) { lo[i] = 0; hi[i] = Ty; }
      } else {
        const int cx = Tx / 2, r = p.radius;  // coarse rows 0..cx-1 all carry path cells
        for (int i = tid; i < Tx; i += BLOCK) {
          const int a = i >> 1;
          int mn = INT_MAX, mx = -1;
          for (int aa = max(0, a - r); aa <= min(cx - 1, a + r); ++aa) { mn = min(mn, jmn[aa]); mx = max(mx, jmx[aa]); }
          int l = 2 * (mn - r), h = 2 * (mx + r) + 2;
          if (mx < 0) { l = 0; h = 0; }
          lo[i] = max(0, l);
          hi[i] = min(Ty, h);
        }
      }
      block_sync<BLOCK>();
      if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < Tx; ++i) { off[i] = acc; acc += max(0, hi[i] - lo[i]); }
        off[Tx] = acc;
        s_cells += acc;
      }
      block_sync<BLOCK>();
      bp_in_smem = off[Tx] <= p.smem_bp_cap;
      for (int a = tid; a < Tx; a += BLOCK) { if (a < mtx / 2 + 1) { jmn[a] = INT_MAX; jmx[a] = -1; } }
    } else {
      if (tid == 0) s_cells += (long long)Tx * Ty;
    }
    unsigned char* bp = bp_in_smem ? bp_s : bp_g;
    block_sync<BLOCK>();

    // ---- wavefront over anti-diagonals k = i + j -----------------------------------------------
    int imin = 0, imax = -1;
    const int ndiag = Tx + Ty - 1;
    for (int k = 0; k < ndiag; ++k) {
      if (FULL) {
        imin = max(0, k - (Ty - 1));
        imax = min(Tx - 1, k);
      } else {
        while (imax + 1 < Tx && imax + 1 + lo[imax + 1] <= k) ++imax;
        while (imin < Tx && imin + hi[imin] <= k) ++imin;
      }
      double* dk = Dbuf + (size_t)(k % 3) * mtx;
      const double* d1 = Dbuf + (size_t)((k + 2) % 3) * mtx;  // diagonal k-1
      const double* d2 = Dbuf + (size_t)((k + 1) % 3) * mtx;  // diagonal k-2
      for (int i = imin + tid; i <= imax; i += BLOCK) {
        const int j = k - i;
        const double dt = local_cost(xl + (size_t)i * D, yl + (size_t)j * D, D, p.cost_kind, p.logdb);
        bool vu, vl, vd;
        if (FULL) {
          vu = i > 0; vl = j > 0; vd = i > 0 && j > 0;
        } else {
          vu = i > 0 && j >= lo[i - 1] && j < hi[i - 1];
          vl = j - 1 >= lo[i];
          vd = i > 0 && j - 1 >= lo[i - 1] && j - 1 < hi[i - 1];
        }
        const double up = (vu ? d1[i - 1] : CUDART_INF) + dt;
        const double left = (vl ? d1[i] : CUDART_INF) + dt;
        const double diag = ((i == 0 && j == 0) ? 0.0 : (vd ? d2[i - 1] : CUDART_INF)) + dt;
        double best = up;
        unsigned char dir = 0;
        if (left < best) { best = left; dir = 1; }
        if (diag < best) { best = diag; dir = 2; }
        dk[i] = best;
        const size_t cell = FULL ? (size_t)i * Ty + j : (size_t)(off[i] + j - lo[i]);
        bp[cell] = dir;
      }
      if (!bp_in_smem) __threadfence_block();
      block_sync<BLOCK>();
    }

    // ---- backtrack (one thread; the path is a dependent chain) --------------------------------------
    if (tid == 0) {
      int i = Tx - 1, j = Ty - 1, n = 0;
      bool ok = true;
      if (lev == 0) p.dist[pair] = Dbuf[(size_t)((ndiag - 1) % 3) * mtx + (Tx - 1)];
      int32_t* pi = p.path_i + (size_t)pair * p.path_ld;
      int32_t* pj = p.path_j + (size_t)pair * p.path_ld;
      while (i >= 0 && j >= 0) {
        if (!FULL && (j < lo[i] || j >= hi[i])) { ok = false; break; }
        if (lev == 0) {
          if (n >= p.path_ld) { ok = false; break; }
          pi[n] = i; pj[n] = j;
        } else if (!FULL) {
          jmn[i] = min(jmn[i], j);
          jmx[i] = max(jmx[i], j);
        }
        ++n;
        const size_t cell = FULL ? (size_t)i * Ty + j : (size_t)(off[i] + j - lo[i]);
        const unsigned char dir = bp[cell];
        if (dir == 0) --i;
        else if (dir == 1) --j;
        else { --i; --j; }
        if (i < 0 || j < 0) break;
      }
      s_n = ok ? n : -1;
    }
    block_sync<BLOCK>();
  }
  // ---- finalise: reverse the level-0 path in place ------------------------------------------------------
  const int n = s_n;
  if (n > 0) {
    int32_t* pi = p.path_i + (size_t)pair * p.path_ld;
    int32_t* pj = p.path_j + (size_t)pair * p.path_ld;
    for (int a = tid; a < n / 2; a += BLOCK) {
      const int b = n - 1 - a;
      const int32_t ti = pi[a], tj = pj[a];
      pi[a] = pi[b]; pj[a] = pj[b];
      pi[b] = ti; pj[b] = tj;
    }
  }
  if (tid == 0) {
    p.path_len[pair] = n;
    if (p.cells) p.cells[pair] = s_cells;
  }
}

// X_aligned[n, :L] = X[n, path[n, :L]]; rows L.. are zero (alignment.py:52-54, 72-73)
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ X, int64_t x_pair_stride, int x_ld, const int32_t* __restrict__ path,
                                   int path_ld, const int32_t* __restrict__ path_len, T* __restrict__ out,
                                   int64_t out_pair_stride, int out_rows, int D) {
  const int pair = blockIdx.y;
  const int L = path_len[pair];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)out_rows * D; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / D), k = (int)(e % D);
    T v = T(0);
    if (r < L) v = X[(int64_t)pair * x_pair_stride + (int64_t)path[(int64_t)pair * path_ld + r] * x_ld + k];
    out[(int64_t)pair * out_pair_stride + (int64_t)r * D + k] = v;
  }
}

// trim_zeros_frames(x, eps, trim='b') lengths (preprocessing/generic.py:312-323): last frame whose
// sum_d |x| (accumulated in the input dtype, numpy pairwise order for D <= 128) is >= eps, plus one.
template <typename T>
__device__ T abs_pairwise(const T* a, int n) {
  if (n < 8) {
    T res = T(0);
    for (int i = 0; i < n; ++i) res += fabs(a[i]);
    return res;
  }
  if (n <= 128) {
    T r[8];
    for (int j = 0; j < 8; ++j) r[j] = fabs(a[j]);
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += fabs(a[i + j]);
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += fabs(a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return abs_pairwise(a, n2) + abs_pairwise(a + n2, n - n2);
}

template <typename T>
__global__ void trim_len_kernel(const T* __restrict__ X, int64_t pair_stride, int ld, int Tn, int D, T eps, int32_t* __restrict__ len) {
  const int pair = blockIdx.x;
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = 0;
  __syncthreads();
  int last = 0;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    const T s = abs_pairwise(X + (int64_t)pair * pair_stride + (int64_t)t * ld, D);
    if (!(s < eps)) last = max(last, t + 1);
  }
  atomicMax(&s_last, last);
  __syncthreads();
  if (threadIdx.x == 0) len[pair] = s_last;
}

static size_t dtw_series_doubles(int max_t, int D) { return (size_t)2 * (size_t)max_t * D + 8; }

static size_t dtw_smem_bytes(int max_tx, bool full, int bp_cap) {
  size_t b = sizeof(double) * 3 * (size_t)max_tx;
  if (!full) b += sizeof(int) * ((size_t)max_tx * 2 + (max_tx + 1) + 2 * (max_tx / 2 + 1)) + (size_t)bp_cap;
  return b + 16;
}

static size_t dtw_fast_cells_bound(int max_tx, int max_ty, int radius) {
  // window cells of one level <= 4 (2r+1) (coarse path length + 2r+1); the coarsest level is a full
  // rectangle of at most (2(r+2)) x ... cells, covered by the same bound for every realistic size
  const size_t r2 = (size_t)(2 * radius + 1);
  return 4 * r2 * ((size_t)(max_tx + max_ty) / 2 + r2 + 2) + 64;
}

}  // namespace nnk

using namespace nnk;

extern "C" size_t nnk_dtw_workspace_bytes(int32_t n_pairs, int32_t max_tx, int32_t max_ty, int32_t D, int32_t radius) {
  const int mt = max_tx > max_ty ? max_tx : max_ty;
  size_t per = 2 * dtw_series_doubles(mt, D) * sizeof(double);
  if (radius < 0) per += (size_t)max_tx * (size_t)max_ty;
  else per += (size_t)max_tx * (size_t)max_ty < ((size_t)64 << 20) ? (size_t)max_tx * (size_t)max_ty : ((size_t)64 << 20);
  per = (per + 255) / 256 * 256;
  return per * (size_t)n_pairs;
}

extern "C" int nnk_dtw_align(const nnk_dtw_args_t* a, void* stream) {
  NNK_REQUIRE(a != nullptr, NNK_ERR_ARG, "args is NULL");
  NNK_REQUIRE(a->n_pairs >= 0 && a->D > 0 && a->max_tx >= 0 && a->max_ty >= 0, NNK_ERR_ARG, "bad size");
  if (a->n_pairs == 0 || a->max_tx == 0 || a->max_ty == 0) return NNK_OK;
  NNK_REQUIRE(a->X && a->Y && a->len_x && a->len_y && a->path_i && a->path_j && a->path_len && a->dist && a->workspace,
              NNK_ERR_ARG, "NULL device pointer");
  NNK_REQUIRE(a->dtype == NNK_F32 || a->dtype == NNK_F64, NNK_ERR_ARG, "bad dtype");
  NNK_REQUIRE(a->cost_kind == 0 || a->cost_kind == 1, NNK_ERR_UNSUPPORTED, "cost_kind must be 0 (euclid) or 1 (melcd)");
  NNK_REQUIRE(a->path_ld >= a->max_tx + a->max_ty - 1, NNK_ERR_ARG, "path_ld < max_tx + max_ty - 1");
  NNK_REQUIRE(a->radius != 0, NNK_ERR_UNSUPPORTED, "radius 0 is not supported (fastdtw itself fails on odd lengths)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool full = a->radius < 0;
  const int mt = a->max_tx > a->max_ty ? a->max_tx : a->max_ty;
  DtwParams p;
  p.X = a->X; p.Y = a->Y; p.is_f64 = a->dtype == NNK_F64; p.n_pairs = a->n_pairs;
  p.x_pair_stride = a->x_pair_stride; p.y_pair_stride = a->y_pair_stride; p.x_ld = a->x_ld; p.y_ld = a->y_ld; p.D = a->D;
  p.len_x = a->len_x; p.len_y = a->len_y; p.order = a->order; p.cost_kind = a->cost_kind; p.radius = a->radius;
  p.path_i = a->path_i; p.path_j = a->path_j; p.path_ld = a->path_ld; p.path_len = a->path_len; p.dist = a->dist;
  p.cells = (long long*)a->cells; p.max_tx = a->max_tx; p.max_ty = a->max_ty;
  p.ws = (unsigned char*)a->workspace;
  p.series_doubles = dtw_series_doubles(mt, a->D);
  const size_t need = nnk_dtw_workspace_bytes(a->n_pairs, a->max_tx, a->max_ty, a->D, a->radius);
  NNK_REQUIRE(a->workspace_bytes >= need, NNK_ERR_WORKSPACE, "DTW workspace too small");
  p.ws_pair_bytes = need / (size_t)a->n_pairs;
  p.bp_bytes = p.ws_pair_bytes - 2 * p.series_doubles * sizeof(double);
  p.logdb = 10.0 / log(10.0) * sqrt(2.0);  // metrics/__init__.py:5
  int dev = 0, max_smem = 0;
  NNK_CUDA_CHECK(cudaGetDevice(&dev));
  NNK_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (full) {
    constexpr int BLOCK = 256;
    const size_t smem = dtw_smem_bytes(a->max_tx, true, 0);
    NNK_REQUIRE(smem <= (size_t)max_smem, NNK_ERR_UNSUPPORTED, "sequence too long for the wavefront buffers in shared memory");
    p.smem_bp_cap = 0;
    NNK_CUDA_CHECK(cudaFuncSetAttribute(dtw_kernel<BLOCK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dtw_kernel<BLOCK, true><<<a->n_pairs, BLOCK, smem, st>>>(p);
  } else {
    constexpr int BLOCK = 32;
    size_t bound = dtw_fast_cells_bound(a->max_tx, a->max_ty, a->radius);
    size_t smem = dtw_smem_bytes(a->max_tx, false, (int)bound);
    if (smem > (size_t)max_smem / 2) {  // keep >= 2 CTAs per SM; overflow back-pointers go to global scratch
      const size_t base = dtw_smem_bytes(a->max_tx, false, 0);
      NNK_REQUIRE(base + 1024 <= (size_t)max_smem, NNK_ERR_UNSUPPORTED, "sequence too long for shared memory");
      bound = ((size_t)max_smem / 2 > base + 1024) ? (size_t)max_smem / 2 - base : 1024;
      smem = dtw_smem_bytes(a->max_tx, false, (int)bound);
    }
    p.smem_bp_cap = (int)bound;
    NNK_CUDA_CHECK(cudaFuncSetAttribute(dtw_kernel<BLOCK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dtw_kernel<BLOCK, false><<<a->n_pairs, BLOCK, smem, st>>>(p);
  }
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

extern "C" int nnk_gather_rows(const void* X, int32_t dtype, int64_t x_pair_stride, int32_t x_ld, const int32_t* path,
                               int32_t path_ld, const int32_t* path_len, void* out, int64_t out_pair_stride,
                               int32_t out_rows, int32_t D, int32_t n_pairs, void* stream) {
  NNK_REQUIRE(X && path && path_len && out, NNK_ERR_ARG, "NULL pointer");
  if (n_pairs == 0 || out_rows == 0 || D == 0) return NNK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)(((int64_t)out_rows * D + 255) / 256), (unsigned)n_pairs);
  if (grid.x > 1024) grid.x = 1024;
  if (dtype == NNK_F32)
    gather_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)X, x_pair_stride, x_ld, path, path_ld, path_len, (float*)out, out_pair_stride, out_rows, D);
  else
    gather_rows_kernel<double><<<grid, 256, 0, st>>>((const double*)X, x_pair_stride, x_ld, path, path_ld, path_len, (double*)out, out_pair_stride, out_rows, D);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

extern "C" int nnk_trim_lengths(const void* X, int32_t dtype, int64_t pair_stride, int32_t ld, int32_t T, int32_t D,
                                double eps, int32_t n_pairs, int32_t* len, void* stream) {
  NNK_REQUIRE(X && len, NNK_ERR_ARG, "NULL pointer");
  if (n_pairs == 0) return NNK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NNK_F32) trim_len_kernel<float><<<n_pairs, 128, 0, st>>>((const float*)X, pair_stride, ld, T, D, (float)eps, len);
  else trim_len_kernel<double><<<n_pairs, 128, 0, st>>>((const double*)X, pair_stride, ld, T, D, eps, len);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}
