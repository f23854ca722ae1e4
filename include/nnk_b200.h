/*
 * nnk_b200.h -- C ABI of libnnk_b200.so: the B200 (sm_100a) implementation of nnmnkwii's two
 * numeric hot paths, MLPG trajectory smoothing and DTW alignment.
 *
 * The reference (r9y9/nnmnkwii v0.1.3) has no FFI/plugin registry: its boundary is a set of Python
 * callables backed by Cython extensions.  Each entry point below names the reference interface
 * (file:line under /root/reference) whose arithmetic it replaces; INTEGRATION.md shows the ctypes
 * stub a reference maintainer would add at each of those call sites.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - "device" pointers are CUDA device memory owned by the caller; "host" pointers are ordinary
 *     (ideally pinned) host memory.  All matrices are row-major.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Device entry points are asynchronous: they enqueue work on `stream` and return.  Numerical
 *     failures (non-positive pivot) are written to a caller-provided device status record that the
 *     host inspects after synchronising (nnk_status_t).  Host entry points synchronise internally.
 *   - Return value: NNK_OK or a negative NNK_ERR_* (argument / CUDA errors; nnk_last_error() has text).
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef NNK_B200_H
#define NNK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNK_ABI_VERSION 2

#define NNK_OK 0
#define NNK_ERR_ARG -1          /* bad argument                                               */
#define NNK_ERR_UNSUPPORTED -2  /* window set larger than NNK_MAX_WIN / NNK_MAX_HALF           */
#define NNK_ERR_CUDA -3         /* CUDA runtime error                                          */
#define NNK_ERR_WORKSPACE -4    /* workspace too small (see *_workspace_bytes)                 */
#define NNK_ERR_NOT_PD -5       /* host entry points only: non-positive pivot (see status)     */

#define NNK_F32 0
#define NNK_F64 1

#define NNK_MAX_WIN 4  /* windows per stream (static, delta, delta-delta, +1)                  */
#define NNK_MAX_HALF 4 /* max(l, u) of any window                                              */
#define NNK_MAX_TAPS (2 * NNK_MAX_HALF + 1)

/* Windows = the reference's list of (l, u, win_coeff) triples (paramgen/_mlpg.py:16-20).
 * coef[w][0 .. l[w]+u[w]] holds win_coeff of window w.                                        */
typedef struct nnk_windows {
  int32_t nw;
  int32_t l[NNK_MAX_WIN];
  int32_t u[NNK_MAX_WIN];
  double coef[NNK_MAX_WIN][NNK_MAX_TAPS];
} nnk_windows_t;

/* Failure record.  On the device it is ONE zero-initialised uint64 word (`status_word`) that the
 * kernels update atomically so that the lexicographically first failure (utterance, chain, frame)
 * wins -- the order in which the reference's Python loops would have raised.  The host decodes
 * it with nnk_status_decode().
 *   code 0 = ok, 1 = non-positive pivot (reference: scipy.linalg.LinAlgError
 *   "%d-th leading minor not positive definite", _bandmat/linalg.pyx:79-82)                    */
typedef struct nnk_status {
  int32_t code;
  int32_t utt;   /* utterance index                                                           */
  int32_t chain; /* chain index (static dimension within the layout)                          */
  int32_t frame; /* 1-based frame of the pivot, as in the reference's message                 */
} nnk_status_t;

/* One "chain" = one static dimension of one stream = one banded T x T solve.
 * A (T, D) frame matrix in Merlin layout holds several streams side by side (e.g. mgc 180 = 3 x 60,
 * lf0 3, vuv 1, bap 3); chain c reads window w of its stream at column in_col + w * win_stride and
 * writes its trajectory to column out_col.  flags & 1 = pass-through (copy in_col -> out_col, no
 * smoothing: the vuv column of the gallery notebooks).                                          */
typedef struct nnk_chain {
  int32_t in_col;
  int32_t win_stride;
  int32_t out_col;
  int32_t flags;
} nnk_chain_t;

/* Batched MLPG over a flat (n_rows, ld) frame matrix holding n_utt utterances back to back.    */
typedef struct nnk_mlpg_args {
  const void* means;          /* device (n_rows, in_ld)   dtype                                */
  const void* vars;           /* device (n_rows, var_ld) per-frame variances, or (>= D,) global
                                 variances when var_ld == 0 (paramgen/_mlpg.py:169-170)        */
  const void* grad_out;       /* device (n_rows, go_ld): nnk_mlpg_grad: backpropagated gradient;
                                 nnk_mlpg_solve: right-hand sides.  Chain c reads column c.
                                 float32, or float64 when go_f64 != 0                          */
  void* out;                  /* device: fwd (n_rows, out_ld) dtype ; grad (n_rows, out_ld) f32 */
  int32_t dtype;              /* NNK_F32 / NNK_F64 of means, vars (and out for fwd)            */
  int32_t n_utt;
  int64_t in_ld, var_ld, go_ld, out_ld; /* row strides in elements                            */
  const int64_t* utt_off;     /* device (n_utt + 1) row offsets                                */
  const int32_t* utt_len;     /* device (n_utt) frame counts, or NULL => utt_off[u+1]-utt_off[u];
                                 lets zero-padded (B, Tmax, D) batches be used in place        */
  const int32_t* order;       /* device (n_utt) processing order (e.g. longest first) or NULL  */
  const nnk_chain_t* chains;  /* device (n_chain)                                              */
  int32_t n_chain;
  int32_t max_T;              /* max utterance length (host knows it; sizes the workspace)     */
  int32_t go_f64;             /* grad_out / rhs element type: 0 = float32, 1 = float64         */
  nnk_windows_t win;
  void* workspace;            /* device scratch, >= nnk_mlpg_workspace_bytes()                 */
  size_t workspace_bytes;
  uint64_t* status_word;      /* device; must be zeroed by the caller before the first launch  */
  const int64_t* out_off;     /* device (n_utt) first OUTPUT row of every utterance, or NULL =>
                                 utt_off (same rows in and out).  Lets a rank write its slice of
                                 a sharded batch straight into its slot of the all-gather buffer */
} nnk_mlpg_args_t;

void nnk_status_decode(uint64_t status_word, nnk_status_t* out);

/* Replaces paramgen.mlpg (paramgen/_mlpg.py:92-199: build_poe :53-89, bla.solveh
 * _bandmat/linalg.pyx:290-304) for a whole batch: out[:, chain.out_col] = P^{-1} b per chain.   */
int nnk_mlpg_fwd(const nnk_mlpg_args_t* args, void* stream);

/* Replaces paramgen.mlpg_grad (paramgen/_mlpg.py:202-281) in closed form:
 * out[:, in_col + w*win_stride] = tau_w * (W_w P^{-1} grad_out[:, chain]); out is float32.      */
int nnk_mlpg_grad(const nnk_mlpg_args_t* args, void* stream);

/* General banded solve with the same P: out[:, chain.out_col] = P^{-1} rhs[:, chain] (dtype of
 * `out` = args->dtype).  Used to build unit_variance_mlpg_matrix (paramgen/_mlpg.py:297-373):
 * R = P^{-1} Wtilde^T with unit variances, one chain per column of Wtilde^T.                    */
int nnk_mlpg_solve(const nnk_mlpg_args_t* args, void* stream);

/* Scratch needed by the calls above for `n_utt` utterances of at most max_T frames.         */
size_t nnk_mlpg_workspace_bytes(int32_t n_utt, int32_t n_chain, int32_t max_T, const nnk_windows_t* win);

/* Host-buffer convenience = what a cgo/ctypes binding of paramgen.mlpg would call: one utterance,
 * one stream, host pointers (means (T, D), variances (T, D) or (D,), out (T, D / nw)), copies
 * included, synchronous.  Returns NNK_ERR_NOT_PD with *bad_frame = 1-based frame on failure.    */
int nnk_mlpg_host(const void* means, const void* vars, int32_t var_is_1d, int32_t dtype, int64_t T,
                  int64_t D, const nnk_windows_t* win, void* out, int32_t* bad_frame);

/* Host-buffer batched MLPG: flat (n_rows, D) host matrices, chains/offsets on the host.  H2D and
 * D2H copies run chunked on two streams so they overlap the solve.  This is the end-to-end path
 * bench.py times as `e2e`.                                                                      */
int nnk_mlpg_batch_host(const void* means, const void* vars, int32_t var_is_1d, int32_t dtype,
                        int64_t n_rows, int64_t D, int64_t D_out, const int64_t* utt_off, int32_t n_utt,
                        const nnk_chain_t* chains, int32_t n_chain, const nnk_windows_t* win, void* out,
                        nnk_status_t* status);

/* ---- UnitVarianceMLPG (autograd/_impl/mlpg.py:70-172) ------------------------------------------
 * R (T, nw*T) row-major device matrix (float32, or float64 when dtype == NNK_F64) as produced by
 * unit_variance_mlpg_matrix (paramgen/_mlpg.py:297-373).
 * 1. nnk_uv_band_profile: profile[dist] = max |R[t, w*T+s]| over |t-s| == dist (T floats); the
 *    host picks the half-width K from it.
 * 2. nnk_uv_band_extract: Rb, RbT (T, nw, 2K+1) band tables (same dtype as R).
 * 3. nnk_uv_apply: backward == 0: y (B, T, sd) = R x   replacing torch.matmul(R, reshaped_means)
 *    (mlpg.py:138); backward != 0: y = R^T x replacing torch.matmul(R.transpose(0,1), grad_output)
 *    (mlpg.py:158).  The nw-window side is (B, T, nw*sd) when reshaped == 0 or (B, nw*T, sd)
 *    when reshaped != 0 (mlpg.py:124-136, :160-167); the other side is (B, T, sd).               */
int nnk_uv_band_profile(const void* R, int32_t dtype, int32_t T, int32_t nw, float* profile, void* stream);
int nnk_uv_band_extract(const void* R, int32_t dtype, int32_t T, int32_t nw, int32_t K, void* Rb, void* RbT, void* stream);
int nnk_uv_apply(const void* table, const void* x, void* y, int32_t dtype, int32_t B, int32_t T, int32_t sd,
                 int32_t nw, int32_t K, int32_t backward, int32_t reshaped, void* stream);
/* float32 variant exploiting that away from the two ends every row of a window block of R is the
 * same FIR filter: rows [t_lo, t_hi) use `taps` (HOST pointer, nw x (2K+1) floats, passed to the
 * kernel through the constant bank), the edge rows use `table` as above.                          */
int nnk_uv_apply_toeplitz(const void* table, const float* taps, const void* x, void* y, int32_t B, int32_t T,
                          int32_t sd, int32_t nw, int32_t K, int32_t t_lo, int32_t t_hi, int32_t backward,
                          int32_t reshaped, void* stream);

/* Factored float32 variant: in the shift-invariant rows every window block of R is one long filter h0
 * (a row of P^-1) convolved with a short window stencil, h_w = h0 * c_w; the host recovers c_w from R
 * (HOST pointers: h0 2K+1 floats, c nw x (2*KC+1) floats) and the sweep costs (2K+1) + ~7 multiply-adds
 * per output instead of nw*(2K+1) (replaces the same two torch.matmul calls, mlpg.py:138 / :158).
 * Packs two static dims per thread for any static_dim (odd ones, e.g. the reference's perf grid
 * static_dim = 59, perf/autograd_mlpg_perf.py:110-120, use predicated 4-byte accesses).              */
int nnk_uv_apply_factored(const void* table, const float* h0, const float* c, const void* x, void* y, int32_t B,
                          int32_t T, int32_t sd, int32_t nw, int32_t K, int32_t KC, int32_t t_lo, int32_t t_hi,
                          int32_t backward, int32_t reshaped, void* stream);

/* ---- DTW alignment (preprocessing/alignment.py:9-190) ------------------------------------------
 * Batched replacement of `dist, path = fastdtw(x, y, radius=self.radius, dist=self.dist)`
 * (alignment.py:50, :138; third-party slaypni/fastdtw, unpinned in setup.py:139 -- see DESIGN.md
 * "parity unpinned").  One CTA per pair; radius < 0 = exact DTW (full anti-diagonal wavefront),
 * radius >= 1 = FastDTW with that radius.  cost_kind 0 = default lambda x, y: norm(x - y)
 * (alignment.py:35), 1 = metrics.melcd (metrics/__init__.py:27-57).  Arithmetic is float64.     */
typedef struct nnk_dtw_args {
  const void* X;              /* device (n_pairs, .., D): pair p, frame t at X + p*x_pair_stride + t*x_ld */
  const void* Y;
  int32_t dtype;              /* NNK_F32 / NNK_F64 of X and Y                                   */
  int32_t n_pairs;
  int64_t x_pair_stride, y_pair_stride; /* elements                                           */
  int32_t x_ld, y_ld, D;
  const int32_t* len_x;       /* device (n_pairs): frames kept by trim_zeros_frames (alignment.py:49) */
  const int32_t* len_y;
  const int32_t* order;       /* device (n_pairs) processing order or NULL                      */
  int32_t cost_kind;
  int32_t radius;
  int32_t* path_i;            /* device (n_pairs, path_ld): pathx of alignment.py:52             */
  int32_t* path_j;            /* device (n_pairs, path_ld): pathy                                */
  int32_t path_ld;            /* >= max_tx + max_ty - 1                                         */
  int32_t* path_len;          /* device (n_pairs)                                               */
  double* dist;               /* device (n_pairs): accumulated cost D[Tx-1, Ty-1]               */
  int64_t* cells;             /* device (n_pairs) DP cells evaluated, or NULL                   */
  int32_t max_tx, max_ty;
  void* workspace;
  size_t workspace_bytes;     /* >= nnk_dtw_workspace_bytes()                                   */
} nnk_dtw_args_t;

int nnk_dtw_align(const nnk_dtw_args_t* args, void* stream);
size_t nnk_dtw_workspace_bytes(int32_t n_pairs, int32_t max_tx, int32_t max_ty, int32_t D, int32_t radius);

/* out[p, r, :] = X[p, path[p, r], :] for r < path_len[p], zero for r >= path_len[p]
 * (x = x[pathx]; X_aligned[idx][:len(x)] = x over a zero array: alignment.py:46-47, 52-54, 72-73)  */
int nnk_gather_rows(const void* X, int32_t dtype, int64_t x_pair_stride, int32_t x_ld, const int32_t* path,
                    int32_t path_ld, const int32_t* path_len, void* out, int64_t out_pair_stride, int32_t out_rows,
                    int32_t D, int32_t n_pairs, void* stream);

/* len[p] = len(trim_zeros_frames(X[p], eps, trim="b")) (preprocessing/generic.py:291-323)          */
int nnk_trim_lengths(const void* X, int32_t dtype, int64_t pair_stride, int32_t ld, int32_t T, int32_t D, double eps,
                     int32_t n_pairs, int32_t* len, void* stream);

/* ---- delta features (preprocessing/generic.py:229-288) ------------------
 * out[:, w*D + d] = np.correlate(x[:, d], coef_w, mode="same") per utterance of a flat (sum_T, D)
 * batch: window centred at len(coef_w) // 2, zeros outside the utterance, float64 arithmetic,
 * result in the dtype of x.  out has nw*D columns.                                                */
int nnk_delta_features(const void* x, int32_t dtype, int32_t D, int64_t x_ld, const int64_t* utt_off,
                       const int32_t* utt_len, int32_t n_utt, int32_t max_T, const nnk_windows_t* win, void* out,
                       int64_t out_ld, void* stream);

/* ---- length-masked objective metrics (metrics/__init__.py:27-190) --------
 * Padded (B, T, D) batches X, Y (element strides item_stride / frame_stride, D contiguous), lengths
 * (B) int32 on the device or NULL (all T frames valid).  The kernels deliver the SUM (float64) and the
 * COUNT of contributing frames; mean / sqrt / dB constant are the caller's scalar finish.
 *   nnk_frame_metric kind 0: sum of per-frame ||x - y||_2      (melcd,               :59-71)
 *                    kind 1: sum of (x - y)^2 over frames x D  (mean_squared_error,  :103-110)
 *   nnk_f0_metric    kind 0: sum of (x - y)^2 over frames with src_vuv + tgt_vuv >= 2; count = voiced
 *                    kind 1: the same on exp(x), exp(y)        (lf0_mean_squared_error, :141-165)
 *                    kind 2: sum of (src_vuv != tgt_vuv)       (vuv_error,           :181-190)
 * Deterministic (fixed-order fold of per-block partials).  workspace >= nnk_metric_workspace_bytes(B, T),
 * zero-filled before its first use (each call leaves it reusable); one workspace per stream.         */
int64_t nnk_metric_workspace_bytes(int32_t B, int32_t T);
int nnk_frame_metric(const void* X, const void* Y, int32_t dtype, int32_t B, int32_t T, int32_t D,
                     int64_t item_stride, int64_t frame_stride, const int32_t* lengths, int32_t kind, double* sum_out,
                     int64_t* count_out, void* workspace, int64_t workspace_bytes, void* stream);
int nnk_f0_metric(const void* src_f0, const void* src_vuv, const void* tgt_f0, const void* tgt_vuv, int32_t dtype,
                  int32_t B, int32_t T, int64_t item_stride, int64_t frame_stride, const int32_t* lengths,
                  int32_t kind, double* sum_out, int64_t* count_out, void* workspace, int64_t workspace_bytes,
                  void* stream);

/* ---- GMM mapping in front of MLPG (baseline/gmm.py:47-247) ---------------------
 * Device tables of a joint source/target GMM with M mixtures over D-dimensional frames, float64:
 *   src_means, tgt_means (M, D); prec_chol (M, D, D) = sklearn precisions_cholesky_ of the source marginal
 *   (U_m, row-major [d][e]); log_const (M) = log w_m + log det U_m - D/2 log 2pi;
 *   A_t (M, D, D) = (Syx_m Sxx_m^-1)^T, row-major [j][i]; Dm (M, D) = Eq. 23 diagonal variances (may be
 *   NULL when no variances are requested).
 * nnk_gmm_logprob: lp (T, M) = log w_m + log N(x_t | mu_m, Sxx_m)  -- the per-frame predict_proba /
 *   posterior of gmm.py:116-118, 219-221 before normalisation.
 * nnk_gmm_map: mode 0 = MLPG.transform's arg-max mixture sequence (gmm.py:219-237): E[t] = Eq. 22 mean,
 *   Dv[t] = Eq. 23 variance of the chosen mixture, mix[t] = its index (Dv / mix may be NULL);
 *   mode 1 = MLPGBase._transform_frame (gmm.py:97-121): E[t] = posterior-weighted mean, Eq. 13.     */
typedef struct nnk_gmm {
  const double* src_means;
  const double* tgt_means;
  const double* prec_chol;
  const double* log_const;
  const double* A_t;
  const double* Dm;
  int32_t M, D;
} nnk_gmm_t;
int nnk_gmm_logprob(const nnk_gmm_t* gmm, const double* x, int64_t x_ld, int32_t T, double* lp, void* stream);
int nnk_gmm_map(const nnk_gmm_t* gmm, const double* x, int64_t x_ld, int32_t T, const double* lp, int32_t mode, double* E,
                double* Dv, int32_t* mix, void* stream);

/* ---- sharded batches (the reference has no multi-device path) ------------------------
 * Copies n_seg row segments (whole utterances) between two row-major device matrices:
 * dst[dst_row[s] + r, 0:cols] = src[src_row[s] + r, 0:cols] for r < len[s].  Used to bring the
 * all-gathered trajectories (shard order: bucket, rank, utterance) back into the caller's utterance
 * order, i.e. the order the reference's per-utterance loop over paramgen.mlpg
 * (paramgen/_mlpg.py:92) would have produced them in.  elem_bytes 4 or 8; n_seg <= 65535 per call. */
int nnk_segment_copy(const void* src, void* dst, int32_t elem_bytes, int64_t cols, int64_t src_ld, int64_t dst_ld,
                     const int64_t* src_row, const int64_t* dst_row, const int32_t* len, int32_t n_seg,
                     int32_t max_len, void* stream);

/* Peer-memory transport of a sharded result (one process per GPU of one NVLink / NVSwitch box):
 * a cudaMalloc allocation per rank, shared through CUDA IPC handles; nnk_peer_copy pushes a byte range
 * into a peer's mapping with copy-engine DMA over NVLink (no SMs: it overlaps the solve kernels, which
 * an NCCL all-gather kernel cannot while they hold every SM slot).  The 64-byte handle travels between
 * the processes by any host channel (torch.distributed here).                                       */
int nnk_peer_alloc(size_t bytes, void** ptr);
int nnk_peer_free(void* ptr);
int nnk_peer_export(const void* ptr, unsigned char* handle64);
int nnk_peer_open(const unsigned char* handle64, void** ptr);
int nnk_peer_close(void* ptr);
int nnk_peer_copy(void* dst_peer, const void* src_local, size_t bytes, void* stream);

const char* nnk_last_error(void);
int nnk_abi_version(void);
/* Number of kernel launches this library has issued since load (bench.py's gpu_launches).       */
int64_t nnk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NNK_B200_H */
