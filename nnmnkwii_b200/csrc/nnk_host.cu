// nnk_host.cu -- host-buffer entry points of libnnk_b200: the calls a ctypes/cgo binding of the
// reference's Python functions would make (host pointers in, host pointers out, copies included).
//
//   nnk_mlpg_host        <-> paramgen.mlpg(mean_frames, variance_frames, windows)  (_mlpg.py:92)
//   nnk_mlpg_batch_host  <-> the per-utterance / per-stream loop of the gallery notebooks around it
//
// Device memory comes from a grow-only per-process arena (cudaMalloc is milliseconds; the solve is
// tens of microseconds).  The batch call splits the utterances into row-balanced chunks and runs
// H2D(i+1) / solve(i) / D2H(i-1) concurrently on three streams, so end-to-end time approaches the
// PCIe transfer time of the inputs.
#include <string.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <vector>

#include "nnk_common.cuh"

namespace nnk {

struct Arena {
  void* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return NNK_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + (1 << 20);
    NNK_CUDA_CHECK(cudaMalloc(&ptr, want));
    cap = want;
    return NNK_OK;
  }
};

struct HostCtx {
  std::mutex mu;
  Arena in_m, in_v, out, ws, meta;
  cudaStream_t s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev;
  int init() {
    if (s_run) return NNK_OK;
    NNK_CUDA_CHECK(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    NNK_CUDA_CHECK(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    NNK_CUDA_CHECK(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
    return NNK_OK;
  }
  int events(size_t n) {
    while (ev.size() < n) {
      cudaEvent_t e;
      NNK_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    return NNK_OK;
  }
};
// one context (arena + streams) per device: the streams and allocations of a context belong to the
// device that was current when it was first used
static constexpr int kMaxDevices = 64;
static HostCtx g_ctx_by_device[kMaxDevices];

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace nnk

using namespace nnk;

extern "C" int nnk_mlpg_batch_host(const void* means, const void* vars, int32_t var_is_1d, int32_t dtype,
                                   int64_t n_rows, int64_t D, int64_t D_out, const int64_t* utt_off,
                                   int32_t n_utt, const nnk_chain_t* chains, int32_t n_chain,
                                   const nnk_windows_t* win, void* out, nnk_status_t* status) {
  NNK_REQUIRE(means && vars && utt_off && chains && win && out, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE(dtype == NNK_F32 || dtype == NNK_F64, NNK_ERR_ARG, "bad dtype");
  NNK_REQUIRE(n_rows >= 0 && D > 0 && D_out > 0 && n_utt >= 0 && n_chain >= 0, NNK_ERR_ARG, "bad size");
  if (status) { status->code = 0; status->utt = status->chain = status->frame = 0; }
  if (n_utt == 0 || n_rows == 0 || n_chain == 0) return NNK_OK;
  NNK_REQUIRE(utt_off[0] == 0 && utt_off[n_utt] == n_rows, NNK_ERR_ARG, "utt_off must span [0, n_rows]");
  const size_t es = dtype == NNK_F32 ? 4 : 8;

  int device = 0;
  NNK_CUDA_CHECK(cudaGetDevice(&device));
  NNK_REQUIRE(device >= 0 && device < kMaxDevices, NNK_ERR_ARG, "device index out of range");
  HostCtx& g_ctx = g_ctx_by_device[device];
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  int rc = g_ctx.init();
  if (rc) return rc;

  int max_T = 0;
  for (int u = 0; u < n_utt; ++u) {
    const int64_t T = utt_off[u + 1] - utt_off[u];
    NNK_REQUIRE(T >= 0 && T < (1 << 21) - 1, NNK_ERR_ARG, "utterance length out of range");
    max_T = std::max<int>(max_T, (int)T);
  }
  // chunks of whole utterances, balanced by rows; enough of them to overlap copies with the solve
  const int64_t in_bytes = n_rows * D * (int64_t)es;
  int n_chunk = (int)std::min<int64_t>(n_utt, std::max<int64_t>(1, std::min<int64_t>(16, in_bytes / (8 << 20))));
  std::vector<int> cb(n_chunk + 1, 0);
  {
    int u = 0;
    for (int c = 1; c < n_chunk; ++c) {
      const int64_t target = n_rows * c / n_chunk;
      while (u < n_utt && utt_off[u] < target) ++u;
      cb[c] = std::max(u, cb[c - 1]);
    }
    cb[n_chunk] = n_utt;
  }
  int max_chunk_utt = 0;
  for (int c = 0; c < n_chunk; ++c) max_chunk_utt = std::max(max_chunk_utt, cb[c + 1] - cb[c]);

  // per-chunk longest-first order (entries are global utterance ids)
  std::vector<int32_t> order(n_utt);
  std::iota(order.begin(), order.end(), 0);
  for (int c = 0; c < n_chunk; ++c)
    std::stable_sort(order.begin() + cb[c], order.begin() + cb[c + 1], [&](int a, int b) {
      return (utt_off[a + 1] - utt_off[a]) > (utt_off[b + 1] - utt_off[b]);
    });

  const size_t ws_bytes = nnk_mlpg_workspace_bytes(max_chunk_utt, n_chain, max_T, win);
  NNK_REQUIRE(ws_bytes > 0, NNK_ERR_UNSUPPORTED, "unsupported window set");
  const size_t ws_cap = std::min<size_t>(ws_bytes, (size_t)4 << 30);
  if ((rc = g_ctx.in_m.ensure((size_t)n_rows * D * es))) return rc;
  if ((rc = g_ctx.in_v.ensure(var_is_1d ? (size_t)D * es : (size_t)n_rows * D * es))) return rc;
  if ((rc = g_ctx.out.ensure((size_t)n_rows * D_out * es))) return rc;
  if ((rc = g_ctx.ws.ensure(ws_cap))) return rc;
  const size_t off_b = align_up(sizeof(int64_t) * (n_utt + 1), 256);
  const size_t ord_b = align_up(sizeof(int32_t) * n_utt, 256);
  const size_t chn_b = align_up(sizeof(nnk_chain_t) * n_chain, 256);
  if ((rc = g_ctx.meta.ensure(off_b + ord_b + chn_b + 256))) return rc;
  if ((rc = g_ctx.events(2 * (size_t)n_chunk))) return rc;
  char* meta = (char*)g_ctx.meta.ptr;
  int64_t* d_off = (int64_t*)meta;
  int32_t* d_ord = (int32_t*)(meta + off_b);
  nnk_chain_t* d_chn = (nnk_chain_t*)(meta + off_b + ord_b);
  uint64_t* d_status = (uint64_t*)(meta + off_b + ord_b + chn_b);

  cudaStream_t sh = g_ctx.s_h2d, sr = g_ctx.s_run, sd = g_ctx.s_d2h;
  NNK_CUDA_CHECK(cudaMemcpyAsync(d_off, utt_off, sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, sh));
  NNK_CUDA_CHECK(cudaMemcpyAsync(d_ord, order.data(), sizeof(int32_t) * n_utt, cudaMemcpyHostToDevice, sh));
  NNK_CUDA_CHECK(cudaMemcpyAsync(d_chn, chains, sizeof(nnk_chain_t) * n_chain, cudaMemcpyHostToDevice, sh));
  NNK_CUDA_CHECK(cudaMemsetAsync(d_status, 0, sizeof(uint64_t), sh));
  if (var_is_1d) NNK_CUDA_CHECK(cudaMemcpyAsync(g_ctx.in_v.ptr, vars, (size_t)D * es, cudaMemcpyHostToDevice, sh));
  // columns no chain writes get a defined (zero) value
  NNK_CUDA_CHECK(cudaMemsetAsync(g_ctx.out.ptr, 0, (size_t)n_rows * D_out * es, sh));

  for (int c = 0; c < n_chunk; ++c) {
    const int64_t r0 = utt_off[cb[c]], r1 = utt_off[cb[c + 1]];
    const size_t boff = (size_t)r0 * D * es, bytes = (size_t)(r1 - r0) * D * es;
    if (bytes) {
      NNK_CUDA_CHECK(cudaMemcpyAsync((char*)g_ctx.in_m.ptr + boff, (const char*)means + boff, bytes, cudaMemcpyHostToDevice, sh));
      if (!var_is_1d)
        NNK_CUDA_CHECK(cudaMemcpyAsync((char*)g_ctx.in_v.ptr + boff, (const char*)vars + boff, bytes, cudaMemcpyHostToDevice, sh));
    }
    NNK_CUDA_CHECK(cudaEventRecord(g_ctx.ev[2 * c], sh));
    NNK_CUDA_CHECK(cudaStreamWaitEvent(sr, g_ctx.ev[2 * c], 0));
    nnk_mlpg_args_t a;
    memset(&a, 0, sizeof(a));
    a.means = g_ctx.in_m.ptr; a.vars = g_ctx.in_v.ptr; a.out = g_ctx.out.ptr;
    a.dtype = dtype; a.n_utt = cb[c + 1] - cb[c];
    a.in_ld = D; a.var_ld = var_is_1d ? 0 : D; a.out_ld = D_out;
    a.utt_off = d_off; a.order = d_ord + cb[c]; a.chains = d_chn; a.n_chain = n_chain; a.max_T = max_T;
    a.win = *win; a.workspace = g_ctx.ws.ptr; a.workspace_bytes = ws_cap; a.status_word = d_status;
    if (a.n_utt > 0 && (rc = nnk_mlpg_fwd(&a, sr))) return rc;
    NNK_CUDA_CHECK(cudaEventRecord(g_ctx.ev[2 * c + 1], sr));
    NNK_CUDA_CHECK(cudaStreamWaitEvent(sd, g_ctx.ev[2 * c + 1], 0));
    const size_t ooff = (size_t)r0 * D_out * es, obytes = (size_t)(r1 - r0) * D_out * es;
    if (obytes)
      NNK_CUDA_CHECK(cudaMemcpyAsync((char*)out + ooff, (const char*)g_ctx.out.ptr + ooff, obytes, cudaMemcpyDeviceToHost, sd));
  }
  uint64_t word = 0;
  NNK_CUDA_CHECK(cudaStreamSynchronize(sr));
  NNK_CUDA_CHECK(cudaMemcpyAsync(&word, d_status, sizeof(word), cudaMemcpyDeviceToHost, sd));
  NNK_CUDA_CHECK(cudaStreamSynchronize(sd));
  NNK_CUDA_CHECK(cudaStreamSynchronize(sh));
  nnk_status_t st;
  decode_status(word, &st);
  if (status) *status = st;
  if (st.code) {
    set_error("%d-th leading minor not positive definite (utterance %d, chain %d)", st.frame, st.utt, st.chain);
    return NNK_ERR_NOT_PD;
  }
  return NNK_OK;
}

extern "C" int nnk_mlpg_host(const void* means, const void* vars, int32_t var_is_1d, int32_t dtype, int64_t T,
                             int64_t D, const nnk_windows_t* win, void* out, int32_t* bad_frame) {
  NNK_REQUIRE(win && win->nw >= 1 && win->nw <= NNK_MAX_WIN, NNK_ERR_UNSUPPORTED, "unsupported number of windows");
  if (bad_frame) *bad_frame = 0;
  const int64_t sd = D / win->nw;  // static_dim = D // num_windows (paramgen/_mlpg.py:172)
  if (T == 0 || sd == 0) return NNK_OK;
  std::vector<nnk_chain_t> chains((size_t)sd);
  for (int64_t d = 0; d < sd; ++d) {
    chains[d].in_col = (int32_t)d;
    chains[d].win_stride = (int32_t)sd;
    chains[d].out_col = (int32_t)d;
    chains[d].flags = 0;
  }
  int64_t off[2] = {0, T};
  nnk_status_t st;
  int rc = nnk_mlpg_batch_host(means, vars, var_is_1d, dtype, T, D, sd, off, 1, chains.data(), (int32_t)sd, win, out, &st);
  if (rc == NNK_ERR_NOT_PD && bad_frame) *bad_frame = st.frame;
  return rc;
}
