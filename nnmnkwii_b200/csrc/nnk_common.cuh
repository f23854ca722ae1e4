// nnk_common.cuh -- shared host/device helpers for libnnk_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nnk_b200.h"

namespace nnk {

// ---- host-side error plumbing -------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define NNK_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::nnk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return NNK_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define NNK_REQUIRE(cond, code, msg)                    \
  do {                                                  \
    if (!(cond)) {                                      \
      ::nnk::set_error("%s:%d: %s", __FILE__, __LINE__, msg); \
      return code;                                      \
    }                                                   \
  } while (0)

// ---- device guard ---------------------------------------------------------------------------
// Every launching entry point runs on the device that OWNS the data it was given, not on whatever
// device happens to be current in the calling thread (torch's default stream handle 0 is valid on
// every device, so a launch with cuda:1 pointers while cuda:0 is current would otherwise run on the
// wrong GPU or fault).  The previous device is restored on return.
struct DeviceGuard {
  int prev = -1, dev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* device_ptr) {
    cudaPointerAttributes at;
    if (device_ptr && cudaPointerGetAttributes(&at, device_ptr) == cudaSuccess &&
        (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
      dev = at.device;
      if (cudaGetDevice(&prev) == cudaSuccess && prev != dev && cudaSetDevice(dev) == cudaSuccess) switched = true;
    } else {
      cudaGetLastError();  // not a device pointer: leave the current device alone
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ---- status key ------------------------------------------------------------------------------
// The device keeps ONE 64-bit word: 0 = ok, otherwise ~((utt << 42) | (chain << 21) | frame) of the
// lexicographically FIRST failure (utterance, then chain, then frame) -- the order in which the
// reference's Python loops would have raised (paramgen/_mlpg.py:184 over d, linalg.pyx:77 over
// frames).  atomicMax on the complement == atomicMin on the key.
__device__ __forceinline__ void report_not_pd(unsigned long long* status, int utt, int chain, int frame1) {
  unsigned long long key = ((unsigned long long)(unsigned)utt << 42) | ((unsigned long long)(unsigned)chain << 21) |
                           (unsigned long long)(unsigned)frame1;
  atomicMax(status, ~key);
}

static inline void decode_status(unsigned long long word, nnk_status_t* st) {
  if (word == 0ull) { st->code = 0; st->utt = st->chain = st->frame = 0; return; }
  unsigned long long key = ~word;
  st->code = 1;
  st->utt = (int32_t)(key >> 42);
  st->chain = (int32_t)((key >> 21) & 0x1FFFFF);
  st->frame = (int32_t)(key & 0x1FFFFF);
}

// ---- small device helpers ----------------------------------------------------------------------
template <typename T> struct recip_in_dtype;
// 1 / variance is evaluated in the INPUT dtype and then widened, exactly like
// `precisions[:, w] = 1 / variance_frames[:, col]` (paramgen/_mlpg.py:188): IEEE division.
template <> struct recip_in_dtype<float> {
  static __device__ __forceinline__ double f(float v) { return (double)__frcp_rn(v); }
};
template <> struct recip_in_dtype<double> {
  static __device__ __forceinline__ double f(double v) { return __drcp_rn(v); }
};

// streaming loads/stores: inputs are read once -> do not pollute L1, evict first from L2
template <typename T> __device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <typename T> __device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }

// predicated streaming store (no branch: idle lanes simply do not store)
__device__ __forceinline__ void st_stream_if(float* p, float v, bool pred) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q st.global.cs.f32 [%0], %1;\n}" ::"l"(p), "f"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void st_stream_if(double* p, double v, bool pred) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q st.global.cs.f64 [%0], %1;\n}" ::"l"(p), "d"(v), "r"((int)pred) : "memory");
}

}  // namespace nnk
