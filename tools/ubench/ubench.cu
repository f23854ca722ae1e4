// Instruction-throughput micro-benchmark for the pipes the MLPG assembler warps lean on (sm_100a):
// warp-instructions per clock per SM of F2F.F64.F32, MUFU.RCP, DFMA / DADD / DMUL, FFMA, and an integer
// float->double widening sequence.  usage: ./ubench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr int ITER = 4096;
constexpr int UNROLL = 8;

template <int OP>
__global__ void k(double* out, float seed) {
  float f[UNROLL];
  double d[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) { f[i] = seed + threadIdx.x * 1e-3f + i; d[i] = (double)f[i] * 1.0000001; }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (OP == 0) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); }
      if (OP == 1) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i])); }
      if (OP == 2) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
      if (OP == 3) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
      if (OP == 4) { asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
      if (OP == 5) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(1.0000001f)); }
      if (OP == 6) {  // integer widening of a normal float (4 ALU ops)
        uint32_t b = __float_as_uint(f[i]);
        uint32_t hi, lo;
        asm volatile("{\n.reg .u32 t;\nshr.u32 t, %2, 3;\nand.b32 t, t, 0x0fffffff;\nadd.u32 t, t, 0x38000000;\nlop3.b32 %0, t, %2, 0x80000000, 0xf8;\nshl.b32 %1, %2, 29;\n}" : "=r"(hi), "=r"(lo) : "r"(b));
        d[i] = __hiloint2double(hi, lo);
        f[i] = __uint_as_float(b + 1);
      }
      if (OP == 7) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); }
      if (OP == 8) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d[i]) : "d"(d[i])); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) s += d[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int warps_per_sm) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  const int threads = 32 * warps_per_sm;
  cudaMalloc(&out, sizeof(double) * sms * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms, threads>>>(out, 1.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP><<<sms, threads>>>(out, 1.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * khz * 1e3;
  const double winst = (double)ITER * UNROLL * warps_per_sm;  // warp instructions per SM
  printf("%-22s warps/SM=%2d  %8.3f ms  %7.3f warp-inst/clk/SM  (%.2f cycles per warp-inst per SMSP)\n", name, warps_per_sm, ms,
         winst / cycles, cycles / (winst / 4));
  cudaFree(out);
}

int main() {
  for (int w : {4, 16}) {
    run<0>("F2F.F64.F32", w);
    run<7>("F2F.F32.F64", w);
    run<1>("MUFU.RCP (f32)", w);
    run<8>("MUFU.RCP64H", w);
    run<2>("DFMA", w);
    run<3>("DADD", w);
    run<4>("DMUL", w);
    run<5>("FFMA", w);
    run<6>("int widen (5 ops)", w);
  }
  return 0;
}
