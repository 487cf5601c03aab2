// Microbenchmark (B200): does an FP64 warp instruction (16 lanes per scheduler: 2 cycles on the pipe) also hold the scheduler's
// issue port for 2 cycles, or can integer / FP32 instructions issue in its shadow?  Per loop iteration every thread runs
// NF independent DFMA chains and NI independent integer (LOP3 / IADD) chains; with W warps per scheduler resident.
//   overlap:  cycles per iteration per scheduler ~ W * max(2 * NF, NF + NI)
//   blocking: cycles per iteration per scheduler ~ W * (2 * NF + NI)
// usage: fp64_issue            (prints a table)
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI>
__global__ void __launch_bounds__(256) k(double* out, int* iout, int iters, double a, int b) {
  double x[NF > 0 ? NF : 1];
  int y[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NI; ++i) y[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, 1e-9);
#pragma unroll
      for (int i = 0; i < NI; ++i) y[i] = (y[i] ^ b) + it;
    }
  }
  double s = 0;
  int t = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < NI; ++i) t += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int NF, int NI>
void run(double* out, int* iout, int ctas_per_sm, int sms, double ghz) {
  const int iters = 2000;
  k<NF, NI><<<sms * ctas_per_sm, 256>>>(out, iout, 10, 0.999, 5);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NF, NI><<<sms * ctas_per_sm, 256>>>(out, iout, iters, 0.999, 5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * ghz * 1e9;
  const int w = ctas_per_sm * 2;                       // warps per scheduler (256 threads = 8 warps = 2 per scheduler)
  const double per_iter = cycles / (iters * 4.0) / w;  // cycles per unrolled group per warp on one scheduler
  printf("NF=%2d NI=%2d warps/sched=%d  %8.3f ms  cycles per (NF dfma + NI int) per warp: %6.2f   overlap model %5.1f  blocking model %5.1f\n",
         NF, NI, w, ms, per_iter, (double)(2 * NF > NF + 2 * NI ? 2 * NF : NF + 2 * NI), (double)(2 * NF + 2 * NI));
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double ghz = khz * 1e-6;
  double* out; int* iout;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaMalloc(&iout, sizeof(int) * sms * 8 * 256);
  printf("SMs %d, clock %.3f GHz (each integer chain step is 2 instructions: LOP3 + IADD)\n", sms, ghz);
  for (int c : {1, 2, 4}) {
    run<8, 0>(out, iout, c, sms, ghz);
    run<8, 2>(out, iout, c, sms, ghz);
    run<8, 4>(out, iout, c, sms, ghz);
    run<8, 8>(out, iout, c, sms, ghz);
    run<0, 8>(out, iout, c, sms, ghz);
    run<4, 8>(out, iout, c, sms, ghz);
  }
  return 0;
}
