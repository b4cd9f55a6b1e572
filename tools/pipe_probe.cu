// Micro-probe (not part of the library): do the FP64 FMA pipe (DFMA) and the FP64 tensor pipe (DMMA.8x8x4) of sm_100a
// run concurrently, or do they share issue/throughput?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// gpurun_out/pipe_probe tools/pipe_probe.cu ; run on the GPU box.  Prints TFLOP/s of DFMA only, DMMA only, both
// interleaved inside every warp, and both split over warps of the same CTA.
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// mode 0: DFMA only, 1: DMMA only, 2: interleaved in each warp, 3: even warps DFMA / odd warps DMMA
template <int NF, int NM>
__global__ void __launch_bounds__(256) probe(double* sink, int iters, int mode) {
  double f[NF];
  double c[NM][2];
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = 1e-9 * (threadIdx.x + i);
  const double m = 1.0 + 1e-12, q = 1e-13, a = 1e-3 * (threadIdx.x & 7), b = 1e-3;
  const bool do_f = mode == 0 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 0);
  const bool do_m = mode == 1 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 1);
  if (do_f && do_m) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        dmma(c[i], a, b);
#pragma unroll
        for (int j = 0; j < NF / NM; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i * (NF / NM) + j]) : "d"(m), "d"(q));
      }
    }
  } else if (do_f) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(m), "d"(q));
    }
  } else if (do_m) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma(c[i], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) sink[0] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink;
  cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  constexpr int NF = 64, NM = 8;  // per iteration and warp: 64 DFMA (64*64 flop) and 8 DMMA (8*512 flop): balanced
  const int iters = 20000, threads = 256;
  const char* names[] = {"DFMA only", "DMMA only", "interleaved per warp", "split over warps"};
  for (int cps = 1; cps <= 4; cps *= 2) {
    const int blocks = sms * cps;
    for (int mode = 0; mode < 4; ++mode) {
      probe<NF, NM><<<blocks, threads>>>(sink, iters / 10, mode);
      cudaEventRecord(e0);
      probe<NF, NM><<<blocks, threads>>>(sink, iters, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double warps = (double)blocks * threads / 32;
      double wf = (mode == 0 || mode == 2) ? warps : (mode == 3 ? warps / 2 : 0);
      double wm = (mode == 1 || mode == 2) ? warps : (mode == 3 ? warps / 2 : 0);
      const double ff = wf * iters * NF * 64.0, fm = wm * iters * NM * 512.0;
      printf("ctas/SM %d  %-22s  %.3f ms  DFMA %.2f TF/s  DMMA %.2f TF/s  total %.2f TF/s\n", cps, names[mode], ms,
             ff / ms / 1e9, fm / ms / 1e9, (ff + fm) / ms / 1e9);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
