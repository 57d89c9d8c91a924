// fp64_pipes.cu -- micro-benchmark: are the FP64 FMA pipe (DFMA) and the FP64 tensor pipe (DMMA.8x8x4) of sm_100a
// independent issue targets?  Three kernels with register-only operands: DMMA only, DFMA only, and a mix in which
// every warp interleaves both.  Prints TFLOP/s of each; if mix > max(dmma, dfma) the covariance GEMM can put part of
// its tile on the FMA pipe.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NMMA, int NFMA>
__global__ void __launch_bounds__(128) k_mix(int iters, double* out) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double acc[NMMA > 0 ? NMMA : 1][2];
  double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < (NMMA > 0 ? NMMA : 1); i++) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); i++) f[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < NMMA; i++) dmma884(acc[i][0], acc[i][1], a, b);
#pragma unroll
      for (int i = 0; i < NFMA; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;\n" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < (NMMA > 0 ? NMMA : 1); i++) s += acc[i][0] + acc[i][1];
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); i++) s += f[i];
  if (s == 123.456) out[0] = s;
}

template <int NMMA, int NFMA>
void run(const char* name, int ctas_per_sm) {
  int iters = 4000, nsm = 148;
  double* out;
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_mix<NMMA, NFMA><<<nsm * ctas_per_sm, 128>>>(10, out);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_mix<NMMA, NFMA><<<nsm * ctas_per_sm, 128>>>(iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double warps = (double)nsm * ctas_per_sm * 4;
  double fl_mma = warps * iters * 4.0 * NMMA * 512.0, fl_fma = warps * iters * 4.0 * NFMA * 64.0;
  printf("%-28s ctas/sm=%d  %.3f ms  dmma %.2f TF  dfma %.2f TF  total %.2f TF\n", name, ctas_per_sm, best, fl_mma / best / 1e9, fl_fma / best / 1e9,
         (fl_mma + fl_fma) / best / 1e9);
  cudaFree(out);
}

int main() {
  for (int c = 1; c <= 4; c *= 2) {
    run<16, 0>("dmma only (16 acc)", c);
    run<0, 16>("dfma only (16 acc)", c);
    run<16, 16>("16 dmma + 16 dfma", c);
    run<16, 32>("16 dmma + 32 dfma", c);
    run<8, 32>("8 dmma + 32 dfma", c);
    run<8, 64>("8 dmma + 64 dfma", c);
    run<4, 64>("4 dmma + 64 dfma", c);
  }
  return 0;
}
