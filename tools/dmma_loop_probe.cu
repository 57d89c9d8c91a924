// dmma_loop_probe.cu -- where do the last 8-14 % of the covariance GEMM's main loop go?  Strips k_dgemm_nt's inner loop
// (64 x 112 x 16 CTA tile, 4 warps as 2 x 2, 4 x 7 DMMA tiles per warp, two CTAs per SM) down in steps:
//   V0  DMMAs only, operands fixed in registers (the fp64_pipes ceiling, with this kernel's 56 accumulators)
//   V1  + fragment loads from shared memory every k-step (same addresses as the GEMM, padded rows of 20 doubles)
//   V2  + __syncthreads() every slab (4 k-steps)
//   V3  + cp.async of the next slab from global memory (L2-resident operands) and cp.async.wait_group, i.e. the full main loop
//   V4  = V3 with the slab's copies issued in four parts between the k-steps instead of one burst after the barrier
// Prints TFLOP/s of each.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop_probe dmma_loop_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int BN = 112, BK = 16, LDS_ROW = BK + 4, MT = 4, NTL = 7, WTM = 32, WTN = 56;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int V, int STAGES, int BM, int CA>
__global__ void __launch_bounds__(BM * 2, BM == 64 ? 2 : 1) k_probe(const double* __restrict__ A, const double* __restrict__ B, int lda, int slabs, double* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = (double*)smem_raw;
  double* Bs = As + (size_t)STAGES * BM * LDS_ROW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1, fr = lane >> 2, fk = lane & 3;
  for (int k = threadIdx.x; k < STAGES * (BM + BN) * LDS_ROW; k += BM * 2) As[k] = 1.0 + 1e-9 * k;
  __syncthreads();
  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  double af[MT], bf[NTL];
#pragma unroll
  for (int i = 0; i < MT; i++) af[i] = 1.0 + i * 1e-9 + lane * 1e-12;
#pragma unroll
  for (int j = 0; j < NTL; j++) bf[j] = 1.0 - j * 1e-9 - lane * 1e-12;
  // copy addresses as in load_stage
  constexpr int RPP = BM * 2 / 8;  // rows per copy pass
  const int lrow = threadIdx.x / 8, lcc = (threadIdx.x % 8) * 2;
  const double* gA = A + (size_t)(blockIdx.x % (4096 / BM) * BM + lrow) * lda + lcc;
  const double* gB = B + (size_t)(blockIdx.x % 16 * BN + lrow) * lda + lcc;
  const unsigned sA0 = (unsigned)__cvta_generic_to_shared(As + lrow * LDS_ROW + lcc), sB0 = (unsigned)__cvta_generic_to_shared(Bs + lrow * LDS_ROW + lcc);
  // one quarter (part = 0..3) of a slab's copies: issued between the k-steps so that the fragment loads never queue behind a burst
  auto load_part = [&](int s, int kn, int part) {
    constexpr int NA = BM / RPP, NB = (BN + RPP - 1) / RPP, NT = NA + NB;
#pragma unroll
    for (int c = 0; c < NT; c++) {
      if (c * 4 / NT != part) continue;
      if (c < NA) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sA0 + (s * BM + c * RPP) * LDS_ROW * 8), "l"(gA + (size_t)c * RPP * lda + (kn % 20) * BK));
      } else {
        const int it = c - NA;
        if ((it + 1) * RPP <= BN || it * RPP + lrow < BN)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sB0 + (s * BN + it * RPP) * LDS_ROW * 8), "l"(gB + (size_t)it * RPP * lda + (kn % 20) * BK));
      }
    }
  };
  auto load_stage = [&](int s, int kn) {
#pragma unroll
    for (int it = 0; it < BM / RPP; it++) {
      if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sA0 + (s * BM + it * RPP) * LDS_ROW * 8), "l"(gA + (size_t)it * RPP * lda + (kn % 20) * BK));
      else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sA0 + (s * BM + it * RPP) * LDS_ROW * 8), "l"(gA + (size_t)it * RPP * lda + (kn % 20) * BK));
    }
#pragma unroll
    for (int it = 0; it < (BN + RPP - 1) / RPP; it++)
      if ((it + 1) * RPP <= BN || it * RPP + lrow < BN) {
        if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sB0 + (s * BN + it * RPP) * LDS_ROW * 8), "l"(gB + (size_t)it * RPP * lda + (kn % 20) * BK));
        else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sB0 + (s * BN + it * RPP) * LDS_ROW * 8), "l"(gB + (size_t)it * RPP * lda + (kn % 20) * BK));
      }
  };
  if (V >= 3) {
    for (int s = 0; s < STAGES - 1; s++) { load_stage(s, s); asm volatile("cp.async.commit_group;\n" ::); }
  }
  for (int kt = 0; kt < slabs; kt++) {
    if (V >= 3) asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 2));
    if (V >= 2) __syncthreads();
    if (V == 3) { load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1); asm volatile("cp.async.commit_group;\n" ::); }
    const double* as = As + (size_t)(kt % STAGES) * BM * LDS_ROW + (wm * WTM + fr) * LDS_ROW + fk;
    const double* bs = Bs + (size_t)(kt % STAGES) * BN * LDS_ROW + (wn * WTN + fr) * LDS_ROW + fk;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      if (V >= 1) {
#pragma unroll
        for (int i = 0; i < MT; i++) af[i] = as[i * 8 * LDS_ROW + kk * 4];
#pragma unroll
        for (int j = 0; j < NTL; j++) bf[j] = bs[j * 8 * LDS_ROW + kk * 4];
      }
      if (V == 4) {
        load_part((kt + STAGES - 1) % STAGES, kt + STAGES - 1, kk);
        if (kk == BK / 4 - 1) asm volatile("cp.async.commit_group;\n" ::);
      }
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  if (V >= 3) asm volatile("cp.async.wait_group 0;\n" ::);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int V, int STAGES = 3, int BM = 64, int CA = 0>
void run(const char* name, const double* A, const double* B, int lda, double* out) {
  const int slabs = 2000, ctas = BM == 64 ? 296 : 148;
  const size_t smem = (size_t)STAGES * (BM + BN) * LDS_ROW * sizeof(double);
  cudaFuncSetAttribute(k_probe<V, STAGES, BM, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_probe<V, STAGES, BM, CA><<<ctas, BM * 2, smem>>>(A, B, lda, 10, out);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_probe<V, STAGES, BM, CA><<<ctas, BM * 2, smem>>>(A, B, lda, slabs, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = (double)ctas * slabs * 2.0 * BM * BN * BK;
  printf("%-64s %.3f ms  %.2f TFLOP/s  (%s)\n", name, best, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int lda = 336, rows = 4096 + 2048;
  double *A, *out;
  cudaMalloc(&A, sizeof(double) * (size_t)rows * lda);
  cudaMemset(A, 0, sizeof(double) * (size_t)rows * lda);
  cudaMalloc(&out, 8);
  const double* B = A + (size_t)4096 * lda;
  run<0>("V0 DMMA only, register operands", A, B, lda, out);
  run<1>("V1 + fragment loads from shared memory", A, B, lda, out);
  run<2>("V2 + __syncthreads per slab", A, B, lda, out);
  run<3>("V3 + cp.async next slab / wait_group (full main loop)", A, B, lda, out);
  run<4>("V4 = V3 with the copies spread over the 4 k-steps", A, B, lda, out);
  run<4, 4>("V4 with 4 stages", A, B, lda, out);
  run<3, 4>("V3 with 4 stages", A, B, lda, out);
  run<3, 3, 64, 1>("V3 with cp.async.ca (L1-allocating)", A, B, lda, out);
  run<2, 3, 128>("V2 on a 128 x 112 tile, 8 warps, one CTA per SM", A, B, lda, out);
  run<3, 3, 128>("V3 on a 128 x 112 tile, 8 warps, one CTA per SM", A, B, lda, out);
  run<3, 4, 128>("V3 on a 128 x 112 tile, 4 stages", A, B, lda, out);
  return 0;
}
