// dgemm_tma_proto.cu -- prototype of the covariance GEMM with TMA-fed stages (cp.async.bulk.tensor + mbarrier) instead of
// per-thread cp.async, to be compared with k_dgemm_nt on the two shapes of config A:
//   GEMM-1  C[4096 x 2016] = A[4096 x 336] . B[2016 x 336]^T        (K = 336)
//   GEMM-2  C[4096 x 336]  = A[4096 x 2000] . B[336 x 2000]^T       (K = 2000, split-K 3)
// Stages are dense 128-byte rows (16 doubles) written by the TMA engine with the hardware 128B swizzle; lane (fr, fk) of a DMMA
// fragment owns k = 4 fk .. 4 fk + 3 of every slab (a permutation of k common to both operands), so its operands are two LDS.128
// per row and slab, conflict-free under the swizzle (chunk = (2 fk + h) ^ fr).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dgemm_tma_proto dgemm_tma_proto.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BM = 64, BK = 16, NTHREADS = 128, MT = 4;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst), "l"(map),
               "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2) k_dgemm_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                                                            double* __restrict__ C, int ldc, size_t split_stride) {
  constexpr int WTN = BN / 2, NTL = WTN / 8;
  constexpr unsigned A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  const unsigned base = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;  // 128B swizzle: 1024-byte aligned stages
  const unsigned bars = base + STAGES * STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1, fr = lane >> 2, fk = lane & 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int KT_all = (K + BK - 1) / BK;
  const int kt_beg = (int)((long long)KT_all * blockIdx.z / gridDim.z), kt_end = (int)((long long)KT_all * (blockIdx.z + 1) / gridDim.z);
  const int KT = kt_end - kt_beg;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int kt) {  // thread 0: slab kt -> stage kt % STAGES
    const int s = kt % STAGES;
    const unsigned bar = bars + 8 * s, sa = base + s * STAGE_BYTES;
    mbar_expect_tx(bar, STAGE_BYTES);
    tma_load_2d(sa, &tmA, (kt_beg + kt) * BK, m0, bar);
    tma_load_2d(sa + A_BYTES, &tmB, (kt_beg + kt) * BK, n0, bar);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES - 1 && s < KT; s++) issue(s);

  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // this lane's fragment addresses inside a stage: row * 128 + (((2 fk + h) ^ fr) << 4)
  const unsigned offA = (wm * 32 + fr) * 128, offB = A_BYTES + (wn * WTN + fr) * 128;
  const unsigned ch0 = ((2 * fk) ^ fr) << 4, ch1 = ((2 * fk + 1) ^ fr) << 4;
  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % STAGES;
    mbar_wait(bars + 8 * s, (kt / STAGES) & 1);
    __syncthreads();  // everybody has finished slab kt-1: its stage may be refilled
    if (threadIdx.x == 0 && kt + STAGES - 1 < KT) issue(kt + STAGES - 1);
    const unsigned st = base + s * STAGE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const unsigned ch = h ? ch1 : ch0;
      double a0[MT], a1[MT], b0[NTL], b1[NTL];
#pragma unroll
      for (int i = 0; i < MT; i++) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a0[i]), "=d"(a1[i]) : "r"(st + offA + i * 8 * 128 + ch));
#pragma unroll
      for (int j = 0; j < NTL; j++) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(b0[j]), "=d"(b1[j]) : "r"(st + offB + j * 8 * 128 + ch));
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
    }
  }
  double* out = C + (size_t)blockIdx.z * split_stride;
#pragma unroll
  for (int i = 0; i < MT; i++) {
    const int row = m0 + wm * 32 + i * 8 + fr;
#pragma unroll
    for (int j = 0; j < NTL; j++) {
      const int col = n0 + wn * WTN + j * 8 + 2 * fk;
      *reinterpret_cast<double2*>(out + (size_t)row * ldc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// Persistent variant: one CTA per resident slot, tiles taken with stride gridDim.x; the slab sequence runs across tile boundaries, so
// the first slabs of the next tile are already in flight while the epilogue of the current one stores its results.
template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2) k_dgemm_tma_persist(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                    int K, int tiles_n, int tiles_m, int ksplit, double* __restrict__ C, int ldc,
                                                                    size_t split_stride) {
  constexpr int WTN = BN / 2, NTL = WTN / 8;
  constexpr unsigned A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  const unsigned base = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
  const unsigned bars = base + STAGES * STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1, fr = lane >> 2, fk = lane & 3;
  const int KT_all = (K + BK - 1) / BK, total = tiles_n * tiles_m * ksplit;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  auto tile_coords = [&](int t, int& m0, int& n0, int& z, int& kb, int& kn) {
    const int nt = t % tiles_n, r = t / tiles_n;
    z = r / tiles_m;
    m0 = (r % tiles_m) * BM;
    n0 = nt * BN;
    kb = (int)((long long)KT_all * z / ksplit);
    kn = (int)((long long)KT_all * (z + 1) / ksplit) - kb;
  };
  // producer cursor (thread 0 only)
  int p_tile = blockIdx.x, p_kt = 0, p_q = 0, p_m0 = 0, p_n0 = 0, p_z = 0, p_kb = 0, p_kn = 0;
  if (p_tile < total) tile_coords(p_tile, p_m0, p_n0, p_z, p_kb, p_kn);
  auto produce = [&]() {
    if (p_tile >= total) return;
    const int s = p_q % STAGES;
    const unsigned bar = bars + 8 * s, sa = base + s * STAGE_BYTES;
    mbar_expect_tx(bar, STAGE_BYTES);
    tma_load_2d(sa, &tmA, (p_kb + p_kt) * BK, p_m0, bar);
    tma_load_2d(sa + A_BYTES, &tmB, (p_kb + p_kt) * BK, p_n0, bar);
    p_q++;
    if (++p_kt == p_kn) {
      p_kt = 0;
      p_tile += gridDim.x;
      if (p_tile < total) tile_coords(p_tile, p_m0, p_n0, p_z, p_kb, p_kn);
    }
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES - 1; s++) produce();

  const unsigned offA = (wm * 32 + fr) * 128, offB = A_BYTES + (wn * WTN + fr) * 128;
  const unsigned ch0 = ((2 * fk) ^ fr) << 4, ch1 = ((2 * fk + 1) ^ fr) << 4;
  int q = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int m0, n0, z, kb, KT;
    tile_coords(tile, m0, n0, z, kb, KT);
    double acc[MT][NTL][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < KT; kt++, q++) {
      const int s = q % STAGES;
      mbar_wait(bars + 8 * s, (q / STAGES) & 1);
      __syncthreads();
      if (threadIdx.x == 0) produce();
      const unsigned st = base + s * STAGE_BYTES;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const unsigned ch = h ? ch1 : ch0;
        double a0[MT], a1[MT], b0[NTL], b1[NTL];
#pragma unroll
        for (int i = 0; i < MT; i++) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a0[i]), "=d"(a1[i]) : "r"(st + offA + i * 8 * 128 + ch));
#pragma unroll
        for (int j = 0; j < NTL; j++) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(b0[j]), "=d"(b1[j]) : "r"(st + offB + j * 8 * 128 + ch));
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
      }
    }
    double* out = C + (size_t)z * split_stride;
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const int row = m0 + wm * 32 + i * 8 + fr;
#pragma unroll
      for (int j = 0; j < NTL; j++) {
        const int col = n0 + wn * WTN + j * 8 + 2 * fk;
        *reinterpret_cast<double2*>(out + (size_t)row * ldc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

// reference: one thread per output element
__global__ void k_ref(const double* A, int lda, const double* B, int ldb, int M, int N, int K, double* C, int ldc) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  double t = 0.0;
  for (int k = 0; k < K; k++) t += A[(size_t)m * lda + k] * B[(size_t)n * ldb + k];
  C[(size_t)m * ldc + n] = t;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, const double* ptr, int rows, int K, int ld, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

template <int BN, int STAGES>
void run_case(EncodeFn enc, const char* name, int M, int N, int K, int lda, int ldb, int ksplit) {
  // A [M][lda], B [N_pad][ldb], C [ksplit][M][N_pad]
  const int N_pad = (N + BN - 1) / BN * BN;
  std::vector<double> hA((size_t)M * lda, 0.0), hB((size_t)N_pad * ldb, 0.0);
  srand(1);
  for (int m = 0; m < M; m++)
    for (int k = 0; k < K; k++) hA[(size_t)m * lda + k] = rand() / (double)RAND_MAX - 0.5;
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) hB[(size_t)n * ldb + k] = rand() / (double)RAND_MAX - 0.5;
  double *A, *B, *C, *R;
  cudaMalloc(&A, hA.size() * 8); cudaMalloc(&B, hB.size() * 8);
  cudaMalloc(&C, (size_t)ksplit * M * N_pad * 8); cudaMalloc(&R, (size_t)M * N_pad * 8);
  cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(B, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(C, 0, (size_t)ksplit * M * N_pad * 8);
  k_ref<<<dim3((N_pad + 127) / 128, M), 128>>>(A, lda, B, ldb, M, N_pad, K, R, N_pad);
  CUtensorMap tmA = make_map(enc, A, M, K, lda, BM), tmB = make_map(enc, B, N_pad, K, ldb, BN);
  const size_t smem = (size_t)STAGES * (BM + BN) * 128 + 8 * STAGES + 1024;
  cudaFuncSetAttribute(k_dgemm_tma<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(N_pad / BN, M / BM, ksplit);
  k_dgemm_tma<BN, STAGES><<<grid, NTHREADS, smem>>>(tmA, tmB, K, C, N_pad, (size_t)M * N_pad);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<double> hC((size_t)ksplit * M * N_pad), hR((size_t)M * N_pad);
  cudaMemcpy(hC.data(), C, hC.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hR.data(), R, hR.size() * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0.0;
  for (size_t t = 0; t < hR.size(); t++) {
    double v = 0.0;
    for (int z = 0; z < ksplit; z++) v += hC[(size_t)z * M * N_pad + t];
    maxerr = fmax(maxerr, fabs(v - hR[t]));
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_dgemm_tma<BN, STAGES><<<grid, NTHREADS, smem>>>(tmA, tmB, K, C, N_pad, (size_t)M * N_pad);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-44s BN=%d stages=%d grid=%dx%dx%d  max|err|=%.2e  %.1f us  %.2f TFLOP/s (padded flops)\n", name, BN, STAGES, grid.x, grid.y, grid.z, maxerr,
         best * 1e3, 2.0 * M * N_pad * ((K + 15) / 16 * 16) / best / 1e9);
  {  // persistent variant
    cudaMemset(C, 0, (size_t)ksplit * M * N_pad * 8);
    cudaFuncSetAttribute(k_dgemm_tma_persist<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles_n = N_pad / BN, tiles_m = M / BM, total = tiles_n * tiles_m * ksplit, pg = total < 296 ? total : 296;
    k_dgemm_tma_persist<BN, STAGES><<<pg, NTHREADS, smem>>>(tmA, tmB, K, tiles_n, tiles_m, ksplit, C, N_pad, (size_t)M * N_pad);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: persistent kernel failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(hC.data(), C, hC.size() * 8, cudaMemcpyDeviceToHost);
    double perr = 0.0;
    for (size_t t = 0; t < hR.size(); t++) {
      double v = 0.0;
      for (int z = 0; z < ksplit; z++) v += hC[(size_t)z * M * N_pad + t];
      perr = fmax(perr, fabs(v - hR[t]));
    }
    float pbest = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      k_dgemm_tma_persist<BN, STAGES><<<pg, NTHREADS, smem>>>(tmA, tmB, K, tiles_n, tiles_m, ksplit, C, N_pad, (size_t)M * N_pad);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < pbest) pbest = ms;
    }
    printf("%-44s   persistent, %d CTAs:                   max|err|=%.2e  %.1f us  %.2f TFLOP/s\n", "", pg, perr, pbest * 1e3,
           2.0 * M * N_pad * ((K + 15) / 16 * 16) / pbest / 1e9);
  }
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(R);
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  run_case<112, 3>(enc, "GEMM-1 shape (4096 x 2000 x 325)", 4096, 2000, 325, 336, 336, 1);
  run_case<112, 4>(enc, "GEMM-1 shape (4096 x 2000 x 325)", 4096, 2000, 325, 336, 336, 1);
  run_case<112, 3>(enc, "GEMM-2 shape (4096 x 325 x 2000), split-K 3", 4096, 325, 2000, 2144, 2144, 3);
  run_case<112, 4>(enc, "GEMM-2 shape (4096 x 325 x 2000), split-K 3", 4096, 325, 2000, 2144, 2144, 3);
  run_case<112, 4>(enc, "large (32768 x 2016 x 336)", 32768, 2016, 336, 336, 336, 1);
  return 0;
}
