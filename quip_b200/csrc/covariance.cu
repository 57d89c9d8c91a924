// covariance.cu -- sparse-GP dot-product covariance and its gradient back-contraction on FP64 tensor cores.
//
// Replaces the per-atom BLAS-2 pair of gpCoordinates_Predict (src/GAP/gp_predict.f95:3753-3775, 3854-3859):
//     c = sparseX^T x           (dgemv 'T', once per atom)
//     k = delta^2 c^zeta cutoff ;  E_i = k . alpha ;  a = alpha delta^2 zeta c^(zeta-1) cutoff   (fast_pow_1d :3581)
//     gradPredict = sparseX a   (dgemv 'N', once per atom)
// by two batched GEMMs over all centres at once:
//     GEMM-1  C[Nc x M] = X[Nc x d] . S^T   with the kernel non-linearity, the alpha weighting and the row
//             reduction to E_i fused into the epilogue (C itself never reaches HBM; only a = dE_i/dc does),
//     GEMM-2  G[Nc x d] = A[Nc x M] . S.
// The reference streams sparseX twice per atom; here sparseX is read once per 128 atoms and stays L2 resident.
//
// Both GEMMs are the same "NT" kernel (A[m][k], B[n][k], K contiguous in both), built on the FP64 tensor-core
// instruction mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- tcgen05 has no FP64 kind), fed by a 3-stage
// cp.async pipeline.  CTA tile 128x128x16, 8 warps as 2(M) x 4(N), warp tile 64x32 = 8x4 DMMA tiles,
// 64 FP64 accumulators per thread.  Shared-memory rows are padded to 20 doubles so that the (8 rows x 4 k)
// fragment loads of a half-warp hit 16 distinct 8-byte banks.
#include <type_traits>

#include "gap_device.cuh"

namespace gapb200 {

namespace {

constexpr int BM = COV_BM, BN = COV_BN, BK = COV_BK;
constexpr int WARPS_M = 2, WARPS_N = 4, NTHREADS = WARPS_M * WARPS_N * 32;
constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;  // 64 x 32
constexpr int MT = WTM / 8, NTL = WTN / 8;             // 8 x 4 DMMA tiles per warp
constexpr int LDS_ROW = BK + 4;                        // padded row (doubles)
constexpr int STAGES = 3;
constexpr size_t GEMM_SMEM = (size_t)STAGES * (BM + BN) * LDS_ROW * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// load one K-slab (BK columns) of the A and B tiles into stage `st`
__device__ __forceinline__ void load_stage(double* As, double* Bs, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                           int m0, int n0, int k0) {
  // BM*BK/2 16-byte chunks for A, same for B: 1024 + 1024 chunks / 256 threads = 4 + 4 each
#pragma unroll
  for (int it = 0; it < (BM * BK / 2) / NTHREADS; it++) {
    int chunk = threadIdx.x + it * NTHREADS;
    int row = chunk / (BK / 2), cc = (chunk % (BK / 2)) * 2;
    cp_async16(As + row * LDS_ROW + cc, A + (size_t)(m0 + row) * lda + k0 + cc);
  }
#pragma unroll
  for (int it = 0; it < (BN * BK / 2) / NTHREADS; it++) {
    int chunk = threadIdx.x + it * NTHREADS;
    int row = chunk / (BK / 2), cc = (chunk % (BK / 2)) * 2;
    cp_async16(Bs + row * LDS_ROW + cc, B + (size_t)(n0 + row) * ldb + k0 + cc);
  }
}

struct EpiCov {  // GEMM-1 epilogue
  const double* alpha;
  const double* cutoff;
  CovParams cp;
  double* acoef;
  int lda_out;
  double* epart;
  int n_tiles_n;
  int M;  // real number of sparse points; columns >= M are padding
};
struct EpiStore {  // GEMM-2 epilogue
  double* out;
  int ldo;
};

__device__ __forceinline__ double ipow(double v, int e) {  // fast_pow_1d: v**e_int by repeated multiplication
  double r = 1.0;
  for (int i = 0; i < e; i++) r *= v;
  return r;
}

template <class Epi>
__global__ void __launch_bounds__(NTHREADS, 1) k_dgemm_nt(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, int K,
                                                          Epi epi) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = (double*)smem_raw;                  // [STAGES][BM][LDS_ROW]
  double* Bs = As + (size_t)STAGES * BM * LDS_ROW;  // [STAGES][BN][LDS_ROW]
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int fr = lane >> 2, fk = lane & 3;  // fragment row (0..7) and k (0..3)

  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = K / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) load_stage(As + (size_t)s * BM * LDS_ROW, Bs + (size_t)s * BN * LDS_ROW, A, lda, B, ldb, m0, n0, s * BK);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {  // prefetch slab kt+STAGES-1 into the stage that was consumed in iteration kt-1
      int kn = kt + STAGES - 1;
      if (kn < KT) {
        int s = kn % STAGES;
        load_stage(As + (size_t)s * BM * LDS_ROW, Bs + (size_t)s * BN * LDS_ROW, A, lda, B, ldb, m0, n0, kn * BK);
      }
      cp_async_commit();
    }
    const double* as = As + (size_t)(kt % STAGES) * BM * LDS_ROW + (wm * WTM + fr) * LDS_ROW + fk;
    const double* bs = Bs + (size_t)(kt % STAGES) * BN * LDS_ROW + (wn * WTN + fr) * LDS_ROW + fk;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      double af[MT], bf[NTL];
#pragma unroll
      for (int i = 0; i < MT; i++) af[i] = as[i * 8 * LDS_ROW + kk * 4];
#pragma unroll
      for (int j = 0; j < NTL; j++) bf[j] = bs[j * 8 * LDS_ROW + kk * 4];
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: thread holds C[row = fr][col = 2*fk, 2*fk+1] of every 8x8 tile ----
  if constexpr (std::is_same<Epi, EpiCov>::value) {
    const EpiCov& e = *reinterpret_cast<const EpiCov*>(&epi);
    double* red = (double*)smem_raw;  // [WARPS_N][BM]
    double al[NTL][2], cu[NTL][2];
#pragma unroll
    for (int j = 0; j < NTL; j++)
#pragma unroll
      for (int t = 0; t < 2; t++) {
        int col = n0 + wn * WTN + j * 8 + 2 * fk + t;
        al[j][t] = e.alpha[col];
        cu[j][t] = e.cutoff[col];
      }
#pragma unroll
    for (int i = 0; i < MT; i++) {
      int row = m0 + wm * WTM + i * 8 + fr;
      double esum = 0.0;
#pragma unroll
      for (int j = 0; j < NTL; j++) {
        double outv[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
          double c = acc[i][j][t];
          if (n0 + wn * WTN + j * 8 + 2 * fk + t >= e.M) { outv[t] = 0.0; continue; }
          double pw = e.cp.zeta_int >= 1 ? ipow(c, e.cp.zeta_int - 1) : (e.cp.zeta_int == 0 ? 0.0 : pow(c, e.cp.zeta - 1.0));
          double kval = e.cp.zeta_int == 0 ? e.cp.delta2 : e.cp.delta2 * (pw * c);
          esum += al[j][t] * (kval * cu[j][t]);
          outv[t] = al[j][t] * e.cp.delta2 * e.cp.zeta * pw * cu[j][t];
        }
        int col = n0 + wn * WTN + j * 8 + 2 * fk;
        *reinterpret_cast<double2*>(e.acoef + (size_t)row * e.lda_out + col) = make_double2(outv[0], outv[1]);
      }
      esum += __shfl_xor_sync(0xffffffffu, esum, 1);
      esum += __shfl_xor_sync(0xffffffffu, esum, 2);
      if (fk == 0) red[wn * BM + wm * WTM + i * 8 + fr] = esum;
    }
    __syncthreads();
    if (threadIdx.x < BM) {
      double t = (red[threadIdx.x] + red[BM + threadIdx.x]) + (red[2 * BM + threadIdx.x] + red[3 * BM + threadIdx.x]);
      e.epart[(size_t)(m0 + threadIdx.x) * e.n_tiles_n + blockIdx.x] = t;
    }
  } else {
    const EpiStore& e = *reinterpret_cast<const EpiStore*>(&epi);
#pragma unroll
    for (int i = 0; i < MT; i++) {
      int row = m0 + wm * WTM + i * 8 + fr;
#pragma unroll
      for (int j = 0; j < NTL; j++) {
        int col = n0 + wn * WTN + j * 8 + 2 * fk;
        *reinterpret_cast<double2*>(e.out + (size_t)row * e.ldo + col) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

// E_i = sum over column tiles (fixed order) ; local_e(centre) += E_i  (IPModel_GAP.f95:454-459 with cc = 1, |ci| = 1)
__global__ void k_energy_rows(const double* __restrict__ epart, int n_tiles_n, const int* __restrict__ centres, int n_centres, double e_scale,
                              double* __restrict__ local_e) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_centres) return;
  double t = 0.0;
  for (int k = 0; k < n_tiles_n; k++) t += epart[(size_t)c * n_tiles_n + k];
  local_e[centres[c]] += e_scale * t;
}

}  // namespace

void launch_cov_gemm1(const double* x, int ldx, const double* sp_rows, int lds, int n_rows_pad, int M, int M_pad, int K_pad, const double* alpha,
                      const double* cutoff, CovParams cp, double* acoef, int lda, double* epart, int n_tiles_n, cudaStream_t st,
                      int* launches) {
  EpiCov e{alpha, cutoff, cp, acoef, lda, epart, n_tiles_n, M};
  cudaFuncSetAttribute(k_dgemm_nt<EpiCov>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
  dim3 grid(M_pad / BN, n_rows_pad / BM);
  k_dgemm_nt<EpiCov><<<grid, NTHREADS, GEMM_SMEM, st>>>(x, ldx, sp_rows, lds, K_pad, e);
  *launches += 1;
}

void launch_cov_gemm2(const double* acoef, int lda, const double* st_rows, int ldst, int n_rows_pad, int dn_pad, int K_pad, double* gvec,
                      int ldg, cudaStream_t st, int* launches) {
  EpiStore e{gvec, ldg};
  cudaFuncSetAttribute(k_dgemm_nt<EpiStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
  dim3 grid(dn_pad / BN, n_rows_pad / BM);
  k_dgemm_nt<EpiStore><<<grid, NTHREADS, GEMM_SMEM, st>>>(acoef, lda, st_rows, ldst, K_pad, e);
  *launches += 1;
}

void launch_energy_rows(const double* epart, int n_tiles_n, const int* centres, int n_centres, double e_scale, double* local_e,
                        cudaStream_t st, int* launches) {
  if (n_centres <= 0) return;
  k_energy_rows<<<(n_centres + 255) / 256, 256, 0, st>>>(epart, n_tiles_n, centres, n_centres, e_scale, local_e);
  *launches += 1;
}

}  // namespace gapb200
