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
// cp.async pipeline.  CTA tile 64 x BN x 16 (BN = 128 or 112), 4 warps as 2(M) x 2(N), warp tile 32 x BN/2 =
// 4 x (8 or 7) DMMA tiles, TWO CTAs resident per SM so that one CTA's epilogue (the kernel non-linearity and the
// 64 x BN store) overlaps the other's main loop and the tail is balanced at half-tile granularity.  GEMM-2 can be
// split along K (= the sparse-point index) into `ksplit` partial outputs when it has too few tiles to fill 148 SMs;
// the consumer (the SOAP adjoint kernel) adds the partials in a fixed order.  Shared-memory rows are padded to
// 20 doubles so that the (8 rows x 4 k) fragment loads of a half-warp hit 16 distinct 8-byte banks.
#include <type_traits>

#include "gap_device.cuh"

namespace gapb200 {

namespace {

constexpr int BM = COV_BM, BK = COV_BK;
constexpr int WARPS_M = 2, WARPS_N = 2, NTHREADS = WARPS_M * WARPS_N * 32;
constexpr int WTM = BM / WARPS_M;   // 32
constexpr int MT = WTM / 8;         // 4 DMMA tiles along M per warp
constexpr int LDS_ROW = BK + 4;     // padded row (doubles)
constexpr int STAGES = 3;
template <int BN>
constexpr size_t gemm_smem() { return ((size_t)STAGES * (BM + BN) * LDS_ROW + BN) * sizeof(double); }  // + the tile's GP weights

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

// One K-slab (BK columns) of the A and B tiles -> a pipeline stage.  With BK/2 = 8 sixteen-byte chunks per row and 128
// threads, thread t always copies column chunk (t % 8) of rows t/8, t/8 + 16, t/8 + 32, ...: the global pointers and the
// shared addresses are computed ONCE per tile (gA/gB/sA/sB below) and only advance by constants afterwards -- the
// index arithmetic of the first version cost more issue slots than the DMMAs of a slab.
constexpr int CPR = BK / 2;                       // 16-byte chunks per row
constexpr int ROWS_PER_PASS = NTHREADS / CPR;     // 16
static_assert(BM % ROWS_PER_PASS == 0, "tile rows must be a multiple of the rows copied per pass");
template <int BN>
__device__ __forceinline__ void load_stage(unsigned sA, unsigned sB, const double* gA, const double* gB, size_t lda16, size_t ldb16) {
  static_assert(BN % ROWS_PER_PASS == 0, "tile rows must be a multiple of the rows copied per pass");
#pragma unroll
  for (int it = 0; it < BM / ROWS_PER_PASS; it++)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sA + it * ROWS_PER_PASS * LDS_ROW * 8), "l"(gA + it * lda16));
#pragma unroll
  for (int it = 0; it < BN / ROWS_PER_PASS; it++)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sB + it * ROWS_PER_PASS * LDS_ROW * 8), "l"(gB + it * ldb16));
}

// c^(zeta-1).  ZI = compile-time integer zeta (1..4, the usual GAP settings; fast_pow_1d multiplies repeatedly,
// gp_predict.f95:3581-3605); ZI = 0 is the general case, kept OUT of line: inlined into the fully unrolled epilogue
// (64 copies of the pow() slow path) it made the kernel instruction-cache bound.
__device__ __noinline__ double pow_zm1_general(double c, double zeta, int zeta_int) {
  if (zeta_int >= 1) {
    double r = 1.0;
    for (int i = 0; i < zeta_int - 1; i++) r *= c;
    return r;
  }
  return zeta_int == 0 ? 0.0 : pow(c, zeta - 1.0);
}
template <int ZI>
__device__ __forceinline__ double pow_zm1(double c, const CovParams& cp) {
  if (ZI == 1) return 1.0;
  if (ZI == 2) return c;
  if (ZI == 3) return c * c;
  if (ZI == 4) return (c * c) * c;
  return pow_zm1_general(c, cp.zeta, cp.zeta_int);
}

template <int ZI>
struct EpiCov {  // GEMM-1 epilogue
  const double* w;  // w_s = alpha_s * sparseCutoff_s * delta^2 (0 for padding columns), precombined at model upload
  CovParams cp;
  double* acoef;
  int lda_out;
  double* epart;
  int n_tiles_n;
  static constexpr int zi = ZI;
  __device__ __forceinline__ double pw(double c) const { return pow_zm1<ZI>(c, cp); }
};
struct EpiStore {  // GEMM-2 epilogue
  double* out;
  int ldo;
  size_t split_stride;  // doubles between the partial outputs of consecutive K splits
};

template <int BN, class Epi>
__global__ void __launch_bounds__(NTHREADS, 2) k_dgemm_nt(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, int K,
                                                          int row0, const int* __restrict__ n_rows_dev, Epi epi) {
  constexpr int WTN = BN / WARPS_N, NTL = WTN / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = (double*)smem_raw;                  // [STAGES][BM][LDS_ROW]
  double* Bs = As + (size_t)STAGES * BM * LDS_ROW;  // [STAGES][BN][LDS_ROW]
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (n_rows_dev && row0 + m0 >= *n_rows_dev) return;  // row tile beyond the (device-side) number of centres
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int fr = lane >> 2, fk = lane & 3;  // fragment row (0..7) and k (0..3)

  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // K range of this split (blockIdx.z of gridDim.z), in BK slabs.  K is a multiple of 4 (one DMMA k-step), not of BK:
  // the last slab is loaded whole (both operands are zero-padded to a multiple of BK) but only its first
  // nk_last k-steps are multiplied.
  const int KT_all = (K + BK - 1) / BK;
  const int nk_last = ((K - 1) % BK) / 4 + 1;
  const int kt_beg = (int)((long long)KT_all * blockIdx.z / gridDim.z), kt_end = (int)((long long)KT_all * (blockIdx.z + 1) / gridDim.z);
  const int KT = kt_end - kt_beg;
  // per-thread copy addresses (see load_stage)
  const int lrow = threadIdx.x / CPR, lcc = (threadIdx.x % CPR) * 2;
  const double* gA = A + (size_t)(m0 + lrow) * lda + (size_t)kt_beg * BK + lcc;
  const double* gB = B + (size_t)(n0 + lrow) * ldb + (size_t)kt_beg * BK + lcc;
  const size_t lda16 = (size_t)ROWS_PER_PASS * lda, ldb16 = (size_t)ROWS_PER_PASS * ldb;
  const unsigned sA0 = (unsigned)__cvta_generic_to_shared(As + lrow * LDS_ROW + lcc);
  const unsigned sB0 = (unsigned)__cvta_generic_to_shared(Bs + lrow * LDS_ROW + lcc);
  constexpr unsigned A_STAGE = BM * LDS_ROW * 8, B_STAGE = BN * LDS_ROW * 8;
  if constexpr (!std::is_same<Epi, EpiStore>::value) {
    // the tile's GP weights ride along ahead of the first slabs (read in the epilogue)
    double* wsm = Bs + (size_t)STAGES * BN * LDS_ROW;
    if (threadIdx.x < BN / 2) {
      unsigned sw = (unsigned)__cvta_generic_to_shared(wsm + 2 * threadIdx.x);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sw), "l"(epi.w + n0 + 2 * threadIdx.x));
    }
    cp_async_commit();
  }
  // Software pipeline, two levels.  Global -> shared: all STAGES slabs are requested up front; slab kt + STAGES is requested as
  // soon as every warp has taken its last fragments of slab kt.  Shared -> registers: the fragments of k-step kk+1 (of the next
  // slab when kk is the last step) are loaded into the OTHER fragment buffer before the DMMAs of step kk are issued, so a load
  // never waits for the tensor pipe to release the registers it overwrites and the pipe never waits for a load.
#pragma unroll
  for (int s = 0; s < STAGES; s++) {
    if (s < KT) load_stage<BN>(sA0 + s * A_STAGE, sB0 + s * B_STAGE, gA + s * BK, gB + s * BK, lda16, ldb16);
    cp_async_commit();
  }
  cp_async_wait<STAGES - 1>();  // slab 0 (and the weights) have landed
  __syncthreads();
  double af[2][MT], bf[2][NTL];
  auto load_frag = [&](int buf, int stage, int kk) {
    const double* as = As + (size_t)stage * BM * LDS_ROW + (wm * WTM + fr) * LDS_ROW + fk + kk * 4;
    const double* bs = Bs + (size_t)stage * BN * LDS_ROW + (wn * WTN + fr) * LDS_ROW + fk + kk * 4;
#pragma unroll
    for (int i = 0; i < MT; i++) af[buf][i] = as[i * 8 * LDS_ROW];
#pragma unroll
    for (int j = 0; j < NTL; j++) bf[buf][j] = bs[j * 8 * LDS_ROW];
  };
  if (KT > 0) load_frag(0, 0, 0);
  for (int kt = 0; kt < KT; kt++) {
    const int nk = (kt_beg + kt == KT_all - 1) ? nk_last : BK / 4;
    const bool has_next = kt + 1 < KT;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      if (kk >= nk) break;
      if (kk + 1 < nk) {
        load_frag((kk + 1) & 1, kt % STAGES, kk + 1);
      } else if (has_next) {
        cp_async_wait<STAGES - 2>();  // slab kt+1 has landed ...
        __syncthreads();              // ... for everybody, and everybody holds its last fragments of slab kt: its stage is free
        const int kn = kt + STAGES;
        if (kn < KT) {
          const int st = kt % STAGES;
          load_stage<BN>(sA0 + st * A_STAGE, sB0 + st * B_STAGE, gA + (size_t)kn * BK, gB + (size_t)kn * BK, lda16, ldb16);
        }
        cp_async_commit();
        load_frag((kk + 1) & 1, (kt + 1) % STAGES, 0);
      }
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], af[kk & 1][i], bf[kk & 1][j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: thread holds C[row = fr][col = 2*fk, 2*fk+1] of every 8x8 tile ----
  if constexpr (!std::is_same<Epi, EpiStore>::value) {
    const Epi& e = epi;
    double* red = (double*)smem_raw;  // [WARPS_N][BM] (the pipeline stages are drained)
    const double* wsm = Bs + (size_t)STAGES * BN * LDS_ROW;
    const double zeta = e.cp.zeta;
    const bool zeta0 = e.cp.zeta_int == 0;
    double wv[NTL][2];
#pragma unroll
    for (int j = 0; j < NTL; j++) {
      const double2 t = *reinterpret_cast<const double2*>(wsm + wn * WTN + j * 8 + 2 * fk);
      wv[j][0] = t.x;
      wv[j][1] = t.y;
    }
    __syncthreads();  // everyone has its weights: the stage area may now be reused for the row sums
#pragma unroll
    for (int i = 0; i < MT; i++) {
      int row = m0 + wm * WTM + i * 8 + fr;
      double esum = 0.0;
#pragma unroll
      for (int j = 0; j < NTL; j++) {
        double outv[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const double c = acc[i][j][t];
          // c^(zeta-1); a zero weight (padding column, c = 0 exactly) must not meet 0^(negative) in the general-zeta path
          const double pw = (Epi::zi == 0 && wv[j][t] == 0.0) ? 0.0 : e.pw(c);
          esum += wv[j][t] * (zeta0 ? 1.0 : pw * c);     // alpha_s * delta^2 c^zeta cutoff_s   (gp_predict.f95:3766-3768, 3854)
          outv[t] = wv[j][t] * zeta * pw;                // alpha_s * d k_s / d c                (:3771-3772)
        }
        int col = n0 + wn * WTN + j * 8 + 2 * fk;
        *reinterpret_cast<double2*>(e.acoef + (size_t)row * e.lda_out + col) = make_double2(outv[0], outv[1]);
      }
      esum += __shfl_xor_sync(0xffffffffu, esum, 1);
      esum += __shfl_xor_sync(0xffffffffu, esum, 2);
      if (fk == 0) red[wn * BM + wm * WTM + i * 8 + fr] = esum;
    }
    __syncthreads();
    if (threadIdx.x < BM) e.epart[(size_t)(m0 + threadIdx.x) * e.n_tiles_n + blockIdx.x] = red[threadIdx.x] + red[BM + threadIdx.x];
  } else {
    const EpiStore& e = *reinterpret_cast<const EpiStore*>(&epi);
    double* out = e.out + (size_t)blockIdx.z * e.split_stride;
#pragma unroll
    for (int i = 0; i < MT; i++) {
      int row = m0 + wm * WTM + i * 8 + fr;
#pragma unroll
      for (int j = 0; j < NTL; j++) {
        int col = n0 + wn * WTN + j * 8 + 2 * fk;
        *reinterpret_cast<double2*>(out + (size_t)row * e.ldo + col) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

// E_i = sum over column tiles (fixed order) ; local_e(centre) += E_i  (IPModel_GAP.f95:454-459 with cc = 1, |ci| = 1)
__global__ void k_energy_rows(const double* __restrict__ epart, int n_tiles_n, const int* __restrict__ centres, const int* __restrict__ n_centres_dev,
                              int n_centres_ub, double e_scale, double* __restrict__ local_e) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (n_centres_dev ? *n_centres_dev : n_centres_ub)) return;
  double t = 0.0;
  for (int k = 0; k < n_tiles_n; k++) t += epart[(size_t)c * n_tiles_n + k];
  local_e[centres[c]] += e_scale * t;
}

}  // namespace

// Column tile of GEMM-1 for n_rows_pad rows and M sparse points.  All tiles of one launch cost the same, so the launch
// takes ceil(tiles / slots) rounds of (bn + fixed) each: 4,096 centres x 2,000 sparse points are 1,024 tiles of 128
// columns = 3.46 rounds on the 296 CTA slots (4 paid), but 1,152 tiles of 112 columns = 3.89 rounds of 7/8 the length.
int cov_gemm1_bn(int n_rows_pad, int M, int n_sm) {
  const int cand[4] = {128, 112, 96, 80};
  const long slots = 2L * n_sm, row_tiles = n_rows_pad / BM;
  int best = 128;
  double best_cost = 1e300;
  for (int c = 0; c < 4; c++) {
    const int bn = cand[c];
    const long tiles = row_tiles * ((M + bn - 1) / bn), rounds = (tiles + slots - 1) / slots;
    // per-tile cost model: main loop ~ bn columns, + prologue/epilogue and the A-operand share that does not shrink with bn
    const double cost = (double)rounds * (bn + 24.0);
    if (cost < best_cost * 0.98) { best_cost = cost; best = bn; }
  }
  return best;
}

void launch_cov_gemm1(const double* x, int ldx, const double* sp_rows, int lds, int n_rows_pad, int row0, const int* n_rows_dev, int bn, int M,
                      int K, const double* w, CovParams cp, double* acoef, int lda, double* epart, int n_tiles_n, cudaStream_t st,
                      int* launches) {
  auto go = [&](auto bn_tag, auto e) {
    constexpr int BN = decltype(bn_tag)::value;
    using E = decltype(e);
    dim3 grid((M + BN - 1) / BN, n_rows_pad / BM, 1);
    cudaFuncSetAttribute(k_dgemm_nt<BN, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<BN>());
    k_dgemm_nt<BN, E><<<grid, NTHREADS, gemm_smem<BN>(), st>>>(x, ldx, sp_rows, lds, K, row0, n_rows_dev, e);
  };
  auto by_zeta = [&](auto bn_tag) {
    switch (cp.zeta_int) {
      case 1: go(bn_tag, EpiCov<1>{w, cp, acoef, lda, epart, n_tiles_n}); break;
      case 2: go(bn_tag, EpiCov<2>{w, cp, acoef, lda, epart, n_tiles_n}); break;
      case 3: go(bn_tag, EpiCov<3>{w, cp, acoef, lda, epart, n_tiles_n}); break;
      case 4: go(bn_tag, EpiCov<4>{w, cp, acoef, lda, epart, n_tiles_n}); break;
      default: go(bn_tag, EpiCov<0>{w, cp, acoef, lda, epart, n_tiles_n}); break;
    }
  };
  switch (bn) {
    case 112: by_zeta(std::integral_constant<int, 112>{}); break;
    case 96: by_zeta(std::integral_constant<int, 96>{}); break;
    case 80: by_zeta(std::integral_constant<int, 80>{}); break;
    default: by_zeta(std::integral_constant<int, 128>{}); break;
  }
  *launches += 1;
}

int cov_gemm2_bn(int d) {  // column tile of GEMM-2: whichever of 128 / 112 pads the descriptor dimension less
  int p128 = (d + 127) / 128 * 128, p112 = (d + 111) / 112 * 112;
  return p112 < p128 ? 112 : 128;
}

int cov_gemm2_ksplit(int n_rows_pad, int dn_pad, int bn, int n_sm) {  // K splits that best fill 2 CTA slots per SM
  long tiles = (long)(n_rows_pad / BM) * (dn_pad / bn), slots = 2L * n_sm;
  int best = 1;
  double best_eff = 0.0;
  for (int ks = 1; ks <= COV_MAX_KSPLIT; ks++) {
    long units = tiles * ks, waves = (units + slots - 1) / slots;
    double eff = (double)units / (double)(waves * slots);
    if (eff > best_eff + 0.03) { best_eff = eff; best = ks; }
  }
  return best;
}

void launch_cov_gemm2(const double* acoef, int lda, const double* st_rows, int ldst, int n_rows_pad, int row0, const int* n_rows_dev, int dn_pad,
                      int bn, int ksplit, int K, double* gvec, int ldg, size_t split_stride, cudaStream_t st, int* launches) {
  EpiStore e{gvec, ldg, split_stride};
  dim3 grid(dn_pad / bn, n_rows_pad / BM, ksplit);
  if (bn == 112) {
    cudaFuncSetAttribute(k_dgemm_nt<112, EpiStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<112>());
    k_dgemm_nt<112, EpiStore><<<grid, NTHREADS, gemm_smem<112>(), st>>>(acoef, lda, st_rows, ldst, K, row0, n_rows_dev, e);
  } else {
    cudaFuncSetAttribute(k_dgemm_nt<128, EpiStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<128>());
    k_dgemm_nt<128, EpiStore><<<grid, NTHREADS, gemm_smem<128>(), st>>>(acoef, lda, st_rows, ldst, K, row0, n_rows_dev, e);
  }
  *launches += 1;
}

void launch_energy_rows(const double* epart, int n_tiles_n, const int* centres, const int* n_centres_dev, int n_centres_ub, double e_scale,
                        double* local_e, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  k_energy_rows<<<(n_centres_ub + 255) / 256, 256, 0, st>>>(epart, n_tiles_n, centres, n_centres_dev, n_centres_ub, e_scale, local_e);
  *launches += 1;
}

}  // namespace gapb200
