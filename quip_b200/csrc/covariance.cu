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
// instruction mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- tcgen05 has no FP64 kind).  The operand slabs (16 doubles =
// 128 bytes per row) are brought in by the TMA engine (cp.async.bulk.tensor.2d, SASS UTMALDG) into a 4-stage ring of dense,
// 128B-swizzled stages; one thread arms the stage's mbarrier and issues the two copies, everybody waits on the mbarrier.
// (Per-thread cp.async copies were the one part of the old main loop that cost DMMA issue slots: tools/dmma_loop_probe.cu.)
// Lane (fr, fk) of a DMMA fragment owns k = 4 fk .. 4 fk + 3 of every slab -- a permutation of k common to both operands -- so
// its operands are two LDS.128 per row and slab, conflict-free under the swizzle (16-byte chunk (2 fk + h) ^ fr); the K tail
// (K is a multiple of 4, not of 16; the TMA zero-fills beyond K) runs its last slab in the plain k order with LDS.64.
// CTA tile 64 x BN x 16 (BN = 128, 112, 96 or 80: GEMM-1 picks the one that fills the last round of CTA slots), 4 warps as
// 2(M) x 2(N), warp tile 32 x BN/2 = 4 x (8 .. 5) DMMA tiles, TWO CTAs resident per SM.  GEMM-2 can be split along K
// (= the sparse-point index) into `ksplit` partial outputs when it has too few tiles to fill 2 x 148 CTA slots; the
// consumer (the SOAP adjoint kernel) adds the partials in a fixed order.
#include <cuda.h>

#include <stdexcept>
#include <string>
#include <type_traits>

#include "gap_device.cuh"

namespace gapb200 {

namespace {

constexpr int BM = COV_BM, BK = COV_BK;
constexpr int WARPS_M = 2, WARPS_N = 2, NTHREADS = WARPS_M * WARPS_N * 32;
constexpr int WTM = BM / WARPS_M;   // 32
constexpr int MT = WTM / 8;         // 4 DMMA tiles along M per warp
constexpr int STAGES = 4;
constexpr unsigned ROW_BYTES = BK * sizeof(double);  // 128: one swizzle span
static_assert(ROW_BYTES == 128, "a slab row is one 128-byte swizzle span");
template <int BN>
constexpr size_t gemm_smem() {  // 1 KiB alignment slack + stages + GP weights of the tile + epilogue row sums + mbarriers
  return 1024 + (size_t)STAGES * (BM + BN) * ROW_BYTES + (BN + WARPS_N * BM) * sizeof(double) + STAGES * 8;
}

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
// one box (16 k x box rows) of a row-major FP64 matrix -> a dense, 128B-swizzled stage; completion is signalled on the mbarrier
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int k0, int row0, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst), "l"(map),
               "r"(bar), "r"(k0), "r"(row0)
               : "memory");
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
__global__ void __launch_bounds__(NTHREADS, 2) k_dgemm_nt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                                                          int row0, const int* __restrict__ n_rows_dev, Epi epi) {
  constexpr int WTN = BN / WARPS_N, NTL = WTN / 8;
  constexpr bool COV = !std::is_same<Epi, EpiStore>::value;
  constexpr unsigned A_BYTES = BM * ROW_BYTES, B_BYTES = BN * ROW_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  pdl_launch_dependents();
  unsigned char* const sbase = smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);  // swizzled stages: 1 KiB aligned
  const unsigned base = (unsigned)__cvta_generic_to_shared(sbase);
  double* const wsm = (double*)(sbase + STAGES * STAGE_BYTES);  // [BN] GP weights of this tile's columns
  double* const red = wsm + BN;                                 // [WARPS_N][BM] row sums of the epilogue
  const unsigned bars = base + STAGES * STAGE_BYTES + (BN + WARPS_N * BM) * sizeof(double);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int fr = lane >> 2, fk = lane & 3;  // fragment row (0..7) and k group (0..3)

  // K range of this split (blockIdx.z of gridDim.z), in BK slabs.  K is a multiple of 4 (one DMMA k-step), not of BK.
  const int KT_all = (K + BK - 1) / BK;
  const int nk_last = ((K - 1) % BK) / 4 + 1;
  const int kt_beg = (int)((long long)KT_all * blockIdx.z / gridDim.z), kt_end = (int)((long long)KT_all * (blockIdx.z + 1) / gridDim.z);
  const int KT = kt_end - kt_beg;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  if constexpr (COV) {
    if (threadIdx.x < BN) wsm[threadIdx.x] = epi.w[n0 + threadIdx.x];  // read in the epilogue, many barriers from here (model constants)
  }
  pdl_wait();  // everything above is independent of the preceding kernels; the operands (and the centre count) are not
  if (n_rows_dev && row0 + m0 >= *n_rows_dev) return;  // row tile beyond the (device-side) number of centres
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

  // this lane's fragment rows inside a stage; element (row, k) sits at row * 128 + (((k >> 1) ^ (row & 7)) << 4) + (k & 1) * 8
  const unsigned offA = (wm * WTM + fr) * ROW_BYTES, offB = A_BYTES + (wn * WTN + fr) * ROW_BYTES;
  const unsigned ch0 = ((2 * fk) ^ fr) << 4, ch1 = ((2 * fk + 1) ^ fr) << 4;
  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % STAGES;
    mbar_wait(bars + 8 * s, (kt / STAGES) & 1);
    __syncthreads();  // everybody has finished slab kt-1: its stage may be refilled
    if (threadIdx.x == 0 && kt + STAGES - 1 < KT) issue(kt + STAGES - 1);
    const unsigned st = base + s * STAGE_BYTES;
    if (kt_beg + kt == KT_all - 1 && nk_last < BK / 4) {
      // K tail: plain k order (k = 4 kk + fk), only the k-steps that hold data (two-way bank conflicts, once per tile)
      for (int kk = 0; kk < nk_last; kk++) {
        const int k = 4 * kk + fk;
        const unsigned ch = ((unsigned)((k >> 1) ^ fr) << 4) + (k & 1) * 8;
        double af[MT], bf[NTL];
#pragma unroll
        for (int i = 0; i < MT; i++) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(af[i]) : "r"(st + offA + i * 8 * ROW_BYTES + ch));
#pragma unroll
        for (int j = 0; j < NTL; j++) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(bf[j]) : "r"(st + offB + j * 8 * ROW_BYTES + ch));
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const unsigned ch = h ? ch1 : ch0;
        double a0[MT], a1[MT], b0[NTL], b1[NTL];  // k = 4 fk + 2 h and 4 fk + 2 h + 1
#pragma unroll
        for (int i = 0; i < MT; i++) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a0[i]), "=d"(a1[i]) : "r"(st + offA + i * 8 * ROW_BYTES + ch));
#pragma unroll
        for (int j = 0; j < NTL; j++) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(b0[j]), "=d"(b1[j]) : "r"(st + offB + j * 8 * ROW_BYTES + ch));
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
  }

  // ---- epilogue: thread holds C[row = fr][col = 2*fk, 2*fk+1] of every 8x8 tile ----
  if constexpr (COV) {
    const Epi& e = epi;
    const double zeta = e.cp.zeta;
    const bool zeta0 = e.cp.zeta_int == 0;
    double wv[NTL][2];
#pragma unroll
    for (int j = 0; j < NTL; j++) {
      const double2 t = *reinterpret_cast<const double2*>(wsm + wn * WTN + j * 8 + 2 * fk);
      wv[j][0] = t.x;
      wv[j][1] = t.y;
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



// ---- TMA descriptors (host) ----
// cuTensorMapEncodeTiled comes from the driver; it is looked up once through the runtime, so the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled is not available from this CUDA driver (the covariance GEMM needs TMA)");
  return fn;
}
// row-major FP64 matrix [rows][ld] of which the first K columns are valid (the TMA zero-fills beyond K and beyond rows);
// box = one slab (16 k) x box_rows rows, written to shared memory with the 128-byte swizzle
CUtensorMap operand_map(const double* ptr, long rows, int K, int ld, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return m;
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
    const CUtensorMap tmA = operand_map(x, n_rows_pad, K, ldx, BM), tmB = operand_map(sp_rows, (long)grid.x * BN, K, lds, BN);
    cudaFuncSetAttribute(k_dgemm_nt<BN, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<BN>());
    launch_pdl(k_dgemm_nt<BN, E>, grid, dim3(NTHREADS), gemm_smem<BN>(), st, tmA, tmB, K, row0, n_rows_dev, e);
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
  const CUtensorMap tmA = operand_map(acoef, n_rows_pad, K, lda, BM), tmB = operand_map(st_rows, dn_pad, K, ldst, bn == 112 ? 112 : 128);
  if (bn == 112) {
    cudaFuncSetAttribute(k_dgemm_nt<112, EpiStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<112>());
    launch_pdl(k_dgemm_nt<112, EpiStore>, grid, dim3(NTHREADS), gemm_smem<112>(), st, tmA, tmB, K, row0, n_rows_dev, e);
  } else {
    cudaFuncSetAttribute(k_dgemm_nt<128, EpiStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<128>());
    launch_pdl(k_dgemm_nt<128, EpiStore>, grid, dim3(NTHREADS), gemm_smem<128>(), st, tmA, tmB, K, row0, n_rows_dev, e);
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
