// variance.cu -- predictive variance of the sparse GP (the optional local_gap_variance / gap_variance_gradient outputs of
// IPModel_GAP_Calc, src/Potentials/IPModel_GAP.f95:464-469, 484-487).
//
// Reference arithmetic (src/GAP/gp_predict.f95):
//   gpCoordinates_initialise_variance_estimate (:3970-4085)   k_mm = covariance of the sparse points among themselves,
//       * delta^2 + f0^2, + regularisation^2 on the diagonal, Cholesky-factorised (LA_Matrix_Factorise)
//   gpCoordinates_Predict (:3866-3891)   variance = delta^2 + f0^2 + reg^2 - k . k_mm^-1 k   (Matrix_Solve per descriptor),
//       grad_variance = -2 sparseX (delta^2 zeta c^(zeta-1) cutoff_s * (k_mm^-1 k)_s)          (DOT_PRODUCT)
//                     = -2 grad_k (k_mm^-1 k)                                                  (ARD_SE)
// Here, per SOAP coordinate: C = X S^T on the FP64 tensor cores (the covariance GEMM with a plain store), k element-wise,
// ALL centres solved at once with two triangular solves (cuBLAS dtrsm) against the factor (cuSOLVER dpotrf, once per
// model and regularisation), a warp per centre for the variance and the gradient weights, the covariance GEMM-2 and the
// SOAP adjoint kernel for the pull-back to atoms.  distance_2b (ARD_SE, <= 64 sparse points): a warp per centre with the
// explicit inverse of k_mm in shared memory.  cuBLAS / cuSOLVER are loaded on first use (dlopen), so the energy/force
// path does not depend on them.  This is a diagnostic output, not the hot path.
#include <dlfcn.h>

#include <cublas_v2.h>
#include <cusolverDn.h>

#include <string>

#include "gap_device.cuh"

namespace gapb200 {

namespace {

constexpr double PI_D = 3.14159265358979323846264338327950288;

__device__ __forceinline__ double pow_cov(double c, const CovParams& cp, int dec /* 0: c^zeta, 1: c^(zeta-1) */) {
  if (cp.zeta_int >= 0) {  // fast_pow_1d, gp_predict.f95:3581-3605
    double r = 1.0;
    for (int i = 0; i < cp.zeta_int - dec; i++) r *= c;
    return (cp.zeta_int - dec < 0) ? 0.0 : r;
  }
  return pow(c, cp.zeta - (double)dec);
}

// k_mm(i,j) = delta^2 (s_i . s_j)^zeta + f0^2 (+ reg^2 on the diagonal)   (:4014-4018, 4070-4075); G = S S^T row-major
__global__ void k_kmm_finish(const double* __restrict__ G, int ldg, int M, CovParams cp, double f02, double reg2, double* __restrict__ K) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= M || i >= M) return;
  double v = cp.delta2 * pow_cov(G[(size_t)i * ldg + j], cp, 0) + f02;
  if (i == j) v += reg2;
  K[(size_t)j * M + i] = v;
}

// Q[row][s] = k_s = delta^2 c^zeta cutoff_s   (gp_predict.f95:3766-3768)
__global__ void k_var_prepare(const double* __restrict__ Cm, int ld, int rows, int M, const double* __restrict__ scut, CovParams cp,
                              double* __restrict__ Q) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (s >= M || r >= rows) return;
  Q[(size_t)r * ld + s] = cp.delta2 * pow_cov(Cm[(size_t)r * ld + s], cp, 0) * scut[s];
}

// one warp per centre row: variance = diag - k . (k_mm^-1 k) ; Cm[row][s] <- -2 delta^2 zeta c^(zeta-1) cutoff_s (k_mm^-1 k)_s
__global__ void __launch_bounds__(128) k_var_finish(double* __restrict__ Cm, const double* __restrict__ Q, int ld, int rows, int row0,
                                                    const int* __restrict__ n_rows_dev, int M, int M4, const double* __restrict__ scut,
                                                    CovParams cp, double diag, const int* __restrict__ centres, double* __restrict__ lgv,
                                                    int want_grad, int* __restrict__ neg_flag) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  if (n_rows_dev && row0 + r >= *n_rows_dev) return;
  double acc = 0.0;
  for (int s = lane; s < M4; s += 32) {
    double a2 = 0.0;
    if (s < M) {
      const double c = Cm[(size_t)r * ld + s], q = Q[(size_t)r * ld + s];
      acc += cp.delta2 * pow_cov(c, cp, 0) * scut[s] * q;
      a2 = -2.0 * cp.delta2 * cp.zeta * pow_cov(c, cp, 1) * scut[s] * q;
    }
    if (want_grad) Cm[(size_t)r * ld + s] = a2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const double var = diag - acc;
    if (var < 0.0) atomicOr(neg_flag, 1);  // gp_predict.f95:3877
    lgv[centres[row0 + r]] += var;        // covariance_cutoff = 1, |ci| = 1 (IPModel_GAP.f95:465-468)
  }
}

// distance_2b / ARD_SE variance: one warp per centre, neighbours in turn, lanes over the (<= 64) sparse points
__global__ void __launch_bounds__(128) k_pair2b_var(Pair2bDev p, const double* __restrict__ kinv /* [M][M] */, double diag, int first, int last,
                                                    const int* __restrict__ Zc, const int* __restrict__ nbr_off, const int* __restrict__ nbr_end,
                                                    const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                    const double* __restrict__ pos, const int* __restrict__ Z, Lattice9 lat,
                                                    double* __restrict__ lgv, double* __restrict__ gvg, int* __restrict__ neg_flag) {
  __shared__ double sK[64 * 65];
  const int M = p.M;
  for (int k = threadIdx.x; k < M * M; k += blockDim.x) sK[(k / M) * 65 + (k % M)] = kinv[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, i = first + blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= last || Zc[i] < 0) return;
  const int Zi = Z[i];
  const bool Zi1 = (p.Z1 == 0) || (Zi == p.Z1), Zi2 = (p.Z2 == 0) || (Zi == p.Z2);
  if (!(Zi1 || Zi2)) return;
  const int s0i = lane, s1i = lane + 32;
  const double x0 = s0i < M ? p.sparseX[s0i] : 0.0, c0 = s0i < M ? p.scut[s0i] : 0.0;
  const double x1 = s1i < M ? p.sparseX[s1i] : 0.0, c1 = s1i < M ? p.scut[s1i] : 0.0;
  for (int q = nbr_off[i]; q < nbr_end[i]; q++) {
    int j = nbr_j[q], a0, a1, a2;
    unpack_shift(nbr_s[q], a0, a1, a2);
    double dd[3];
    image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)j, lat.v, a0, a1, a2, dd);
    const double r = norm_nofma(dd);
    if (r >= p.cutoff) continue;  // descriptors.f95:4729
    const int Zj = Z[j];
    const bool Zj1 = (p.Z1 == 0) || (Zj == p.Z1), Zj2 = (p.Z2 == 0) || (Zj == p.Z2);
    if (!((Zi1 && Zj2) || (Zi2 && Zj1))) continue;  // :4733
    // k_s and grad_k_s (gp_predict.f95:3795-3816)
    double t0 = (x0 - r) * p.inv_theta[0], t1 = (x1 - r) * p.inv_theta[0];
    double e0 = p.delta2 * exp(-0.5 * t0 * t0), e1 = p.delta2 * exp(-0.5 * t1 * t1);
    double k0 = s0i < M ? (e0 + p.f02) * c0 : 0.0, k1 = s1i < M ? (e1 + p.f02) * c1 : 0.0;
    double g0 = s0i < M ? e0 * t0 * p.inv_theta[0] * c0 : 0.0, g1 = s1i < M ? e1 * t1 * p.inv_theta[0] * c1 : 0.0;
    double q0 = 0.0, q1 = 0.0;  // (k_mm^-1 k)_s
    for (int t = 0; t < M; t++) {
      const double kt = __shfl_sync(0xffffffffu, t < 32 ? k0 : k1, t & 31);
      if (s0i < M) q0 += sK[s0i * 65 + t] * kt;
      if (s1i < M) q1 += sK[s1i * 65 + t] * kt;
    }
    double kk = k0 * q0 + k1 * q1, gk = g0 * q0 + g1 * q1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kk += __shfl_xor_sync(0xffffffffu, kk, o);
      gk += __shfl_xor_sync(0xffffffffu, gk, o);
    }
    if (lane == 0) {
      const double var = diag - kk, gv1 = -2.0 * gk;
      if (var < 0.0) atomicOr(neg_flag, 1);
      double fc, dfc;  // coordination_function, linearalgebra.f95:7488-7516
      if (r > p.cutoff - p.ctw) {
        double sn, cn;
        sincos(PI_D * (r - p.cutoff + p.ctw) / p.ctw, &sn, &cn);
        fc = 0.5 * (cn + 1.0);
        dfc = -0.5 * PI_D * sn / p.ctw;
      } else { fc = 1.0; dfc = 0.0; }
      const double vc = 0.5 * var * fc * fc;  // IPModel_GAP.f95:465, |ci| = 2
      atomicAdd(&lgv[i], vc);
      atomicAdd(&lgv[j], vc);
      if (gvg) {  // :485-487, rows n = 0 (atom i: grad_data = -u, grad_cc = -fc' u) and n = 1 (atom j: the opposite)
        const double rinv = 1.0 / r;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double u = dd[k] * rinv, tk = gv1 * (-u) * fc * fc + 2.0 * var * fc * (-dfc * u);
          atomicAdd(&gvg[3 * (size_t)i + k], tk);
          atomicAdd(&gvg[3 * (size_t)j + k], -tk);
        }
      }
    }
  }
}

// ---- cuBLAS / cuSOLVER, loaded on first use ----
struct Libs {
  bool tried = false, ok = false;
  std::string err;
  cublasHandle_t hb = nullptr;
  cusolverDnHandle_t hs = nullptr;
  cublasStatus_t (*bCreate)(cublasHandle_t*) = nullptr;
  cublasStatus_t (*bSetStream)(cublasHandle_t, cudaStream_t) = nullptr;
  cublasStatus_t (*bDtrsm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, cublasOperation_t, cublasDiagType_t, int, int, const double*,
                           const double*, int, double*, int) = nullptr;
  cusolverStatus_t (*sCreate)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*sSetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*sPotrfBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*sPotrf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
};
Libs g_libs;

void* open_any(const char* const* names) {
  for (int k = 0; names[k]; k++)
    if (void* h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL)) return h;
  return nullptr;
}
bool load_libs() {
  Libs& L = g_libs;
  if (L.tried) return L.ok;
  L.tried = true;
  const char* bnames[] = {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so", nullptr};
  const char* snames[] = {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", nullptr};
  void* hb = open_any(bnames);
  void* hs = hb ? open_any(snames) : nullptr;
  if (!hb || !hs) {
    L.err = std::string("cannot load ") + (hb ? "cuSOLVER" : "cuBLAS") + " (needed for the GAP variance only): " + (dlerror() ? dlerror() : "");
    return false;
  }
#define SYM(lib, field, name)                                     \
  *(void**)(&L.field) = dlsym(lib, name);                         \
  if (!L.field) { L.err = std::string("missing symbol ") + name; return false; }
  SYM(hb, bCreate, "cublasCreate_v2")
  SYM(hb, bSetStream, "cublasSetStream_v2")
  SYM(hb, bDtrsm, "cublasDtrsm_v2")
  SYM(hs, sCreate, "cusolverDnCreate")
  SYM(hs, sSetStream, "cusolverDnSetStream")
  SYM(hs, sPotrfBuf, "cusolverDnDpotrf_bufferSize")
  SYM(hs, sPotrf, "cusolverDnDpotrf")
#undef SYM
  if (L.bCreate(&L.hb) != CUBLAS_STATUS_SUCCESS) { L.err = "cublasCreate failed"; return false; }
  if (L.sCreate(&L.hs) != CUSOLVER_STATUS_SUCCESS) { L.err = "cusolverDnCreate failed"; return false; }
  L.ok = true;
  return true;
}

}  // namespace

void launch_var_kmm_finish(const double* G, int ldg, int M, CovParams cp, double f02, double reg2, double* K, cudaStream_t st, int* launches) {
  dim3 grid((M + 127) / 128, M);
  k_kmm_finish<<<grid, 128, 0, st>>>(G, ldg, M, cp, f02, reg2, K);
  *launches += 1;
}

// in-place lower Cholesky of the column-major M x M matrix K (synchronises the stream); returns LAPACK info (0 = ok), -1000 = library failure
int var_factorise(double* K, int M, cudaStream_t st, std::string* err) {
  if (!load_libs()) { *err = g_libs.err; return -1000; }
  Libs& L = g_libs;
  L.sSetStream(L.hs, st);
  int lwork = 0;
  if (L.sPotrfBuf(L.hs, CUBLAS_FILL_MODE_LOWER, M, K, M, &lwork) != CUSOLVER_STATUS_SUCCESS) { *err = "cusolverDnDpotrf_bufferSize failed"; return -1000; }
  double* work = nullptr;
  int* info = nullptr;
  if (cudaMalloc(&work, sizeof(double) * (size_t)(lwork > 0 ? lwork : 1)) != cudaSuccess || cudaMalloc(&info, sizeof(int)) != cudaSuccess) {
    *err = "cudaMalloc of the Cholesky workspace failed";
    return -1000;
  }
  cusolverStatus_t s = L.sPotrf(L.hs, CUBLAS_FILL_MODE_LOWER, M, K, M, work, lwork, info);
  int h_info = -1000;
  cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaFree(work);
  cudaFree(info);
  if (s != CUSOLVER_STATUS_SUCCESS) { *err = "cusolverDnDpotrf failed"; return -1000; }
  return h_info;
}

// Q (row-major [rows][ldq], i.e. column-major ldq x rows) <- k_mm^-1 Q with the lower factor Lf (column-major M x M)
int var_solve(const double* Lf, int M, double* Q, int ldq, int rows, cudaStream_t st, std::string* err) {
  if (!load_libs()) { *err = g_libs.err; return 1; }
  Libs& L = g_libs;
  L.bSetStream(L.hb, st);
  const double one = 1.0;
  if (L.bDtrsm(L.hb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, rows, &one, Lf, M, Q, ldq) != CUBLAS_STATUS_SUCCESS ||
      L.bDtrsm(L.hb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, rows, &one, Lf, M, Q, ldq) != CUBLAS_STATUS_SUCCESS) {
    *err = "cublasDtrsm failed";
    return 1;
  }
  return 0;
}

void launch_var_prepare(const double* Cm, int ld, int rows, int M, const double* scut, CovParams cp, double* Q, cudaStream_t st, int* launches) {
  dim3 grid((M + 127) / 128, rows);
  k_var_prepare<<<grid, 128, 0, st>>>(Cm, ld, rows, M, scut, cp, Q);
  *launches += 1;
}

void launch_var_finish(double* Cm, const double* Q, int ld, int rows, int row0, const int* n_rows_dev, int M, const double* scut, CovParams cp,
                       double diag, const int* centres, double* lgv, int want_grad, int* neg_flag, cudaStream_t st, int* launches) {
  k_var_finish<<<(rows + 3) / 4, 128, 0, st>>>(Cm, Q, ld, rows, row0, n_rows_dev, M, (M + 3) & ~3, scut, cp, diag, centres, lgv, want_grad, neg_flag);
  *launches += 1;
}

void launch_pair2b_var(Pair2bDev p, const double* kinv, double diag, int first, int last, const int* Zc, const int* nbr_off, const int* nbr_end,
                       const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, double* lgv, double* gvg, int* neg_flag,
                       cudaStream_t st, int* launches) {
  const int n = last - first;
  if (n <= 0) return;
  k_pair2b_var<<<(n + 3) / 4, 128, 0, st>>>(p, kinv, diag, first, last, Zc, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, lgv, gvg, neg_flag);
  *launches += 1;
}

}  // namespace gapb200
