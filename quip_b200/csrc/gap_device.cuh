// gap_device.cuh -- device-side data layout and kernel launchers of the B200 GAP path.
//
// HBM layout (all FP64 / int32):
//   pos[N][3], Z[N]                      inputs (xyz per atom = Fortran pos(3,N))
//   nbr_off[N+1], nbr_j[nnz], nbr_s[nnz] full neighbour list in CSR; shift packed 3 x int8
//   x[Nc_pad][d_pad]                     normalised SOAP vectors, one row per centre
//   xlm[Nc][nlm*K1]                      real-harmonic density expansion kept for the adjoint
//   acoef[Nc_pad][M_pad]                 alpha_s * dk_s/dc  (GEMM-1 epilogue output)
//   gvec[Nc_pad][dn_pad]                 dE_i/dx (GEMM-2 output = the reference's gradPredict)
//   out_force[N][3], out_le[N], vir_part[slots][9]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

namespace gapb200 {

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------------
// The kernels of one evaluation form a chain on one stream.  Launched with the programmatic-stream-serialization attribute, a
// kernel's CTAs may become resident while its predecessor drains (its last round of CTAs), so the launch latency and the ramp-up
// disappear from the critical path.  Contract: such a kernel calls pdl_wait() before it touches anything a predecessor wrote
// (griddepcontrol.wait returns when the prerequisite grids have completed and their writes are visible), and pdl_launch_dependents()
// as early as possible.  Both are no-ops in a kernel launched the ordinary way.  GAP_B200_PDL=0 turns the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("GAP_B200_PDL"); return !(e && *e == '0'); }();
  return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

constexpr int SOAP_NMAX_CAP = 16;   // n_max
constexpr int SOAP_LMAX_CAP = 12;   // l_max
constexpr int SOAP_SPECIES_CAP = 8; // n_species, n_Z

struct SoapDev {
  double cutoff, ctw, alpha, central_weight, sigma0, chol00;
  double cutoff_scale, cutoff_rate, norm_radial_decay;
  int cutoff_dexp;
  int l_max, n_max, n_species, n_Z, K1, d, d_pad, nlm;
  int normalise, cras, two_lp1;
  int species_Z[SOAP_SPECIES_CAP];
  int centre_Z[SOAP_SPECIES_CAP];
  double r_basis[SOAP_NMAX_CAP];
  double T[SOAP_NMAX_CAP * SOAP_NMAX_CAP];                            // T[a + n_max*a'] (column-major)
  double ynorm[(SOAP_LMAX_CAP + 1) * (SOAP_LMAX_CAP + 2) / 2];        // N_lm * (m>0 ? sqrt2 : 1), index l(l+1)/2+m
  double tlpo[SOAP_LMAX_CAP + 1];                                     // 1/sqrt(2l+1) or 1
};

struct CellGrid {
  double lat[9];   // column-major lattice(k,m) = lat[k+3m]
  double g[9];     // inverse: frac_r = sum_c g[r+3c] * pos_c
  double toff[3];  // non-periodic directions: minimum fractional coordinate
  double tscale[3];// non-periodic directions: 1/extent
  int n[3], R[3], pbc[3];
  double cutoff;
};

struct Lattice9 {
  double v[9];
};

__host__ __device__ inline int pack_shift(int a, int b, int c) { return (a & 0xff) | ((b & 0xff) << 8) | ((c & 0xff) << 16); }
__host__ __device__ inline void unpack_shift(int p, int& a, int& b, int& c) {
  a = (int)(int8_t)(p & 0xff);
  b = (int)(int8_t)((p >> 8) & 0xff);
  c = (int)(int8_t)((p >> 16) & 0xff);
}

// Displacement pos_j - pos_i + lattice*shift with the reference's operation order and NO fused
// multiply-add (Connection.f95:486-489, 2120-2125), so distances are bit-identical to the oracle.
__device__ __forceinline__ void image_diff(const double* __restrict__ pi, const double* __restrict__ pj, const double* lat, int s0,
                                           int s1, int s2, double* dd) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double t = __dsub_rn(pj[k], pi[k]);
    t = __dadd_rn(t, __dmul_rn(lat[k + 0], (double)s0));
    t = __dadd_rn(t, __dmul_rn(lat[k + 3], (double)s1));
    t = __dadd_rn(t, __dmul_rn(lat[k + 6], (double)s2));
    dd[k] = t;
  }
}
__device__ __forceinline__ double norm_nofma(const double* dd) {
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dd[0], dd[0]), __dmul_rn(dd[1], dd[1])), __dmul_rn(dd[2], dd[2])));
}

// ---- neighbour.cu ------------------------------------------------------------------------
struct NeighbourWork {  // device scratch owned by the potential handle; sized for N atoms / ncell cells
  int* cell_of = nullptr;      // [N]
  int* mshift = nullptr;       // [N] packed map_shift
  int* sort_keys = nullptr;    // [N]
  int* sort_idx = nullptr;     // [N] atom ids sorted by cell (stable)
  int* slot_of = nullptr;      // [N] inverse: sorted slot of atom i
  int* iota = nullptr;         // [N]
  int* keys_tmp = nullptr;     // [N]
  int* cell_count = nullptr;   // [ncell+1] counting-sort path only
  int* cell_start = nullptr;   // [ncell+1]
  double* spos = nullptr;      // [N][3] positions in sorted order
  int* smshift = nullptr;      // [N]
  int* nn = nullptr;           // [N+1] neighbour counts (original atom order)
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  double* minmax = nullptr;    // [6]
  int* err_flag = nullptr;     // [1] set by the kernels on unrepresentable image shifts
};
size_t neighbour_cub_bytes(int N, int ncell);
void launch_frac_minmax(const double* pos, int N, const double* g9_dev_unused, const CellGrid& grid, double* minmax6, cudaStream_t st, int* launches);
void launch_bin_atoms(const double* pos, int N, const CellGrid& grid, int ncell, NeighbourWork& w, cudaStream_t st, int* launches);
// rows are built only for the centres [first, last) (the partition of this handle); other rows are empty
// exact layout: count (+ the largest row -> *max_row), exclusive scan into nbr_off[N+1], fill packed CSR rows (nbr_end = nbr_off + 1)
void launch_neigh_count(const double* pos, int N, int first, int last, const CellGrid& grid, NeighbourWork& w, int* nbr_off, int* max_row,
                        cudaStream_t st, int* launches);
void launch_neigh_fill(const double* pos, int N, int first, int last, const CellGrid& grid, NeighbourWork& w, int* nbr_off, int* nbr_j,
                       int* nbr_s, double* nbr_d, int cap, cudaStream_t st, int* launches);
// speculative layout: ONE pass into rows of row_cap entries at (i - first) * row_cap; entries beyond the capacity are dropped
// and *max_row (> row_cap) tells the host to repeat the call with the exact layout
void launch_neigh_onepass(const double* pos, int N, int first, int last, const CellGrid& grid, NeighbourWork& w, int* nbr_off, int* nbr_end,
                          int* nbr_j, int* nbr_s, int row_cap, int* max_row, cudaStream_t st, int* launches);

// ---- soap.cu -------------------------------------------------------------------------------
void launch_select_centres(const int* Z, int first, int last, const SoapDev* sp, int* flags, cudaStream_t st, int* launches);
void launch_compact(const int* flags_scan, const int* flags, int first, int n, int* centres, cudaStream_t st, int* launches);
size_t soap_forward_smem(const SoapDev& h);
size_t soap_adjoint_smem(const SoapDev& h);
// n_centres_dev: device int holding the number of centres (the grid is sized by the host-side upper bound n_centres_ub)
void launch_soap_forward(const SoapDev* sp, const SoapDev& h, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off, const int* nbr_end,
                         const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, double* x, double* xlm, double* pnorm,
                         cudaStream_t st, int* launches, int skip_power = 0);
// epart/n_tiles_n/local_e: if epart != NULL the kernel also folds the GEMM-1 row sums into local_e (E_i)
void launch_soap_adjoint(const SoapDev* sp, const SoapDev& h, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off, const int* nbr_end,
                         const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, const double* x, const double* xlm,
                         const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride, const double* epart, int n_tiles_n,
                         double* local_e, double e_scale, double* force, double* vir_part, double* local_virial, double* fpair, cudaStream_t st,
                         int* launches, const double* lambda_in = nullptr);
// fpair != NULL selects the DETERMINISTIC scatter: the force on neighbour j of a pair goes to fpair[3 * slot] (slot = position of the
// entry in the neighbour list), the centre's own sum to force[3 * i] by a plain store (force = a zeroed per-atom scratch buffer); the
// caller then adds both per atom in a fixed order (potential.cu k_det_gather).  fpair == NULL: FP64 atomics on force[].

// ---- soap_general.cu: compression modes and GTO / POLY radial bases (tables built by gap_model.cpp soap_general_setup) ----
struct SoapGenDev {
  int n_grid, Ka, Kb, n_pairs;
  int global_mode;         // average=T: one descriptor per configuration (sum of the centres' density expansions)
  double* Xg;              // global_mode: [nlm][K1] summed density expansion
  double* Lt;              // global_mode: [nlm][n_species * n_grid] dE/dX on the radial grid, shared by all centres
  const double* r_grid;    // [n_grid]
  const double* P;         // [l_max+1][n_grid][n_max]
  const double* c0;        // [n_max]
  const double* W1;        // [K1][Ka]
  const double* W2;        // [K1][Kb]
  const int* pair_ia;      // [n_pairs]
  const int* pair_jb;
  const double* pair_fac;
  // the elements grouped by their first / second channel (for dE/dY1, dE/dY2): offsets [Ka + 1] / [Kb + 1] into element lists [n_pairs]
  const int *by_ia_off, *by_ia, *by_jb_off, *by_jb;
};
size_t soap_general_smem(const SoapDev& h, const SoapGenDev& g);
void launch_soap_forward_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                                 const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                                 double* x, double* xlm, double* pnorm, cudaStream_t st, int* launches);
void launch_soap_adjoint_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                                 const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                                 const double* x, const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride,
                                 const double* epart, int n_tiles_n, double* local_e, double e_scale, double* force, double* vir_part,
                                 double* local_virial, double* fpair, cudaStream_t st, int* launches);

// ---- covariance.cu -------------------------------------------------------------------------
struct CovParams {
  double delta2;   // delta^2
  double zeta;
  int zeta_int;    // >=0: integer fast path (fast_pow_1d), -1: general pow
};
constexpr int COV_BM = 64, COV_BN1_MAX = 128, COV_BK = 16, COV_MAX_KSPLIT = 4;
// GEMM-1 + epilogue: c = X S^T ; k = delta^2 c^zeta cutoff_s ; acoef = alpha_s delta^2 zeta c^(zeta-1) cutoff_s ;
// epart[row][n_tile] = sum over the tile's columns of alpha_s k
// n_rows_dev (may be NULL): device int, tiles whose first row is >= *n_rows_dev are skipped
// w[M_pad] = alpha_s * sparseCutoff_s * delta^2 (0 in the padding)
// bn = column tile (cov_gemm1_bn: 128, 112, 96 or 80, whichever wastes the fewest CTA slots of the last round);
// n_tiles_n = ceil(M / bn); sp_rows, w and acoef are allocated M_pad >= M + 127 wide so that any tiling stays in bounds.
// K = descriptor dimension rounded up to 4 (operand rows are zero-padded to a multiple of COV_BK).
int cov_gemm1_bn(int n_rows_pad, int M, int n_sm);
void launch_cov_gemm1(const double* x, int ldx, const double* sp_rows, int lds, int n_rows_pad, int row0, const int* n_rows_dev, int bn, int M,
                      int K, const double* w, CovParams cp, double* acoef, int lda, double* epart, int n_tiles_n, cudaStream_t st,
                      int* launches);
// GEMM-2: gvec[split] = acoef[:, K range of split] S   (S given transposed: st_rows[q][s]); column tile bn (112 or 128),
// ksplit partial outputs split_stride doubles apart
int cov_gemm2_bn(int d);
int cov_gemm2_ksplit(int n_rows_pad, int dn_pad, int bn, int n_sm);
void launch_cov_gemm2(const double* acoef, int lda, const double* st_rows, int ldst, int n_rows_pad, int row0, const int* n_rows_dev, int dn_pad,
                      int bn, int ksplit, int K /* sparse points rounded up to 4 */, double* gvec, int ldg, size_t split_stride, cudaStream_t st,
                      int* launches);
void launch_energy_rows(const double* epart, int n_tiles_n, const int* centres, const int* n_centres_dev, int n_centres_ub, double e_scale,
                        double* local_e, cudaStream_t st, int* launches);

// compression modes on the EQUISPACED_GAUSS basis ("hybrid"): soap.cu's kernels for the density expansion (skip_power) and the neighbour phase
// (lambda_in), these two for the variant-specific middle: X_lm -> descriptor, dE/dx -> Lambda = dE/dX_lm [centre][nlm][K1]
void launch_soap_power_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* n_centres_dev, int n_centres_ub, const double* xlm,
                               double* x, double* pnorm, cudaStream_t st, int* launches);
void launch_soap_lambda_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* n_centres_dev, int n_centres_ub, const double* x,
                                const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride,
                                double* lambda_out, cudaStream_t st, int* launches);

// GTO / POLY in grid passes of gp <= 16 points through soap.cu's kernels (see soap_general.cu): pass buffers [pass][centre][nlm][n_species * gp]
void launch_soap_power_grid(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                            const int* Z, const double* xt_pass, size_t pass_stride, int gp, double* x, double* xlm, double* pnorm, cudaStream_t st,
                            int* launches);
void launch_soap_lambda_grid(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* n_centres_dev, int n_centres_ub, const double* x,
                             const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride, double* lam_pass,
                             size_t pass_stride, int gp, cudaStream_t st, int* launches);

// ---- pair2b.cu -----------------------------------------------------------------------------
constexpr int PAIR2B_MAX_EXP = 4;
struct Pair2bDev {
  double cutoff, ctw, delta2, f02;
  double inv_theta[PAIR2B_MAX_EXP];   // 1 / theta_k
  double exponents[PAIR2B_MAX_EXP];   // data_k = r^exponents_k (descriptors.f95:4750)
  int n_exp;
  int tail_exponent;                  // covariance_cutoff *= (erf(tail_range r) / r)^tail_exponent (:4743-4747)
  double tail_range;
  int intra_mode;                     // 0: all pairs, 1: only_intra, 2: only_inter (residue ids, :4735-4738)
  const int* resid;                   // [N] or NULL
  int Z1, Z2, M;
  const double* sparseX;  // [M][n_exp]
  const double* alpha;    // [M]
  const double* scut;     // [M]
};
void launch_pair2b(Pair2bDev p, int first, int last, const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z,
                   const int* Zc /* < 0: not a centre */, int scatter /* 1: atom mask active, scatter to the pair partner */, Lattice9 lat, double e_scale, int do_grad, double* local_e, double* force, double* vir_part, double* local_virial,
                   cudaStream_t st, int* launches, int* n_blocks_out);

// ---- angle3b.cu ----------------------------------------------------------------------------
struct Angle3bDev {
  double cutoff, ctw;
  double inv_theta[3];
  double e_f0;            // f0^2 sum_s alpha_s sparseCutoff_s: the constant part of every instance's energy (gp_predict.f95:3812)
  int Zc, Z1, Z2, M;
  const double* table;    // [M][4]: sparseX_s / theta (3), alpha_s sparseCutoff_s delta^2
};
// cidx: int scratch parallel to the list slots (compacted in-cutoff entries of each row).  fpair != NULL: deterministic mode (pair forces per
// list slot, `force` = the per-atom buffer of the centres' own sums)
void launch_angle3b(Angle3bDev p, int first, int last, const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos,
                    const int* Z, const int* Zc /* < 0: not a centre */, Lattice9 lat, double e_scale, int do_grad, double* local_e, double* force,
                    double* fpair, double* vir_part, double* local_virial, int* cidx, cudaStream_t st, int* launches, int* n_blocks_out);

// ---- variance.cu (optional local_gap_variance output; cuBLAS / cuSOLVER loaded on first use) -----------------
}  // namespace gapb200
#include <string>
namespace gapb200 {
void launch_var_kmm_finish(const double* G, int ldg, int M, CovParams cp, double f02, double reg2, double* K, cudaStream_t st, int* launches);
int var_factorise(double* K, int M, cudaStream_t st, std::string* err);
int var_solve(const double* Lf, int M, double* Q, int ldq, int rows, cudaStream_t st, std::string* err);
void launch_var_prepare(const double* Cm, int ld, int rows, int M, const double* scut, CovParams cp, double* Q, cudaStream_t st, int* launches);
void launch_var_finish(double* Cm, const double* Q, int ld, int rows, int row0, const int* n_rows_dev, int M, const double* scut, CovParams cp,
                       double diag, const int* centres, double* lgv, int want_grad, int* neg_flag, cudaStream_t st, int* launches);
void launch_pair2b_var(Pair2bDev p, const double* kinv, double diag, int first, int last, const int* Zc, const int* nbr_off, const int* nbr_end,
                       const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, double* lgv, double* gvg, int* neg_flag,
                       cudaStream_t st, int* launches);

// ---- finalize (potential.cu) ----------------------------------------------------------------
void launch_finalize(const int* Z, int N, int first, int last, const double* e0_dev, double e_scale, double* local_e, const double* vir_part,
                     int n_vir_slots, double* packed_out /* [E, virial(9)] */, cudaStream_t st, int* launches);

}  // namespace gapb200
