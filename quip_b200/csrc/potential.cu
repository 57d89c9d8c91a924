// potential.cu -- the gap_potential handle: device-resident model, workspaces, the calc pipeline and the C ABI
// declared in include/gap_b200.h.
//
// Orchestration replaced: potential_calc's automatic calc_connect (src/Potentials/Potential.f95:844-859) and
// IPModel_GAP_Calc (src/Potentials/IPModel_GAP.f95:233-605): loop over coordinates -> descriptor -> gp_predict ->
// scatter -> totals -> e0 -> E_scale.  Everything between the H2D copy of (pos, Z) and the D2H copy of the results
// runs on one CUDA stream; the host only reads back two integers per call (neighbour-entry count, centre count).
#include <cub/cub.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/gap_b200.h"
#include "gap_comm.h"
#include "gap_device.cuh"
#include "gap_model.h"

using namespace gapb200;

namespace {

thread_local std::string g_last_error;

#define CUDA_OK(expr)                                                                                               \
  do {                                                                                                              \
    cudaError_t _e = (expr);                                                                                        \
    if (_e != cudaSuccess)                                                                                          \
      throw GapError(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T>
  T* as() const { return (T*)p; }
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cap = 0;
      throw GapError(std::string("cudaMalloc of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    }
    cap = want;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct CoordDev {
  int kind = 0;
  // soap
  SoapDev h;
  SoapDev* d_sp = nullptr;
  bool general = false;        // compression modes / GTO / POLY: soap_general.cu kernels
  bool hybrid = false;         // general on the EQUISPACED_GAUSS basis: soap.cu's kernels for the density expansion and the neighbour phase
  // GTO / POLY through soap.cu's kernels in passes over the radial grid (grid_passes slices of grid_gp <= 16 points; 0 = not used):
  // h_grid / d_sp_grid = clones of the descriptor whose basis points are the slice, identity transform, no central term
  int grid_passes = 0, grid_gp = 0;
  SoapDev h_grid;
  SoapDev* d_sp_grid[3] = {nullptr, nullptr, nullptr};
  SoapGenDev gen;
  void* gen_blob = nullptr;    // one device allocation behind the pointers of gen
  double* gen_global = nullptr; // average=T: [Xg | Lt] (see SoapGenDev)
  double *sp_rows = nullptr, *st_rows = nullptr, *alpha = nullptr, *scut = nullptr;
  int M = 0, M_pad = 0, d_pad = 0, dn_pad = 0, bn2 = 128;
  CovParams cp;
  // distance_2b
  Pair2bDev p2;
  double *x2 = nullptr, *a2 = nullptr, *c2 = nullptr;
  // angle_3b
  Angle3bDev p3;
  double* t3 = nullptr;
  // variance estimate (built on first use for a given regularisation): soap = lower Cholesky factor of k_mm (M x M,
  // column-major), distance_2b = explicit inverse of k_mm (M x M)
  double* var_mat = nullptr;
  double var_reg = -1.0, f0 = 0.0, delta = 0.0;
};

}  // namespace

struct gap_potential {
  GapModel model;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<CoordDev> cd;
  double* d_e0 = nullptr;
  int rank = 0, n_ranks = 1;
  GapComm* comm = nullptr;     // set by gap_potential_set_comm: the reduction over ranks happens inside the library
  int n_sm = 148;
  int g_splits = 1;            // K splits of the last GEMM-2 (partial gvec buffers)
  int g_tiles_n = 1;           // column tiles of the last GEMM-1 (partial energies per row in epart)
  size_t g_split_stride = 0;
  // speculative neighbour-list sizing: the entry count of the previous call sizes the buffers of the next one, the
  // real count comes back asynchronously (pinned) and is verified after the final synchronisation of the call
  unsigned char* h_stage = nullptr;  // pinned staging buffer of the host-pointer entry point (inputs in, results out)
  size_t h_stage_cap = 0;
  int* h_pin = nullptr;        // pinned + mapped [3]: entry count (exact layout only), error flag, largest row
  int* d_hpin = nullptr;       // device address of h_pin (k_finalize writes the status there: no copy node in the stream)
  bool stat_clean = false;     // the device status words were reset by the last k_finalize (speculative build needs no memset)
  int row_hint = -1;           // largest neighbour row of the previous call with the same (N, first, last)
  int hint_N = -1, hint_first = -1, hint_last = -1;
  bool pending_check = false;
  long pending_cap = 0;        // row capacity of the speculative layout in flight
  unsigned int* d_fin_counter = nullptr;
  // neighbour list the descriptor kernels read: the handle's own (build_connect) or one supplied by the caller (LAMMPS entry)
  // (row i = entries [cv_off[i], cv_end[i]); cv_end = cv_off + 1 for packed CSR rows)
  const int *cv_off = nullptr, *cv_end = nullptr, *cv_j = nullptr, *cv_s = nullptr;
  DevBuf b_xoff, b_xj, b_xs, b_zc;  // device copies of an external list and of the centre mask
  long launches = 0;
  double last_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // optional inputs / outputs of the last calc (IPModel_GAP.f95:324-337)
  DevBuf b_mask;               // atom mask (ints) set with gap_potential_set_atom_mask
  int mask_N = -1;
  DevBuf b_resid;              // residue ids (ints) set with gap_potential_set_resid (distance_2b only_intra / only_inter)
  int resid_N = -1;
  DevBuf b_epc;                // running sums of local_e after each coordinate -> energy_per_coordinate
  bool epc_valid = false;
  DevBuf b_lgv, b_gvg, b_varflag, b_vc, b_vq, b_vk;  // local_gap_variance[N], gap_variance_gradient[3N], negative-variance flag, work
  int var_N = -1;
  bool var_grad = false;
  cudaStream_t last_stream = nullptr;

  // deterministic force scatter (gap_potential_set_deterministic): pair forces stored per list slot, summed per atom in a fixed order
  bool deterministic = false;
  bool list_reused = false;    // the last build_connect kept the previous list (the reverse index of the slots is still valid)
  long ext_nnz = 0;            // entries of an external (LAMMPS) list
  long det_slots = 0;          // slots covered by the reverse index
  int list_row_cap = 0;        // layout of the handle's own list: fixed-capacity rows of this many slots, or 0 = packed CSR
  long list_slots = 0;         // slots of the handle's own list (row_cap * centres, or the entry count)
  DevBuf b_fpair, b_fself, b_dkeys, b_dkeys2, b_dvals, b_dvals2, b_joff, b_dcub;
  DevBuf b_a3idx;              // angle_3b: compacted in-cutoff entries of each list row (int per slot)
  DevBuf b_lambda;             // hybrid SOAP coordinates: Lambda = dE/dX_lm [centre][nlm][K1]
  DevBuf b_xt_pass, b_lam_pass; // GTO / POLY grid passes: [pass][centre][nlm][n_species * gp]
  // skin-based reuse of the neighbour list (calc_connect with cutoff_skin, Connection.f95:1085-1128)
  double cutoff_skin = 0.0;
  bool list_valid = false;     // cv_* describe a list built with last_cut for the geometry remembered below
  DevBuf b_lastpos, b_disp;    // positions at the last build; per-block maxima of the squared displacement
  double last_lat[9] = {0}, last_cut = 0.0;
  int last_pbc[3] = {0, 0, 0}, last_N = -1, last_first = -1, last_last = -1;
  long n_rebuilds = 0, n_reuses = 0;
  double* h_disp = nullptr;    // pinned [DISP_BLOCKS]
  // neighbour list state
  NeighbourWork nw;
  DevBuf b_cell_of, b_mshift, b_keys, b_idx, b_slot, b_iota, b_cstart, b_ccount, b_spos, b_smshift, b_nn, b_cub, b_minmax;
  DevBuf b_off, b_end, b_j, b_s, b_d;
  int conn_N = 0, conn_nnz = 0;
  // inputs / outputs owned for the host-pointer API
  DevBuf b_pos, b_Z, b_packed, b_le, b_lv;
  // per-coordinate workspaces
  DevBuf b_velo, b_velo2, b_acc, b_mass, b_ke;  // MD driver state
  DevBuf b_flags, b_scan, b_centres, b_x, b_xlm, b_pnorm, b_acoef, b_gvec, b_epart, b_vir, b_fin;
  int timing = 0;              // 0: no events; 1: per-stage CUDA events; 2: only around the covariance GEMMs (gap_potential_set_timing)
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_stage;
  size_t ev_used = 0;
};

namespace {

// ---------------------------------------------------------------------------------------------------
// finalize kernels: local_e += e0 ; E = sum local_e ; virial = sum vir_part  (IPModel_GAP.f95:523-532, 576-596)
// ---------------------------------------------------------------------------------------------------
constexpr int FIN_BLOCKS = 128, FIN_THREADS = 256;

// One kernel: per-block partial sums (warp shuffles, then one pass over the 8 warp rows), and the LAST block to finish
// (atomic ticket) adds the partials in block order, so the totals are deterministic.
__global__ void __launch_bounds__(FIN_THREADS) k_finalize(const int* __restrict__ Z, int N, int first, int last, const double* __restrict__ e0,
                                                          double e_scale, double* __restrict__ local_e, const double* __restrict__ vir_part,
                                                          int n_slots, double* __restrict__ part /* [gridDim][10] */,
                                                          unsigned int* __restrict__ counter, double* __restrict__ packed,
                                                          int* __restrict__ stat_dev /* [entry count, error flag, largest row] or NULL */, long cap,
                                                          int* __restrict__ stat_host /* the same three words in mapped host memory */) {
  __shared__ double wsum[FIN_THREADS / 32][10];
  __shared__ bool is_last;
  pdl_launch_dependents();
  pdl_wait();
  const int stride = gridDim.x * FIN_THREADS;
  double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * FIN_THREADS + threadIdx.x; i < N; i += stride) {
    double le = local_e[i];
    if (i >= first && i < last) {
      int z = Z[i];
      le += e_scale * ((z >= 0 && z < 128) ? e0[z] : 0.0);
      local_e[i] = le;
    }
    v[0] += le;
  }
  for (int s = blockIdx.x * FIN_THREADS + threadIdx.x; s < n_slots; s += stride)
#pragma unroll
    for (int k = 0; k < 9; k++) v[1 + k] += vir_part[9 * (size_t)s + k];
#pragma unroll
  for (int k = 0; k < 10; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < FIN_THREADS / 32; w++) t += wsum[w][threadIdx.x];
    part[10 * blockIdx.x + threadIdx.x] = t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    // the partials of all blocks are fetched in parallel (a chain of dependent L2 loads costs ~0.3 us each), then added in block order
    __shared__ double sp[FIN_BLOCKS * 10];
    for (int k = threadIdx.x; k < 10 * (int)gridDim.x; k += FIN_THREADS) sp[k] = __ldcg(&part[k]);
    __syncthreads();
    if (threadIdx.x < 10) {
      double t = 0.0;
      for (int b = 0; b < (int)gridDim.x; b++) t += sp[10 * b + threadIdx.x];
      // a speculatively sized neighbour list that overflowed poisons the energy: after the all-reduce EVERY rank sees the NaN
      // and repeats the evaluation, with no extra collective
      if (threadIdx.x == 0 && stat_dev && (long)stat_dev[2] > cap) t = __longlong_as_double(0x7ff8000000000000LL);
      packed[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) {
      *counter = 0u;
      if (stat_dev) {  // the host reads the list's status from mapped memory after its synchronisation (no copy node in the stream);
                       // the device words are left clean for the next speculative build
        stat_host[1] = stat_dev[1];  // (visible to the host once the kernel has completed: no fence needed)
        stat_host[2] = stat_dev[2];
        stat_dev[0] = stat_dev[1] = stat_dev[2] = 0;
      }
    }
  }
}

// fixed-order sum of local_e (one block): the running total after each GP coordinate gives energy_per_coordinate
__global__ void __launch_bounds__(1024) k_sum_range(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double ws[32];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) t += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    t = ws[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *out = t;
  }
}
// atomic numbers as the centre selection sees them: -1 for atoms outside the mask (atom_mask_name, IPModel_GAP.f95:344-346)
__global__ void k_apply_mask(const int* __restrict__ Z, const int* __restrict__ mask, int N, int* __restrict__ Zc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) Zc[i] = mask[i] ? Z[i] : -1;
}

// zero the outputs of a calc in one launch: packed [E | virial | F], local_e, local_virial (may be NULL)
__global__ void k_zero_outputs(double* __restrict__ a, size_t na, double* __restrict__ b, size_t nb, double* __restrict__ c, size_t nc) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t t = t0; t < na; t += stride) a[t] = 0.0;
  for (size_t t = t0; t < nb; t += stride) b[t] = 0.0;
  for (size_t t = t0; t < nc; t += stride) c[t] = 0.0;
}

// Centre selection of a SOAP coordinate for small partitions, ONE block: flag (descriptors.f95:7962), block scan and ordered
// compaction, 4 atoms per thread and 4,096 per round; the count goes where the multi-kernel path leaves it (scan[n]).
// Blocks 1.. of the grid (if any) zero the outputs of the calc instead (one launch for both jobs at the start of a step).
constexpr int SEL_THREADS = 1024, SEL_ITEMS = 4, SEL_MAX_N = 16384, SEL_ZERO_BLOCKS = 8;
struct CentreZ {  // Z(:) of a soap descriptor (descriptors.f95:7962), by value: no dependent loads at the head of the step
  int n_Z;
  int Z[SOAP_SPECIES_CAP];
};
__global__ void __launch_bounds__(SEL_THREADS) k_select_compact_block(const int* __restrict__ Z, int first, int last, CentreZ cz,
                                                                       int* __restrict__ centres, int* __restrict__ count_out,
                                                                       double* __restrict__ za, size_t nza, double* __restrict__ zb, size_t nzb,
                                                                       double* __restrict__ zc, size_t nzc) {
  typedef cub::BlockScan<int, SEL_THREADS> BS;
  __shared__ typename BS::TempStorage tmp;
  pdl_launch_dependents();
  pdl_wait();
  if (blockIdx.x > 0) {
    const size_t stride = (size_t)(gridDim.x - 1) * SEL_THREADS, t0 = (size_t)(blockIdx.x - 1) * SEL_THREADS + threadIdx.x;
    for (size_t t = t0; t < nza; t += stride) za[t] = 0.0;
    for (size_t t = t0; t < nzb; t += stride) zb[t] = 0.0;
    for (size_t t = t0; t < nzc; t += stride) zc[t] = 0.0;
    return;
  }
  const int n = last - first, nZ = cz.n_Z;
  int base = 0;
  for (int t0 = 0; t0 < n; t0 += SEL_THREADS * SEL_ITEMS) {
    int f[SEL_ITEMS], pos[SEL_ITEMS], total;
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; k++) {
      const int t = t0 + SEL_ITEMS * threadIdx.x + k;
      f[k] = 0;
      if (t < n) {
        const int Zi = Z[first + t];
#pragma unroll
        for (int q = 0; q < SOAP_SPECIES_CAP; q++)
          if (q < nZ && Zi >= 0 && (cz.Z[q] == Zi || cz.Z[q] == 0)) f[k] = 1;
      }
    }
    BS(tmp).ExclusiveSum(f, pos, total);
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; k++)
      if (f[k]) centres[base + pos[k]] = first + t0 + SEL_ITEMS * threadIdx.x + k;
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count_out = base;
}

// ---------------------------------------------------------------------------------------------------
// velocity Verlet (advance_verlet1 / advance_verlet2, src/libAtoms/DynamicalSystem.f95:1814-2132, 2159-2383; plain atoms,
// no thermostat / barostat / constraints):  v += a dt/2 ; x += v dt   |   a = f/m ; v += a dt/2
// ---------------------------------------------------------------------------------------------------
__global__ void k_verlet1(int N, double dt, double* __restrict__ pos, double* __restrict__ velo, const double* __restrict__ acc) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 3 * N) return;
  double v = velo[k] + 0.5 * acc[k] * dt;
  velo[k] = v;
  pos[k] = pos[k] + v * dt;
}
__global__ void k_verlet2(int N, double dt, const double* __restrict__ force, const double* __restrict__ mass, const double* __restrict__ velo_in,
                          double* __restrict__ velo_out, double* __restrict__ acc, int half_kick) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 3 * N) return;
  double a = force[k] / mass[k / 3];
  acc[k] = a;
  velo_out[k] = half_kick ? velo_in[k] + 0.5 * a * dt : velo_in[k];
}
// kinetic energy partials: sum_i m_i |v_i|^2 / 2 (kinetic_energy, DynamicalSystem.f95:1351)
__global__ void __launch_bounds__(256) k_kinetic(int N, const double* __restrict__ mass, const double* __restrict__ velo, double* __restrict__ part) {
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  double t = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256)
    t += 0.5 * mass[i] * (velo[3 * i] * velo[3 * i] + velo[3 * i + 1] * velo[3 * i + 1] + velo[3 * i + 2] * velo[3 * i + 2]);
  t = BR(tmp).Sum(t);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// ---- deterministic force scatter -----------------------------------------------------------------------------------------
// The SOAP adjoint kernels store the force a pair exerts on its neighbour at the SLOT of that list entry (fpair) instead of adding
// it to force[j] with an FP64 atomic; here the slots are indexed by receiving atom (stable radix sort of (j, slot)), and one warp per
// atom adds its contributions in slot order with a fixed reduction tree, plus the atom's own sum as a centre (fself).
__global__ void k_det_keys(long n_slots, int row_cap, int first, const int* __restrict__ nbr_end, const int* __restrict__ nbr_j, int N,
                           int* __restrict__ keys, int* __restrict__ vals) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_slots) return;
  bool valid = true;
  if (row_cap > 0) valid = p < (long)nbr_end[first + (int)(p / row_cap)];  // fixed-capacity rows: slots beyond the row's fill are empty
  keys[p] = valid ? nbr_j[p] : N;
  vals[p] = (int)p;
}
__global__ void k_det_joff(const int* __restrict__ keys_sorted, long n_slots, int N, int* __restrict__ joff /* [N + 2] */) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_slots) return;
  const int key = keys_sorted[p], prev = p > 0 ? keys_sorted[p - 1] : -1;
  for (int c = prev + 1; c <= key; c++) joff[c] = (int)p;
  if (p == n_slots - 1)
    for (int c = key + 1; c <= N + 1; c++) joff[c] = (int)n_slots;
}
__global__ void __launch_bounds__(128) k_det_gather(int N, const int* __restrict__ joff, const int* __restrict__ slot_of_entry,
                                                    const double* __restrict__ fpair, double* __restrict__ fself, double* __restrict__ force) {
  const int lane = threadIdx.x & 31, j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= N) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int e = joff[j] + lane; e < joff[j + 1]; e += 32) {
    const size_t p = (size_t)slot_of_entry[e];
    a0 += fpair[3 * p]; a1 += fpair[3 * p + 1]; a2 += fpair[3 * p + 2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if (lane == 0) {
    force[3 * (size_t)j] += a0 + fself[3 * (size_t)j];
    force[3 * (size_t)j + 1] += a1 + fself[3 * (size_t)j + 1];
    force[3 * (size_t)j + 2] += a2 + fself[3 * (size_t)j + 2];
    fself[3 * (size_t)j] = fself[3 * (size_t)j + 1] = fself[3 * (size_t)j + 2] = 0.0;  // clean for the next coordinate
  }
}

// largest squared displacement since the last list build, per block (the host takes the maximum of the block values)
constexpr int DISP_BLOCKS = 256;
__global__ void __launch_bounds__(256) k_max_disp2(const double* __restrict__ pos, const double* __restrict__ last, int N, double* __restrict__ part) {
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  double m = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
    const double dx = pos[3 * (size_t)i] - last[3 * (size_t)i], dy = pos[3 * (size_t)i + 1] - last[3 * (size_t)i + 1], dz = pos[3 * (size_t)i + 2] - last[3 * (size_t)i + 2];
    const double d2 = dx * dx + dy * dy + dz * dz;
    m = (d2 > m || d2 != d2) ? d2 : m;  // a NaN position forces the rebuild (and its error reporting)
  }
  m = BR(tmp).Reduce(m, [](double a, double b) { return (a != a || a > b) ? a : b; });
  if (threadIdx.x == 0) part[blockIdx.x] = m;
}

__global__ void k_iota(int* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// stage slots of gap_potential_last_timings
enum { ST_CONNECT = 0, ST_SOAP_FWD = 1, ST_COV_GEMM1 = 2, ST_COV_GEMM2 = 3, ST_SOAP_ADJ = 4, ST_PAIR2B = 5, ST_OTHER = 6, ST_TOTAL = 7 };

void mark(gap_potential* P, cudaStream_t st, int stage) {
  if (!P->timing) return;  // stage timing is instrumentation: off unless gap_potential_set_timing asked for it
  // level 2: only the events that bracket the covariance GEMMs (end of the SOAP forward stage, end of GEMM-1, end of GEMM-2)
  if (P->timing == 2 && stage != ST_SOAP_FWD && stage != ST_COV_GEMM1 && stage != ST_COV_GEMM2) return;
  if (P->ev_used == P->ev.size()) {
    cudaEvent_t e;
    CUDA_OK(cudaEventCreate(&e));
    P->ev.push_back(e);
    P->ev_stage.push_back(0);
  }
  P->ev_stage[P->ev_used] = stage;
  CUDA_OK(cudaEventRecord(P->ev[P->ev_used], st));
  P->ev_used++;
}
void collect_timings(gap_potential* P) {
  for (double& v : P->last_ms) v = 0.0;
  if (P->ev_used < 2) return;
  for (size_t k = 1; k < P->ev_used; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, P->ev[k - 1], P->ev[k]) == cudaSuccess) {
      int st = P->ev_stage[k];
      if (P->timing == 2 && st != ST_COV_GEMM1 && st != ST_COV_GEMM2) continue;  // intervals between the brackets are not stages
      if (st >= 0 && st < 7) P->last_ms[st] += ms;
    }
  }
  float tot = 0.f;
  if (cudaEventElapsedTime(&tot, P->ev[0], P->ev[P->ev_used - 1]) == cudaSuccess) P->last_ms[7] = tot;
}

// ---------------------------------------------------------------------------------------------------
// geometry on the host: inverse lattice, cell grid
// ---------------------------------------------------------------------------------------------------
inline double det3(const double* a) {
  return a[0] * (a[4] * a[8] - a[7] * a[5]) - a[3] * (a[1] * a[8] - a[7] * a[2]) + a[6] * (a[1] * a[5] - a[4] * a[2]);
}
void inv3(const double* a, double* g) {  // column-major both
  double det = det3(a);
  auto A = [&](int r, int c) { return a[r + 3 * c]; };
  g[0 + 3 * 0] = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) / det;
  g[0 + 3 * 1] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) / det;
  g[0 + 3 * 2] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) / det;
  g[1 + 3 * 0] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) / det;
  g[1 + 3 * 1] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) / det;
  g[1 + 3 * 2] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) / det;
  g[2 + 3 * 0] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) / det;
  g[2 + 3 * 1] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) / det;
  g[2 + 3 * 2] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) / det;
}

// the geometry a list was built for (this%last_connect_pos / last_connect_lattice / last_connect_cutoff, Connection.f95:1122-1124)
void remember_build(gap_potential* P, int N, int first, int last, const double* d_pos, const double* lattice, const int* pbc, double cutoff_used,
                    cudaStream_t st) {
  P->b_lastpos.ensure(sizeof(double) * 3 * (size_t)N);
  CUDA_OK(cudaMemcpyAsync(P->b_lastpos.p, d_pos, sizeof(double) * 3 * (size_t)N, cudaMemcpyDeviceToDevice, st));
  for (int k = 0; k < 9; k++) P->last_lat[k] = lattice[k];
  for (int k = 0; k < 3; k++) P->last_pbc[k] = pbc[k] ? 1 : 0;
  P->last_N = N; P->last_first = first; P->last_last = last; P->last_cut = cutoff_used;
  P->list_valid = true;
  P->n_rebuilds++;
}

// speculative = true: if the previous call had the same (N, first, last), build the list in ONE pass into rows of fixed
// capacity (largest row of that call + 25 %) and do NOT wait for anything: the largest row of this call arrives in
// P->h_pin and verify_connect() checks it after the caller's final synchronisation.  Otherwise: exact packed CSR rows
// (count, scan, one synchronisation for the entry count, fill).  Sets P->cv_off / cv_end / cv_j / cv_s.
void build_connect(gap_potential* P, int N, int first, int last, const double* d_pos, const double* lattice, const int* pbc, double cutoff,
                   bool want_dist, bool speculative, cudaStream_t st) {
  P->pending_check = false;
  P->list_reused = false;
  if (N < 0) throw GapError("calc_connect: negative number of atoms");
  if (cutoff < 0.0) throw GapError("calc_connect: Negative cutoff radius " + std::to_string(cutoff));  // Connection.f95:1069
  // calc_connect with cutoff_skin (Connection.f95:1085-1128): the list is built out to cutoff + skin and kept while no atom has moved
  // more than skin / 2 since the build (same N, partition, lattice and pbc); the descriptor kernels recompute every distance from the
  // current positions and drop pairs beyond their own cutoff (descriptors.f95:8190, 4729), as calc_dists + the reference's filters do
  const bool use_skin = speculative && !want_dist && P->cutoff_skin > 0.0 && N > 0 && cutoff > 0.0;
  if (use_skin) {
    cutoff += P->cutoff_skin;
    bool same = P->list_valid && P->last_N == N && P->last_first == first && P->last_last == last && cutoff <= P->last_cut;
    for (int k = 0; k < 9 && same; k++) same = P->last_lat[k] == lattice[k];
    for (int k = 0; k < 3 && same; k++) same = P->last_pbc[k] == (pbc[k] ? 1 : 0);
    if (same) {
      P->b_disp.ensure(sizeof(double) * DISP_BLOCKS);
      if (!P->h_disp) CUDA_OK(cudaHostAlloc((void**)&P->h_disp, sizeof(double) * DISP_BLOCKS, cudaHostAllocDefault));
      const int nb = std::min(DISP_BLOCKS, (N + 255) / 256);
      k_max_disp2<<<nb, 256, 0, st>>>(d_pos, P->b_lastpos.as<double>(), N, P->b_disp.as<double>());
      P->launches += 1;
      CUDA_OK(cudaMemcpyAsync(P->h_disp, P->b_disp.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      double m = 0.0;
      for (int k = 0; k < nb; k++) m = (P->h_disp[k] != P->h_disp[k] || P->h_disp[k] > m) ? P->h_disp[k] : m;
      if (m == m && std::sqrt(m) < 0.5 * P->cutoff_skin) {  // :1110-1116: reuse (cv_* still describe the list)
        P->n_reuses++;
        P->list_reused = true;
        return;
      }
    }
    P->list_valid = false;
  }
  P->conn_N = N;
  P->conn_nnz = 0;
  P->b_off.ensure(sizeof(int) * (N + 4));  // [0..N] row offsets, then: [N+1] error flag, [N+2] largest row
  P->cv_off = P->b_off.as<int>(); P->cv_end = P->b_off.as<int>() + 1; P->cv_j = P->b_j.as<int>(); P->cv_s = P->b_s.as<int>();
  if (N == 0 || cutoff == 0.0) {  // cutoff == 0: "don't compute neighbours" (:1079)
    P->list_row_cap = 0;
    P->list_slots = 0;
    CUDA_OK(cudaMemsetAsync(P->b_off.p, 0, sizeof(int) * (N + 4), st));
    return;
  }
  CellGrid grid;
  memset(&grid, 0, sizeof(grid));
  grid.cutoff = cutoff;
  double lat[9];
  for (int k = 0; k < 9; k++) {
    lat[k] = lattice[k];
    if (!std::isfinite(lat[k])) throw GapError("calc_connect: lattice is not finite");
  }
  for (int k = 0; k < 3; k++) grid.pbc[k] = pbc[k] ? 1 : 0;
  // non-periodic directions with a missing (zero) cell vector get a unit vector that completes the basis; it never
  // multiplies a non-zero shift
  for (int k = 0; k < 3; k++) {
    double nrm = std::sqrt(lat[3 * k] * lat[3 * k] + lat[3 * k + 1] * lat[3 * k + 1] + lat[3 * k + 2] * lat[3 * k + 2]);
    if (nrm < 1e-12) {
      if (grid.pbc[k]) throw GapError("calc_connect: periodic direction with zero-length lattice vector");
      lat[3 * k] = lat[3 * k + 1] = lat[3 * k + 2] = 0.0;
    }
  }
  auto col_norm = [&](int k) { return std::sqrt(lat[3 * k] * lat[3 * k] + lat[3 * k + 1] * lat[3 * k + 1] + lat[3 * k + 2] * lat[3 * k + 2]); };
  for (int k = 0; k < 3; k++) {
    if (col_norm(k) >= 1e-12) continue;
    double best = -1.0;
    int bi = 0;
    for (int c = 0; c < 3; c++) {
      double trial[9];
      memcpy(trial, lat, sizeof(trial));
      trial[3 * k + c] = 1.0;
      // temporarily fill the other empty columns so that det is meaningful
      for (int k2 = 0; k2 < 3; k2++)
        if (k2 != k && std::sqrt(trial[3 * k2] * trial[3 * k2] + trial[3 * k2 + 1] * trial[3 * k2 + 1] + trial[3 * k2 + 2] * trial[3 * k2 + 2]) < 1e-12)
          trial[3 * k2 + ((c + 1 + (k2 > k ? 1 : 0)) % 3)] = 1.0;
      double dv = std::fabs(det3(trial));
      if (dv > best) { best = dv; bi = c; }
    }
    lat[3 * k + bi] = 1.0;
  }
  double det = det3(lat);
  double scale = col_norm(0) * col_norm(1) * col_norm(2);
  if (!(std::fabs(det) > 1e-10 * scale)) throw GapError("calc_connect: singular lattice");
  memcpy(grid.lat, lat, sizeof(lat));
  inv3(lat, grid.g);
  // heights of the cell along each reciprocal direction: V / |b x c| etc. = 1 / |row k of g|
  double height[3];
  for (int k = 0; k < 3; k++) height[k] = 1.0 / std::sqrt(grid.g[k] * grid.g[k] + grid.g[k + 3] * grid.g[k + 3] + grid.g[k + 6] * grid.g[k + 6]);

  P->b_minmax.ensure(sizeof(double) * (6 + 6 * 64));
  int launches = 0;
  double mm[6] = {0, 0, 0, 0, 0, 0};
  if (!grid.pbc[0] || !grid.pbc[1] || !grid.pbc[2]) {
    launch_frac_minmax(d_pos, N, nullptr, grid, P->b_minmax.as<double>(), st, &launches);
    CUDA_OK(cudaMemcpyAsync(mm, P->b_minmax.p, sizeof(mm), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int k = 0; k < 6; k++)
      if (!std::isfinite(mm[k])) throw GapError("calc_connect: atomic positions are not finite");
  }
  double width[3];
  for (int k = 0; k < 3; k++) {
    if (grid.pbc[k]) {
      int n = (int)std::floor(height[k] / cutoff);
      grid.n[k] = n < 1 ? 1 : n;
    } else {
      double ext = mm[3 + k] - mm[k];
      grid.toff[k] = mm[k];
      grid.tscale[k] = ext > 0 ? 1.0 / ext : 0.0;
      int n = (int)std::floor(ext * height[k] / cutoff);
      grid.n[k] = n < 1 ? 1 : n;
      height[k] = ext * height[k];
    }
  }
  // bound the number of cells by the number of atoms (dilute systems / vacuum)
  {
    double cap = 4.0 * N + 1024.0;
    while ((double)grid.n[0] * grid.n[1] * grid.n[2] > cap) {
      int kmax = 0;
      for (int k = 1; k < 3; k++)
        if (grid.n[k] > grid.n[kmax]) kmax = k;
      grid.n[kmax] = (grid.n[kmax] + 1) / 2;
    }
  }
  for (int k = 0; k < 3; k++) {
    width[k] = height[k] / grid.n[k];
    if (grid.pbc[k]) {
      grid.R[k] = (int)std::ceil(cutoff / width[k] + 1e-9);
      if (grid.R[k] < 1) grid.R[k] = 1;
      if ((grid.R[k] + grid.n[k] - 1) / grid.n[k] + 2 > 60) throw GapError("calc_connect: cutoff is too large for this cell (more than 60 images)");
    } else {
      grid.R[k] = grid.n[k] > 1 ? (int)std::ceil(cutoff / width[k] + 1e-9) : 0;
      if (grid.R[k] > grid.n[k] - 1) grid.R[k] = grid.n[k] - 1;
    }
  }
  const int ncell = grid.n[0] * grid.n[1] * grid.n[2];

  NeighbourWork& w = P->nw;
  P->b_cell_of.ensure(sizeof(int) * N); w.cell_of = P->b_cell_of.as<int>();
  P->b_mshift.ensure(sizeof(int) * N); w.mshift = P->b_mshift.as<int>();
  P->b_keys.ensure(sizeof(int) * N); w.sort_keys = P->b_keys.as<int>();
  P->b_idx.ensure(sizeof(int) * N); w.sort_idx = P->b_idx.as<int>();
  P->b_slot.ensure(sizeof(int) * N); w.slot_of = P->b_slot.as<int>();
  P->b_iota.ensure(sizeof(int) * N); w.iota = P->b_iota.as<int>();
  P->b_cstart.ensure(sizeof(int) * (ncell + 2)); w.cell_start = P->b_cstart.as<int>();
  {  // cell counts of the counting sort: zeroed when (re)allocated; every build hands them back zeroed (k_place)
    const void* before = P->b_ccount.p;
    P->b_ccount.ensure(sizeof(int) * (ncell + 2));
    if (P->b_ccount.p != before) CUDA_OK(cudaMemsetAsync(P->b_ccount.p, 0, P->b_ccount.cap, st));
    w.cell_count = P->b_ccount.as<int>();
  }
  P->b_spos.ensure(sizeof(double) * 3 * N); w.spos = P->b_spos.as<double>();
  P->b_smshift.ensure(sizeof(int) * N); w.smshift = P->b_smshift.as<int>();
  P->b_nn.ensure(sizeof(int) * (N + 2)); w.nn = P->b_nn.as<int>();
  size_t cb = neighbour_cub_bytes(N, ncell);
  P->b_cub.ensure(cb); w.cub_tmp = P->b_cub.p; w.cub_bytes = P->b_cub.cap;

  int* const stat = P->b_off.as<int>() + N;  // [0] entry count (written by the scan of the exact layout) [1] error flag [2] largest row
  w.err_flag = stat + 1;
  const bool spec = speculative && !want_dist && P->row_hint >= 0 && P->hint_N == N && P->hint_first == first && P->hint_last == last;
  if (!(spec && P->stat_clean)) CUDA_OK(cudaMemsetAsync(stat, 0, 3 * sizeof(int), st));  // (k_finalize of the previous evaluation left them clean)
  P->stat_clean = false;
  launch_bin_atoms(d_pos, N, grid, ncell, w, st, &launches);
  if (spec) {
    const int row_cap = round_up(P->row_hint + P->row_hint / 4 + 8, 4);
    const long cap = (long)row_cap * (last - first);
    if (cap <= 2147483000L) {
      P->b_end.ensure(sizeof(int) * (N + 1));
      P->b_j.ensure(sizeof(int) * (size_t)(cap + 1));
      P->b_s.ensure(sizeof(int) * (size_t)(cap + 1));
      launch_neigh_onepass(d_pos, N, first, last, grid, w, P->b_off.as<int>(), P->b_end.as<int>(), P->b_j.as<int>(), P->b_s.as<int>(), row_cap,
                           stat + 2, st, &launches);
      P->pending_check = true;  // the status words reach the host through k_finalize (mapped memory), see verify_connect
      P->pending_cap = row_cap;
      P->list_row_cap = row_cap;
      P->list_slots = cap;
      P->cv_end = P->b_end.as<int>(); P->cv_j = P->b_j.as<int>(); P->cv_s = P->b_s.as<int>();
      P->launches += launches;
      if (use_skin) remember_build(P, N, first, last, d_pos, lattice, pbc, cutoff, st);
      CUDA_OK(cudaGetLastError());
      return;
    }
  }
  launch_neigh_count(d_pos, N, first, last, grid, w, P->b_off.as<int>(), stat + 2, st, &launches);
  CUDA_OK(cudaMemcpyAsync(P->h_pin, stat, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  const int nnz = P->h_pin[0];
  if (P->h_pin[1]) throw GapError("calc_connect: an atom lies more than 30 periodic images away from the cell; wrap the positions first");
  if (nnz < 0) throw GapError("calc_connect: neighbour list exceeds 2^31 entries");
  P->conn_nnz = nnz;
  P->list_row_cap = 0;
  P->list_slots = nnz;
  P->row_hint = P->h_pin[2]; P->hint_N = N; P->hint_first = first; P->hint_last = last;
  P->b_j.ensure(sizeof(int) * (size_t)(nnz + 1));
  P->b_s.ensure(sizeof(int) * (size_t)(nnz + 1));
  if (want_dist) P->b_d.ensure(sizeof(double) * (size_t)(nnz + 1));
  launch_neigh_fill(d_pos, N, first, last, grid, w, P->b_off.as<int>(), P->b_j.as<int>(), P->b_s.as<int>(), want_dist ? P->b_d.as<double>() : nullptr,
                    nnz, st, &launches);
  P->cv_j = P->b_j.as<int>(); P->cv_s = P->b_s.as<int>();
  P->launches += launches;
  if (use_skin) remember_build(P, N, first, last, d_pos, lattice, pbc, cutoff, st);
  CUDA_OK(cudaGetLastError());
}

// After the stream has been synchronised: did the speculatively sized list hold every entry?  false = repeat the call.
bool verify_connect(gap_potential* P) {
  if (!P->pending_check) return true;
  P->pending_check = false;
  if (P->h_pin[1]) throw GapError("calc_connect: an atom lies more than 30 periodic images away from the cell; wrap the positions first");
  const int max_row = P->h_pin[2];
  const bool ok = max_row <= P->pending_cap;
  P->row_hint = ok ? max_row : -1;  // overflow: the repeat takes the exact (synchronising) path
  if (!ok) P->list_valid = false;   // (and never reuses the truncated list)
  return ok;
}

// ---------------------------------------------------------------------------------------------------
// model upload
// ---------------------------------------------------------------------------------------------------
double factorial_d(int n) {
  double f = 1.0;
  for (int i = 2; i <= n; i++) f *= i;
  return f;
}

void upload_model(gap_potential* P) {
  const double pi = 3.14159265358979323846264338327950288;
  CUDA_OK(cudaMalloc(&P->d_fin_counter, sizeof(unsigned int)));
  CUDA_OK(cudaMemset(P->d_fin_counter, 0, sizeof(unsigned int)));
  CUDA_OK(cudaHostAlloc((void**)&P->h_pin, 4 * sizeof(int), cudaHostAllocMapped));
  CUDA_OK(cudaHostGetDevicePointer((void**)&P->d_hpin, P->h_pin, 0));
  P->h_pin[0] = P->h_pin[1] = P->h_pin[2] = P->h_pin[3] = 0;
  CUDA_OK(cudaMalloc(&P->d_e0, sizeof(double) * 128));
  CUDA_OK(cudaMemcpy(P->d_e0, P->model.e0, sizeof(double) * 128, cudaMemcpyHostToDevice));
  for (const Coordinate& c : P->model.coord) {
    CoordDev cd;
    cd.kind = c.kind;
    cd.delta = c.delta;
    cd.f0 = c.f0;
    if (c.kind == DESC_SOAP) {
      const SoapSpec& s = c.soap;
      if (s.n_max > SOAP_NMAX_CAP) throw GapError("soap n_max > " + std::to_string(SOAP_NMAX_CAP) + " is not supported by the B200 path");
      if (s.l_max > SOAP_LMAX_CAP) throw GapError("soap l_max > " + std::to_string(SOAP_LMAX_CAP) + " is not supported by the B200 path");
      if (s.n_species > SOAP_SPECIES_CAP || s.n_Z > SOAP_SPECIES_CAP)
        throw GapError("soap n_species/n_Z > " + std::to_string(SOAP_SPECIES_CAP) + " is not supported by the B200 path");
      SoapDev& h = cd.h;
      memset(&h, 0, sizeof(h));
      h.cutoff = s.cutoff; h.ctw = s.cutoff_transition_width; h.alpha = s.alpha; h.central_weight = s.central_weight;
      h.sigma0 = s.covariance_sigma0; h.chol00 = s.cholesky_overlap[0];
      h.cutoff_scale = s.cutoff_scale; h.cutoff_rate = s.cutoff_rate; h.cutoff_dexp = s.cutoff_dexp;
      h.norm_radial_decay = (s.cutoff_dexp > 0 && s.cutoff_rate != 0.0) ? s.cutoff_rate / (1.0 + s.cutoff_rate) : 1.0;
      h.l_max = s.l_max; h.n_max = s.n_max; h.n_species = s.n_species; h.n_Z = s.n_Z; h.K1 = s.K1(); h.d = s.d;
      h.d_pad = round_up(s.d, COV_BK); h.nlm = (s.l_max + 1) * (s.l_max + 1);
      h.normalise = s.normalise; h.cras = s.central_reference_all_species; h.two_lp1 = s.do_two_l_plus_one;
      for (int k = 0; k < s.n_species; k++) h.species_Z[k] = s.species_Z[k];
      for (int k = 0; k < s.n_Z; k++) h.centre_Z[k] = s.Z[k];
      for (int k = 0; k < s.n_max; k++) h.r_basis[k] = s.r_basis[k];
      for (int k = 0; k < s.n_max * s.n_max; k++) h.T[k] = s.transform_basis[k];
      for (int l = 0; l <= s.l_max; l++) {
        h.tlpo[l] = s.do_two_l_plus_one ? 1.0 / std::sqrt(2.0 * l + 1.0) : 1.0;
        for (int m = 0; m <= l; m++)
          h.ynorm[l * (l + 1) / 2 + m] = std::sqrt((2.0 * l + 1.0) / (4.0 * pi) * factorial_d(l - m) / factorial_d(l + m)) * (m > 0 ? std::sqrt(2.0) : 1.0);
      }
      cudaDeviceProp prop;
      CUDA_OK(cudaGetDeviceProperties(&prop, P->device));
      P->n_sm = prop.multiProcessorCount;
      if ((!s.general || (!s.global && s.radial_basis == "EQUISPACED_GAUSS")) &&
          (soap_adjoint_smem(h) > prop.sharedMemPerBlockOptin || soap_forward_smem(h) > prop.sharedMemPerBlockOptin))
        throw GapError("soap descriptor too large for the shared memory of this device");
      CUDA_OK(cudaMalloc(&cd.d_sp, sizeof(SoapDev)));
      CUDA_OK(cudaMemcpy(cd.d_sp, &h, sizeof(SoapDev), cudaMemcpyHostToDevice));
      memset(&cd.gen, 0, sizeof(cd.gen));
      cd.general = s.general;
      cd.hybrid = s.general && !s.global && s.radial_basis == "EQUISPACED_GAUSS" && s.n_grid == s.n_max && getenv("GAP_B200_SOAP_HYBRID") == nullptr;
      if (s.general && !s.global && s.radial_basis != "EQUISPACED_GAUSS" && getenv("GAP_B200_SOAP_HYBRID") == nullptr) {
        const int G = s.n_grid;  // = 3 n_max: always divisible by 3, and G / 3 = n_max <= 16
        const int npass = G <= SOAP_NMAX_CAP ? 1 : ((G % 2 == 0 && G / 2 <= SOAP_NMAX_CAP) ? 2 : 3);
        if (G % npass == 0 && G / npass <= SOAP_NMAX_CAP) {
          SoapDev hg = h;
          const int gp = G / npass;
          hg.n_max = gp; hg.K1 = s.n_species * gp; hg.central_weight = 0.0; hg.chol00 = 0.0;
          memset(hg.T, 0, sizeof(hg.T));
          for (int a = 0; a < gp; a++) hg.T[a + gp * a] = 1.0;
          if (soap_adjoint_smem(hg) <= prop.sharedMemPerBlockOptin && soap_forward_smem(hg) <= prop.sharedMemPerBlockOptin) {
            cd.grid_passes = npass; cd.grid_gp = gp; cd.h_grid = hg;
            for (int p = 0; p < npass; p++) {
              for (int a = 0; a < gp; a++) hg.r_basis[a] = s.r_grid[(size_t)p * gp + a];
              CUDA_OK(cudaMalloc(&cd.d_sp_grid[p], sizeof(SoapDev)));
              CUDA_OK(cudaMemcpy(cd.d_sp_grid[p], &hg, sizeof(SoapDev), cudaMemcpyHostToDevice));
            }
          }
        }
      }
      if (s.general) {  // tables of the general path, packed into one allocation (doubles first, then the two int lists)
        const size_t np = s.pair_ia.size();
        std::vector<double> blob;
        auto put = [&](const std::vector<double>& v) { size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); if (blob.size() & 1) blob.push_back(0.0); return o; };
        const size_t o_r = put(s.r_grid), o_P = put(s.P), o_c0 = put(s.c0), o_W1 = put(s.W1), o_W2 = put(s.W2), o_f = put(s.pair_fac), o_int = blob.size();
        // int tables: ia | jb | elements grouped by ia (offsets, list) | elements grouped by jb (offsets, list), each group in element order
        std::vector<int> ints;
        ints.insert(ints.end(), s.pair_ia.begin(), s.pair_ia.end());
        ints.insert(ints.end(), s.pair_jb.begin(), s.pair_jb.end());
        size_t o_ia_off = 0, o_ia = 0, o_jb_off = 0, o_jb = 0;
        for (int side = 0; side < 2; side++) {
          const std::vector<int>& key = side ? s.pair_jb : s.pair_ia;
          const int K = side ? s.Kb : s.Ka;
          (side ? o_jb_off : o_ia_off) = ints.size();
          std::vector<int> off((size_t)K + 1, 0);
          for (size_t e = 0; e < np; e++) off[(size_t)key[e] + 1]++;
          for (int k = 0; k < K; k++) off[(size_t)k + 1] += off[k];
          ints.insert(ints.end(), off.begin(), off.end());
          (side ? o_jb : o_ia) = ints.size();
          std::vector<int> lst(np), fill(off.begin(), off.end() - 1);
          for (size_t e = 0; e < np; e++) lst[(size_t)fill[key[e]]++] = (int)e;
          ints.insert(ints.end(), lst.begin(), lst.end());
        }
        const size_t bytes = blob.size() * sizeof(double) + ints.size() * sizeof(int);
        CUDA_OK(cudaMalloc(&cd.gen_blob, bytes + 16));
        CUDA_OK(cudaMemcpy(cd.gen_blob, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice));
        int* d_int = (int*)((double*)cd.gen_blob + o_int);
        CUDA_OK(cudaMemcpy(d_int, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice));
        const double* b = (const double*)cd.gen_blob;
        cd.gen.n_grid = s.n_grid; cd.gen.Ka = s.Ka; cd.gen.Kb = s.Kb; cd.gen.n_pairs = (int)np;
        cd.gen.r_grid = b + o_r; cd.gen.P = b + o_P; cd.gen.c0 = b + o_c0; cd.gen.W1 = b + o_W1; cd.gen.W2 = b + o_W2; cd.gen.pair_fac = b + o_f;
        cd.gen.pair_ia = d_int; cd.gen.pair_jb = d_int + np;
        cd.gen.by_ia_off = d_int + o_ia_off; cd.gen.by_ia = d_int + o_ia; cd.gen.by_jb_off = d_int + o_jb_off; cd.gen.by_jb = d_int + o_jb;
        cd.gen.global_mode = s.global ? 1 : 0;
        if (s.global) {  // average=T: the summed density expansion and the shared dE/dX on the radial grid
          CUDA_OK(cudaMalloc(&cd.gen_global, sizeof(double) * (size_t)h.nlm * (h.K1 + s.n_species * s.n_grid)));
          cd.gen.Xg = cd.gen_global;
          cd.gen.Lt = cd.gen_global + (size_t)h.nlm * h.K1;
        }
        if (soap_general_smem(h, cd.gen) > prop.sharedMemPerBlockOptin)
          throw GapError("soap descriptor (general path) too large for the shared memory of this device");
      }
      cd.M = c.M;
      cd.M_pad = round_up(c.M > 0 ? c.M : 1, COV_BK) + COV_BN1_MAX;  // any GEMM-1 column tiling (cov_gemm1_bn) stays in bounds
      cd.d_pad = h.d_pad;
      cd.bn2 = cov_gemm2_bn(s.d);
      cd.dn_pad = round_up(s.d, cd.bn2);
      // sparse points: rows [M_pad][d_pad] (GEMM-1 B operand) and transposed [dn_pad][M_pad] (GEMM-2 B operand)
      std::vector<double> rows((size_t)cd.M_pad * cd.d_pad, 0.0), trn((size_t)cd.dn_pad * cd.M_pad, 0.0), al(cd.M_pad, 0.0), cu(cd.M_pad, 0.0);
      for (int m = 0; m < c.M; m++) {
        for (int q = 0; q < s.d; q++) {
          double v = c.sparseX[(size_t)m * s.d + q];
          rows[(size_t)m * cd.d_pad + q] = v;
          trn[(size_t)q * cd.M_pad + m] = v;
        }
        al[m] = c.alpha[m] * c.sparseCutoff[m] * (c.delta * c.delta);  // GP weight of the fused epilogue (0 in the padding)
        cu[m] = c.sparseCutoff[m];
      }
      CUDA_OK(cudaMalloc(&cd.sp_rows, rows.size() * sizeof(double)));
      CUDA_OK(cudaMalloc(&cd.st_rows, trn.size() * sizeof(double)));
      CUDA_OK(cudaMalloc(&cd.alpha, al.size() * sizeof(double)));
      CUDA_OK(cudaMalloc(&cd.scut, cu.size() * sizeof(double)));
      CUDA_OK(cudaMemcpy(cd.sp_rows, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(cd.st_rows, trn.data(), trn.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(cd.alpha, al.data(), al.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(cd.scut, cu.data(), cu.size() * sizeof(double), cudaMemcpyHostToDevice));
      cd.cp.delta2 = c.delta * c.delta;
      cd.cp.zeta = c.zeta;
      double zi = std::nearbyint(c.zeta);
      cd.cp.zeta_int = (std::fabs(c.zeta - zi) < 1e-12 * std::fmax(1.0, std::fabs(c.zeta)) && zi >= 0 && zi <= 64) ? (int)zi : -1;
    } else if (c.kind == DESC_ANGLE_3B) {
      std::vector<double> tb((size_t)4 * (c.M > 0 ? c.M : 1), 0.0);
      double ef0 = 0.0;
      for (int m = 0; m < c.M; m++) {  // sparseX / theta as gpCoordinates_precalculate_sparse does (gp_predict.f95:3901-3923)
        for (int q = 0; q < 3; q++) tb[(size_t)4 * m + q] = c.sparseX[(size_t)m * 3 + q] / c.theta[q];
        tb[(size_t)4 * m + 3] = c.alpha[m] * c.sparseCutoff[m] * c.delta * c.delta;
        ef0 += c.alpha[m] * c.sparseCutoff[m];
      }
      CUDA_OK(cudaMalloc(&cd.t3, tb.size() * sizeof(double)));
      CUDA_OK(cudaMemcpy(cd.t3, tb.data(), tb.size() * sizeof(double), cudaMemcpyHostToDevice));
      memset(&cd.p3, 0, sizeof(cd.p3));
      cd.p3.cutoff = c.a3b.cutoff; cd.p3.ctw = c.a3b.cutoff_transition_width;
      for (int q = 0; q < 3; q++) cd.p3.inv_theta[q] = 1.0 / c.theta[q];
      cd.p3.e_f0 = c.f0 * c.f0 * ef0;
      cd.p3.Zc = c.a3b.Zc; cd.p3.Z1 = c.a3b.Z1; cd.p3.Z2 = c.a3b.Z2; cd.p3.M = c.M;
      cd.p3.table = cd.t3;
    } else {
      const int ne = (int)c.d2b.exponents.size();
      std::vector<double> xs((size_t)(c.M > 0 ? c.M : 1) * ne, 0.0), al(c.M > 0 ? c.M : 1, 0.0), cu(al.size(), 0.0);
      for (int m = 0; m < c.M; m++) {
        for (int q = 0; q < ne; q++) xs[(size_t)m * ne + q] = c.sparseX[(size_t)m * ne + q];
        al[m] = c.alpha[m]; cu[m] = c.sparseCutoff[m];
      }
      CUDA_OK(cudaMalloc(&cd.x2, xs.size() * sizeof(double)));
      CUDA_OK(cudaMalloc(&cd.a2, al.size() * sizeof(double)));
      CUDA_OK(cudaMalloc(&cd.c2, al.size() * sizeof(double)));
      CUDA_OK(cudaMemcpy(cd.x2, xs.data(), xs.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(cd.a2, al.data(), al.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(cd.c2, cu.data(), al.size() * sizeof(double), cudaMemcpyHostToDevice));
      memset(&cd.p2, 0, sizeof(cd.p2));
      cd.p2.cutoff = c.d2b.cutoff; cd.p2.ctw = c.d2b.cutoff_transition_width; cd.p2.delta2 = c.delta * c.delta; cd.p2.f02 = c.f0 * c.f0;
      cd.p2.n_exp = ne;
      for (int q = 0; q < ne; q++) { cd.p2.inv_theta[q] = 1.0 / c.theta[q]; cd.p2.exponents[q] = c.d2b.exponents[q]; }
      cd.p2.tail_exponent = c.d2b.tail_exponent; cd.p2.tail_range = c.d2b.tail_range; cd.p2.intra_mode = c.d2b.intra_mode; cd.p2.resid = nullptr;
      cd.p2.Z1 = c.d2b.Z1; cd.p2.Z2 = c.d2b.Z2; cd.p2.M = c.M;
      cd.p2.sparseX = cd.x2; cd.p2.alpha = cd.a2; cd.p2.scut = cd.c2;
    }
    P->cd.push_back(cd);
  }
}

gap_potential* create_potential(const std::string& args, const std::string& xml, const std::string& base_dir, int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw GapError(std::string("gap_potential_initialise: no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (device < 0 || device >= ndev) throw GapError("gap_potential_initialise: CUDA device " + std::to_string(device) + " out of range");
  GapModel m = load_gap_model(args, xml, base_dir);
  CUDA_OK(cudaSetDevice(device));
  gap_potential* P = new gap_potential();
  try {
    P->model = std::move(m);
    P->device = device;
    CUDA_OK(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
    upload_model(P);
  } catch (...) {
    gap_potential_finalise(P);
    throw;
  }
  return P;
}

struct CalcArgs {
  int only_descriptor = 0;  // 1-based, 0 = all
  bool use_mask = false, do_epc = false, do_var = false;
  double var_reg = 0.001;   // gap_variance_regularisation (IPModel_GAP.f95:332)
};
CalcArgs parse_calc_args(const gap_potential* P, const char* args_str) {
  CalcArgs a;
  if (!args_str || !*args_str) return a;
  ArgDict d(args_str);
  if (d.has("r_scale") || d.has("E_scale"))  // IPModel_GAP.f95:348-350
    throw GapError("IPModel_GAP_Calc: rescaling of potential at the calc() stage with r_scale and E_scale not yet implemented!");
  if (d.has("atom_mask_name") && d.str("atom_mask_name", "NONE") != "NONE") {
    if (P->mask_N < 0)  // :345-346
      throw GapError("IPModel_GAP_Calc did not find " + d.str("atom_mask_name", "") + " property in the atoms object (supply it with gap_potential_set_atom_mask)");
    if (P->n_ranks > 1)  // :375-378
      throw GapError("IPModel_GAP: atom_mask_name " + d.str("atom_mask_name", "") + " present while running the partitioned (MPI) version");
    a.use_mask = true;
  }
  if (d.logical("print_gap_variance", false))
    throw GapError("IPModel_GAP_Calc: print_gap_variance is not supported by the B200 path (use local_gap_variance)");
  a.do_var = !d.str("local_gap_variance", "").empty();
  a.var_reg = d.real("gap_variance_regularisation", 0.001);
  if (a.do_var && a.var_reg < 0.0)  // gp_predict.f95:4002-4003
    throw GapError("gpCoordinates_initialise_variance_estimate: regularisation (" + std::to_string(a.var_reg) + ") is negative.");
  a.do_epc = !d.str("energy_per_coordinate", "").empty();
  if (d.has("only_descriptor")) {
    a.only_descriptor = (int)d.integer("only_descriptor", 0);
    if (P->model.coord.size() <= 1) a.only_descriptor = 0;  // :399
  }
  return a;
}

// SOAP adjoint + scatter of a coordinate: the DMMA kernels of soap.cu, or the general path of soap_general.cu.  fpair != NULL: deterministic
// scatter (force = the per-atom scratch of the centres' own sums, see gap_device.cuh)
int soap_adjoint_any(gap_potential* P, const CoordDev& cd, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off, const int* nbr_end,
                      const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, const double* x, const double* xlm,
                      const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride, const double* epart, int n_tiles_n,
                      double* local_e, double e_scale, double* force, double* vir_part, double* local_virial, cudaStream_t st, int* launches,
                      double* fpair = nullptr) {
  if (cd.grid_passes && gvec && !fpair) {  // GTO / POLY: dE/dx -> Lambda~ on the radial grid, then one run of the default kernels per grid slice
    const size_t stride = (size_t)(n_centres_ub > 0 ? n_centres_ub : 1) * cd.h.nlm * cd.h_grid.K1;
    P->b_lam_pass.ensure(sizeof(double) * stride * cd.grid_passes);
    launch_soap_lambda_grid(cd.d_sp, cd.h, cd.gen, n_centres_dev, n_centres_ub, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride,
                            P->b_lam_pass.as<double>(), stride, cd.grid_gp, st, launches);
    for (int p = 0; p < cd.grid_passes; p++)  // the energy fold (epart) belongs to one pass only; every pass has its own block of virial partials
      launch_soap_adjoint(cd.d_sp_grid[p], cd.h_grid, centres, n_centres_dev, n_centres_ub, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec,
                          ldg, g_splits, g_split_stride, p == 0 ? epart : nullptr, n_tiles_n, local_e, e_scale, force,
                          vir_part ? vir_part + 9 * (size_t)n_centres_ub * p : nullptr, local_virial, nullptr, st, launches,
                          P->b_lam_pass.as<double>() + stride * p);
    return cd.grid_passes;
  }
  if (cd.hybrid && gvec) {  // variant-specific pull-back dE/dx -> Lambda, then the default path's transform pull-back and neighbour phase
    P->b_lambda.ensure(sizeof(double) * (size_t)(n_centres_ub > 0 ? n_centres_ub : 1) * cd.h.nlm * cd.h.K1);
    launch_soap_lambda_general(cd.d_sp, cd.h, cd.gen, n_centres_dev, n_centres_ub, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride,
                               P->b_lambda.as<double>(), st, launches);
    launch_soap_adjoint(cd.d_sp, cd.h, centres, n_centres_dev, n_centres_ub, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg,
                        g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, fpair, st, launches,
                        P->b_lambda.as<double>());
  } else if (cd.general)
    launch_soap_adjoint_general(cd.d_sp, cd.h, cd.gen, centres, n_centres_dev, n_centres_ub, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec,
                                ldg, g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, fpair, st, launches);
  else
    launch_soap_adjoint(cd.d_sp, cd.h, centres, n_centres_dev, n_centres_ub, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg,
                        g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, fpair, st, launches);
  return 1;  // blocks of n_centres_ub virial partials written
}

// reverse index of the neighbour-list slots by receiving atom, for the deterministic scatter; returns the number of slots
long det_build_index(gap_potential* P, int N, int first, int last, bool ext, cudaStream_t st) {
  const int row_cap = ext ? 0 : P->list_row_cap;  // fixed-capacity rows (speculative layout) or packed CSR
  const long n_slots = ext ? P->ext_nnz : P->list_slots;
  (void)last;
  P->b_joff.ensure(sizeof(int) * (size_t)(N + 2));
  if (P->b_fself.cap < sizeof(double) * 3 * (size_t)(N + 1)) {
    P->b_fself.ensure(sizeof(double) * 3 * (size_t)(N + 1));
    CUDA_OK(cudaMemsetAsync(P->b_fself.p, 0, P->b_fself.cap, st));
  }
  P->b_fpair.ensure(sizeof(double) * 3 * (size_t)(n_slots + 1));
  if (P->list_reused && P->det_slots == n_slots) return n_slots;  // same list as in the previous call: the index still holds
  P->det_slots = n_slots;
  if (n_slots == 0) {
    CUDA_OK(cudaMemsetAsync(P->b_joff.p, 0, sizeof(int) * (size_t)(N + 2), st));
    return 0;
  }
  if (n_slots > 2147483000L) throw GapError("deterministic scatter: neighbour list too large");
  for (DevBuf* b : {&P->b_dkeys, &P->b_dkeys2, &P->b_dvals, &P->b_dvals2}) b->ensure(sizeof(int) * (size_t)n_slots);
  const int nb = (int)((n_slots + 255) / 256);
  k_det_keys<<<nb, 256, 0, st>>>(n_slots, row_cap, first, P->cv_end, P->cv_j, N, P->b_dkeys.as<int>(), P->b_dvals.as<int>());
  int bits = 1;
  while ((1L << bits) <= N && bits < 31) bits++;
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)n_slots, 0, bits, st);
  P->b_dcub.ensure(bytes + 256);
  bytes = P->b_dcub.cap;
  cub::DeviceRadixSort::SortPairs(P->b_dcub.p, bytes, P->b_dkeys.as<int>(), P->b_dkeys2.as<int>(), P->b_dvals.as<int>(), P->b_dvals2.as<int>(), (int)n_slots, 0,
                                  bits, st);
  k_det_joff<<<nb, 256, 0, st>>>(P->b_dkeys2.as<int>(), n_slots, N, P->b_joff.as<int>());
  P->launches += 3;
  return n_slots;
}

// select + compact the centres of SOAP coordinate cd among atoms [first,last).  The number of centres stays on the
// device (P->b_scan[n], see nc_dev()); the host sizes grids and buffers with the upper bound n = last - first.
// zero_* (optional): output buffers cleared by the same launch when the one-block path is taken; *zeroed tells whether it was.
int select_centres(gap_potential* P, const CoordDev& cd, const int* d_Z, int first, int last, cudaStream_t st, double* za = nullptr,
                   size_t nza = 0, double* zb = nullptr, size_t nzb = 0, double* zc = nullptr, size_t nzc = 0, bool* zeroed = nullptr) {
  int n = last - first;
  if (zeroed) *zeroed = false;
  if (n <= 0) return 0;
  int launches = 0;
  P->b_flags.ensure(sizeof(int) * (n + 1));
  P->b_scan.ensure(sizeof(int) * (n + 1));
  P->b_centres.ensure(sizeof(int) * (n + 1));
  if (n <= SEL_MAX_N) {
    const bool fuse = za != nullptr;
    CentreZ cz;
    cz.n_Z = cd.h.n_Z;
    for (int q = 0; q < SOAP_SPECIES_CAP; q++) cz.Z[q] = cd.h.centre_Z[q];
    launch_pdl(k_select_compact_block, dim3(fuse ? 1 + SEL_ZERO_BLOCKS : 1), dim3(SEL_THREADS), 0, st, d_Z, first, last, cz, P->b_centres.as<int>(),
               P->b_scan.as<int>() + n, za, nza, zb, nzb, zc, nzc);
    if (zeroed) *zeroed = fuse;
    P->launches += 1;
    return n;
  }
  launch_select_centres(d_Z, first, last, cd.d_sp, P->b_flags.as<int>(), st, &launches);
  size_t cb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, n + 1);
  P->b_cub.ensure(cb + 256);
  size_t bytes = P->b_cub.cap;
  cub::DeviceScan::ExclusiveSum(P->b_cub.p, bytes, P->b_flags.as<int>(), P->b_scan.as<int>(), n + 1, st);
  launches++;
  launch_compact(P->b_scan.as<int>(), P->b_flags.as<int>(), first, n, P->b_centres.as<int>(), st, &launches);
  P->launches += launches;
  return n;
}
const int* nc_dev(const gap_potential* P, int n_ub) { return P->b_scan.as<int>() + n_ub; }

void soap_forward_stage(gap_potential* P, const CoordDev& cd, int n_ub, const double* d_pos, const int* d_Z, const Lattice9& lat, cudaStream_t st) {
  int launches = 0;
  int nc_pad = round_up(n_ub > 0 ? n_ub : 1, COV_BM);
  P->b_x.ensure(sizeof(double) * (size_t)nc_pad * cd.d_pad);
  P->b_xlm.ensure(sizeof(double) * (size_t)(n_ub > 0 ? n_ub : 1) * cd.h.nlm * cd.h.K1);
  P->b_pnorm.ensure(sizeof(double) * (size_t)nc_pad);
  if (cd.grid_passes) {  // GTO / POLY: the radial grid in slices through the default path's kernels, then the per-l map, mixing and element list
    const size_t stride = (size_t)(n_ub > 0 ? n_ub : 1) * cd.h.nlm * cd.h_grid.K1;
    P->b_xt_pass.ensure(sizeof(double) * stride * cd.grid_passes);
    for (int p = 0; p < cd.grid_passes; p++)
      launch_soap_forward(cd.d_sp_grid[p], cd.h_grid, P->b_centres.as<int>(), nc_dev(P, n_ub), n_ub, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, lat,
                          P->b_x.as<double>(), P->b_xt_pass.as<double>() + stride * p, P->b_pnorm.as<double>(), st, &launches, 1);
    launch_soap_power_grid(cd.d_sp, cd.h, cd.gen, P->b_centres.as<int>(), nc_dev(P, n_ub), n_ub, d_Z, P->b_xt_pass.as<double>(), stride, cd.grid_gp,
                           P->b_x.as<double>(), P->b_xlm.as<double>(), P->b_pnorm.as<double>(), st, &launches);
  } else if (cd.hybrid) {  // density expansion on the default path's kernels, channel mixing + element list from the stored X_lm
    launch_soap_forward(cd.d_sp, cd.h, P->b_centres.as<int>(), nc_dev(P, n_ub), n_ub, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, lat,
                        P->b_x.as<double>(), P->b_xlm.as<double>(), P->b_pnorm.as<double>(), st, &launches, 1);
    launch_soap_power_general(cd.d_sp, cd.h, cd.gen, nc_dev(P, n_ub), n_ub, P->b_xlm.as<double>(), P->b_x.as<double>(), P->b_pnorm.as<double>(), st,
                              &launches);
  } else if (cd.general)
    launch_soap_forward_general(cd.d_sp, cd.h, cd.gen, P->b_centres.as<int>(), nc_dev(P, n_ub), n_ub, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z,
                                lat, P->b_x.as<double>(), P->b_xlm.as<double>(), P->b_pnorm.as<double>(), st, &launches);
  else
    launch_soap_forward(cd.d_sp, cd.h, P->b_centres.as<int>(), nc_dev(P, n_ub), n_ub, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, lat,
                        P->b_x.as<double>(), P->b_xlm.as<double>(), P->b_pnorm.as<double>(), st, &launches);
  P->launches += launches;
}

// covariance for rows [0, nc_pad) of b_x: epart (all rows) and, if want_grad, gvec (all rows)
// rows_dev: device-side row count (NULL = all nc rows are real)
void covariance_stage(gap_potential* P, const CoordDev& cd, int nc, const int* rows_dev, bool want_grad, bool allow_split, cudaStream_t st) {
  int launches = 0;
  int nc_pad = round_up(nc > 0 ? nc : 1, COV_BM);
  size_t budget = (size_t)1 << 30;  // bytes of acoef kept live at once
  int chunk = (int)(budget / ((size_t)cd.M_pad * sizeof(double)) / COV_BM) * COV_BM;
  if (chunk < COV_BM) chunk = COV_BM;
  if (chunk > nc_pad) chunk = nc_pad;
  const int Mr = cd.M > 0 ? cd.M : 1;
  const int bn1 = cov_gemm1_bn(chunk, Mr, P->n_sm), n_tiles_n = (Mr + bn1 - 1) / bn1;
  P->g_tiles_n = n_tiles_n;
  P->b_acoef.ensure(sizeof(double) * (size_t)chunk * cd.M_pad);
  P->b_epart.ensure(sizeof(double) * (size_t)nc_pad * n_tiles_n);
  int ksplit = allow_split ? cov_gemm2_ksplit(chunk, cd.dn_pad, cd.bn2, P->n_sm) : 1;
  P->g_splits = ksplit;
  P->g_split_stride = (size_t)nc_pad * cd.dn_pad;
  if (want_grad) P->b_gvec.ensure(sizeof(double) * (size_t)nc_pad * cd.dn_pad * ksplit);
  for (int r0 = 0; r0 < nc_pad; r0 += chunk) {
    int rows = std::min(chunk, nc_pad - r0);
    launch_cov_gemm1(P->b_x.as<double>() + (size_t)r0 * cd.d_pad, cd.d_pad, cd.sp_rows, cd.d_pad, rows, r0, rows_dev, bn1, Mr, round_up(cd.h.d, 4),
                     cd.alpha, cd.cp, P->b_acoef.as<double>(), cd.M_pad, P->b_epart.as<double>() + (size_t)r0 * n_tiles_n, n_tiles_n, st, &launches);
    mark(P, st, ST_COV_GEMM1);
    if (want_grad) {
      launch_cov_gemm2(P->b_acoef.as<double>(), cd.M_pad, cd.st_rows, cd.M_pad, rows, r0, rows_dev, cd.dn_pad, cd.bn2, ksplit, round_up(Mr, 4),
                       P->b_gvec.as<double>() + (size_t)r0 * cd.dn_pad, cd.dn_pad, P->g_split_stride, st, &launches);
      mark(P, st, ST_COV_GEMM2);
    }
  }
  P->launches += launches;
}

// ---- predictive variance (optional output; variance.cu) -------------------------------------------------------------
// k_mm of a coordinate, factorised (soap) or inverted (distance_2b), for this regularisation
// (gpCoordinates_initialise_variance_estimate, gp_predict.f95:3970-4085)
void ensure_variance_model(gap_potential* P, size_t ic, double reg, cudaStream_t st) {
  CoordDev& cd = P->cd[ic];
  const Coordinate& c = P->model.coord[ic];
  if (cd.var_mat && cd.var_reg == reg) return;  // :3990-3996
  if (cd.var_mat) { cudaFree(cd.var_mat); cd.var_mat = nullptr; }
  const int M = c.M;
  if (c.kind == DESC_ANGLE_3B) throw GapError("GAP variance for angle_3b coordinates is not supported by the B200 path");
  if (M <= 0) return;
  int launches = 0;
  if (cd.kind == DESC_SOAP) {
    const int n_rows_pad = round_up(M, COV_BM), dnp = round_up(M, 128);
    P->b_vc.ensure(sizeof(double) * (size_t)n_rows_pad * dnp);
    launch_cov_gemm2(cd.sp_rows, cd.d_pad, cd.sp_rows, cd.d_pad, n_rows_pad, 0, nullptr, dnp, 128, 1, round_up(cd.h.d, 4), P->b_vc.as<double>(), dnp,
                     0, st, &launches);
    CUDA_OK(cudaMalloc(&cd.var_mat, sizeof(double) * (size_t)M * M));
    launch_var_kmm_finish(P->b_vc.as<double>(), dnp, M, cd.cp, c.f0 * c.f0, reg * reg, cd.var_mat, st, &launches);
    std::string err;
    const int info = var_factorise(cd.var_mat, M, st, &err);
    if (info != 0) {
      cudaFree(cd.var_mat);
      cd.var_mat = nullptr;
      if (info == -1000) throw GapError("gpCoordinates_initialise_variance_estimate: " + err);
      throw GapError("gpCoordinates_initialise_variance_estimate: k_mm is not positive definite (dpotrf info = " + std::to_string(info) + ")");
    }
  } else {
    if (M > 64) throw GapError("GAP variance for distance_2b coordinates with more than 64 sparse points is not supported by the B200 path");
    if (c.d2b.exponents.size() != 1 || c.d2b.exponents[0] != 1.0 || c.d2b.tail_exponent != 0 || c.d2b.intra_mode != 0)
      throw GapError("GAP variance for distance_2b coordinates with exponents / tail / residue options is not supported by the B200 path");
    // ARD_SE, one permutation (:4034-4039), normalisation (:4055-4068), delta^2, f0^2, regularisation (:4070-4075); host, M <= 64
    std::vector<double> K((size_t)M * M), Kinv((size_t)M * M, 0.0);
    const double th = c.theta[0];
    for (int i = 0; i < M; i++)
      for (int j = 0; j < M; j++) {
        double t = (c.sparseX[i] - c.sparseX[j]) / th, v = std::exp(-0.5 * t * t);
        v = (i == j) ? c.sparseCutoff[i] * c.sparseCutoff[i] : v * c.sparseCutoff[i] * c.sparseCutoff[j];  // k_mm(i,i) = 1 before normalisation
        K[(size_t)i * M + j] = v * c.delta * c.delta + c.f0 * c.f0 + (i == j ? reg * reg : 0.0);
      }
    for (int j = 0; j < M; j++) {  // lower Cholesky in place (row-major; the matrix is symmetric)
      double sdiag = K[(size_t)j * M + j];
      for (int k = 0; k < j; k++) sdiag -= K[(size_t)j * M + k] * K[(size_t)j * M + k];
      if (!(sdiag > 0.0)) throw GapError("gpCoordinates_initialise_variance_estimate: k_mm is not positive definite");
      const double ljj = std::sqrt(sdiag);
      K[(size_t)j * M + j] = ljj;
      for (int i = j + 1; i < M; i++) {
        double t = K[(size_t)i * M + j];
        for (int k = 0; k < j; k++) t -= K[(size_t)i * M + k] * K[(size_t)j * M + k];
        K[(size_t)i * M + j] = t / ljj;
      }
    }
    std::vector<double> y(M);
    for (int col = 0; col < M; col++) {  // k_mm^-1 e_col
      for (int i = 0; i < M; i++) {
        double t = (i == col) ? 1.0 : 0.0;
        for (int k = 0; k < i; k++) t -= K[(size_t)i * M + k] * y[k];
        y[i] = t / K[(size_t)i * M + i];
      }
      for (int i = M - 1; i >= 0; i--) {
        double t = y[i];
        for (int k = i + 1; k < M; k++) t -= K[(size_t)k * M + i] * y[k];
        y[i] = t / K[(size_t)i * M + i];
      }
      for (int i = 0; i < M; i++) Kinv[(size_t)i * M + col] = y[i];
    }
    CUDA_OK(cudaMalloc(&cd.var_mat, sizeof(double) * (size_t)M * M));
    CUDA_OK(cudaMemcpyAsync(cd.var_mat, Kinv.data(), sizeof(double) * (size_t)M * M, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  cd.var_reg = reg;
  P->launches += launches;
}

// variance of every centre of SOAP coordinate ic (b_x rows [0, nc)) and, if want_grad, its gradient scattered to atoms
void variance_soap(gap_potential* P, size_t ic, int nc, const int* ncd, bool want_grad, const double* d_pos, const int* d_Z, const Lattice9& lat,
                   double reg, cudaStream_t st) {
  const CoordDev& cd = P->cd[ic];
  const int M = cd.M;
  if (M <= 0 || nc <= 0) return;
  int launches = 0;
  const int ld = cd.M_pad, nc_pad = round_up(nc, COV_BM);
  const size_t budget = (size_t)512 << 20;
  int chunk = (int)(budget / ((size_t)ld * sizeof(double)) / COV_BM) * COV_BM;
  if (chunk < COV_BM) chunk = COV_BM;
  if (chunk > nc_pad) chunk = nc_pad;
  P->b_vk.ensure(sizeof(double) * (size_t)chunk * ld);
  P->b_vq.ensure(sizeof(double) * (size_t)chunk * ld);
  const int ksplit = want_grad ? cov_gemm2_ksplit(chunk, cd.dn_pad, cd.bn2, P->n_sm) : 1;
  const size_t split_stride = (size_t)nc_pad * cd.dn_pad;
  if (want_grad) P->b_gvec.ensure(sizeof(double) * split_stride * ksplit);
  const double diag = cd.delta * cd.delta + cd.f0 * cd.f0 + reg * reg;  // gp_predict.f95:3874
  double* Cm = P->b_vk.as<double>();
  double* Q = P->b_vq.as<double>();
  for (int r0 = 0; r0 < nc_pad; r0 += chunk) {
    const int rows = std::min(chunk, nc_pad - r0);
    launch_cov_gemm2(P->b_x.as<double>() + (size_t)r0 * cd.d_pad, cd.d_pad, cd.sp_rows, cd.d_pad, rows, r0, ncd, round_up(M, 128), 128, 1,
                     round_up(cd.h.d, 4), Cm, ld, 0, st, &launches);
    launch_var_prepare(Cm, ld, rows, M, cd.scut, cd.cp, Q, st, &launches);
    std::string err;
    if (var_solve(cd.var_mat, M, Q, ld, rows, st, &err)) throw GapError("gpCoordinates_Predict (variance): " + err);
    launches += 2;
    launch_var_finish(Cm, Q, ld, rows, r0, ncd, M, cd.scut, cd.cp, diag, P->b_centres.as<int>(), P->b_lgv.as<double>(), want_grad ? 1 : 0,
                      P->b_varflag.as<int>(), st, &launches);
    if (want_grad)
      launch_cov_gemm2(Cm, ld, cd.st_rows, cd.M_pad, rows, r0, ncd, cd.dn_pad, cd.bn2, ksplit, round_up(M, 4),
                       P->b_gvec.as<double>() + (size_t)r0 * cd.dn_pad, cd.dn_pad, split_stride, st, &launches);
  }
  if (want_grad)  // pull-back through the descriptor: gap_variance_gradient(:,j) += grad_variance . grad_data(:,:,n)  (IPModel_GAP.f95:485-486)
    soap_adjoint_any(P, cd, P->b_centres.as<int>(), ncd, nc, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, lat, P->b_x.as<double>(),
                     P->b_xlm.as<double>(), P->b_pnorm.as<double>(), P->b_gvec.as<double>(), cd.dn_pad, ksplit, split_stride, nullptr, 0, nullptr,
                     -1.0 /* the kernel scatters force = -e_scale f_gp */, P->b_gvg.as<double>(), nullptr, nullptr, st, &launches);
  P->launches += launches;
}

// An externally supplied neighbour list (quip_lammps_wrapper): CSR over all N = nlocal + nghost atoms with zero shifts
// (periodic images are explicit ghost atoms), and d_Zc = Z for the atoms that are centres, -1 for the others
// (the reference's atom_mask_name=local, quip_lammps_wrapper.f95:97-147).
struct ExtList {
  const int *off, *j, *s, *Zc;
  int nlocal;
  long nnz;
};

void calc_device_impl(gap_potential* P, int N, const double* d_pos, const int* d_Z, const double* lattice, const int* pbc,
                      const char* args_str, bool want_grad, double* d_packed, double* d_le_user, double* d_lv, cudaStream_t st,
                      const ExtList* ext = nullptr) {
  CUDA_OK(cudaSetDevice(P->device));
  CalcArgs ca = parse_calc_args(P, args_str);
  if (!lattice || !pbc) throw GapError("gap_potential_calc: lattice and pbc are required");
  int first = (int)((long long)P->rank * N / P->n_ranks), last = (int)((long long)(P->rank + 1) * N / P->n_ranks);
  const int* d_Zc = d_Z;  // atomic numbers as seen by the centre selection and the e0 sum
  P->ev_used = 0;
  P->last_stream = st;
  mark(P, st, -1);
  if (ca.use_mask) {
    if (ext) throw GapError("IPModel_GAP_Calc: atom_mask_name cannot be combined with an external neighbour list");
    if (P->mask_N != N) throw GapError("IPModel_GAP_Calc: the atom mask has " + std::to_string(P->mask_N) + " entries, the configuration " + std::to_string(N) + " atoms");
    P->b_zc.ensure(sizeof(int) * (size_t)(N + 1));
    if (N > 0) k_apply_mask<<<(N + 255) / 256, 256, 0, st>>>(d_Z, P->b_mask.as<int>(), N, P->b_zc.as<int>());
    P->launches += 1;
    d_Zc = P->b_zc.as<int>();
  }
  const size_t n_coord = P->cd.size();
  P->epc_valid = false;
  if (ca.do_epc) P->b_epc.ensure(sizeof(double) * (n_coord + 1));
  P->var_N = -1;
  if (ca.do_var) {
    P->b_lgv.ensure(sizeof(double) * (size_t)(N + 1));
    P->b_varflag.ensure(sizeof(int));
    CUDA_OK(cudaMemsetAsync(P->b_lgv.p, 0, sizeof(double) * (size_t)(N + 1), st));
    CUDA_OK(cudaMemsetAsync(P->b_varflag.p, 0, sizeof(int), st));
    if (want_grad) {
      P->b_gvg.ensure(sizeof(double) * 3 * (size_t)(N + 1));
      CUDA_OK(cudaMemsetAsync(P->b_gvg.p, 0, sizeof(double) * 3 * (size_t)(N + 1), st));
    }
    for (size_t ic = 0; ic < n_coord; ic++) ensure_variance_model(P, ic, ca.var_reg, st);  // IPModel_GAP.f95:412-414
  }
  double* d_le = d_le_user;
  if (!d_le) {
    P->b_le.ensure(sizeof(double) * (size_t)(N + 1));
    d_le = P->b_le.as<double>();
  }
  const size_t nz_packed = 10 + 3 * (size_t)N, nz_le = (size_t)N + (d_le_user ? 0 : 1), nz_lv = d_lv ? 9 * (size_t)N : 0;
  if (ext) {
    first = 0;
    last = ext->nlocal;
    d_Zc = ext->Zc;
    P->cv_off = ext->off; P->cv_end = ext->off + 1; P->cv_j = ext->j; P->cv_s = ext->s;
    P->pending_check = false;
    P->list_valid = false;
    P->list_reused = false;
    P->ext_nnz = ext->nnz;
  }
  // The centre selection of the first SOAP coordinate needs Z only: it is launched ahead of the neighbour-list build and, for small
  // partitions, the same launch zeroes the outputs (one launch instead of two at the head of the step).
  int hoisted_ic = -1, hoisted_nc = 0;
  bool zeroed = false;
  for (size_t ic = 0; ic < n_coord && hoisted_ic < 0; ic++)
    if (P->cd[ic].kind == DESC_SOAP && !(ca.only_descriptor && (int)ic + 1 != ca.only_descriptor)) hoisted_ic = (int)ic;
  if (hoisted_ic >= 0) hoisted_nc = select_centres(P, P->cd[hoisted_ic], d_Zc, first, last, st, d_packed, nz_packed, d_le, nz_le, d_lv, nz_lv, &zeroed);
  if (!zeroed) {
    const size_t nz = (d_lv ? 9 : 3) * (size_t)N + 10;
    int zb = (int)std::min<size_t>((nz + 255) / 256, 2048);
    k_zero_outputs<<<zb, 256, 0, st>>>(d_packed, nz_packed, d_le, nz_le, d_lv, nz_lv);
    P->launches += 1;
  }
  mark(P, st, ST_OTHER);
  if (!ext) build_connect(P, N, first, last, d_pos, lattice, pbc, P->model.cutoff, false, true, st);
  mark(P, st, ST_CONNECT);
  Lattice9 lat;
  for (int k = 0; k < 9; k++) lat.v[k] = lattice[k];

  double* d_force = d_packed + 10;
  const double es = P->model.E_scale;
  // deterministic force scatter: index the list slots by receiving atom once per list
  const bool det = P->deterministic && want_grad;
  long det_slots = 0;
  if (det) {
    bool any_soap = false;  // (or angle_3b: the coordinates whose pair forces go through the slot index)
    for (const CoordDev& cdv : P->cd) any_soap = any_soap || cdv.kind == DESC_SOAP || cdv.kind == DESC_ANGLE_3B;
    if (any_soap) det_slots = det_build_index(P, N, first, last, ext != nullptr, st);
  }

  // virial partial slots
  size_t slots_cap = 0;
  for (const CoordDev& cd : P->cd)
    slots_cap += cd.kind == DESC_SOAP ? (size_t)(last - first) * (cd.grid_passes > 1 ? cd.grid_passes : 1) : (size_t)((last - first + 3) / 4);
  if (want_grad) P->b_vir.ensure(sizeof(double) * 9 * (slots_cap + 1));
  size_t slot = 0;

  for (size_t ic = 0; ic < P->cd.size(); ic++) {
    const CoordDev& cd = P->cd[ic];
    int launches = 0;
    if (ca.only_descriptor && (int)ic + 1 != ca.only_descriptor) {
      // skipped coordinate (:399-401): its energy_per_coordinate entry is zero
    } else if (cd.kind == DESC_SOAP) {
      // upper bound; the count itself stays on the device
      int nc = (int)ic == hoisted_ic ? hoisted_nc : select_centres(P, cd, d_Zc, first, last, st);
      mark(P, st, ST_OTHER);
      if (nc > 0) {
        const int* ncd = nc_dev(P, nc);
        const bool glob = cd.general && cd.gen.global_mode;
        if (glob && (first != 0 || last != N))  // (the reference sums the centres' expansions of the whole configuration, :8357-8367)
          throw GapError("soap average=T (one descriptor per configuration) cannot be evaluated on a partition of the centres");
        soap_forward_stage(P, cd, nc, d_pos, d_Z, lat, st);
        mark(P, st, ST_SOAP_FWD);
        // average=T: ONE descriptor (row 0 of x) whatever the number of centres
        covariance_stage(P, cd, glob ? 1 : nc, glob ? nullptr : ncd, want_grad, true, st);
        if (want_grad) {
          if (det && det_slots > 0) CUDA_OK(cudaMemsetAsync(P->b_fpair.p, 0, sizeof(double) * 3 * (size_t)det_slots, st));  // (unvisited slots)
          const int vir_blocks = soap_adjoint_any(P, cd, P->b_centres.as<int>(), ncd, nc, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, lat,
                           P->b_x.as<double>(), P->b_xlm.as<double>(), P->b_pnorm.as<double>(), P->b_gvec.as<double>(), cd.dn_pad, P->g_splits,
                           P->g_split_stride, P->b_epart.as<double>(), P->g_tiles_n, d_le, es, det ? P->b_fself.as<double>() : d_force,
                           P->b_vir.as<double>() + 9 * slot, d_lv, st, &launches, det ? P->b_fpair.as<double>() : nullptr);
          if (det) {  // F_j += (pair forces on j, in slot order) + (j's own sum as a centre)
            k_det_gather<<<(N + 3) / 4, 128, 0, st>>>(N, P->b_joff.as<int>(), P->b_dvals2.as<int>(), P->b_fpair.as<double>(), P->b_fself.as<double>(), d_force);
            launches += 1;
          }
          slot += (size_t)nc * vir_blocks;
          mark(P, st, ST_SOAP_ADJ);
        } else if (glob) {  // energy only: e_i shared by all centres (IPModel_GAP.f95:454-459)
          soap_adjoint_any(P, cd, P->b_centres.as<int>(), ncd, nc, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, lat, P->b_x.as<double>(),
                           P->b_xlm.as<double>(), P->b_pnorm.as<double>(), nullptr, 0, 0, 0, P->b_epart.as<double>(), P->g_tiles_n, d_le, es, nullptr,
                           nullptr, nullptr, st, &launches);
          mark(P, st, ST_OTHER);
        } else {
          launch_energy_rows(P->b_epart.as<double>(), P->g_tiles_n, P->b_centres.as<int>(), ncd, nc, es, d_le, st, &launches);
          mark(P, st, ST_OTHER);
        }
        if (ca.do_var && glob) throw GapError("local_gap_variance is not supported for soap average=T coordinates on the B200 path");
        if (ca.do_var) {
          variance_soap(P, ic, nc, ncd, want_grad, d_pos, d_Z, lat, ca.var_reg, st);
          mark(P, st, ST_OTHER);
        }
      }
    } else if (cd.kind == DESC_ANGLE_3B) {
      int nb = 0;
      if (ca.do_var) throw GapError("GAP variance for angle_3b coordinates is not supported by the B200 path");
      const long n_slots = ext ? P->ext_nnz : P->list_slots;
      P->b_a3idx.ensure(sizeof(int) * (size_t)(n_slots + 1));
      const bool det3 = det && det_slots > 0;
      if (det3) CUDA_OK(cudaMemsetAsync(P->b_fpair.p, 0, sizeof(double) * 3 * (size_t)det_slots, st));  // (slots outside this descriptor's cutoff)
      launch_angle3b(cd.p3, first, last, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, d_Zc, lat, es, want_grad ? 1 : 0, d_le,
                     want_grad ? (det3 ? P->b_fself.as<double>() : d_force) : nullptr, det3 ? P->b_fpair.as<double>() : nullptr,
                     want_grad ? P->b_vir.as<double>() + 9 * slot : nullptr, want_grad ? d_lv : nullptr, P->b_a3idx.as<int>(), st, &launches, &nb);
      if (det3) {
        k_det_gather<<<(N + 3) / 4, 128, 0, st>>>(N, P->b_joff.as<int>(), P->b_dvals2.as<int>(), P->b_fpair.as<double>(), P->b_fself.as<double>(), d_force);
        launches += 1;
      }
      if (want_grad) slot += nb;
      mark(P, st, ST_PAIR2B);
    } else {
      int nb = 0;
      Pair2bDev p2 = cd.p2;
      if (p2.intra_mode) {  // the residue ids are an Atoms property in the reference (resid_name): supplied with gap_potential_set_resid
        if (P->resid_N != N) throw GapError("distance_2b_calc did not find " + P->model.coord[ic].d2b.resid_name + " property (residue id) in the atoms object (supply it with gap_potential_set_resid)");
        p2.resid = P->b_resid.as<int>();
      }
      launch_pair2b(p2, first, last, P->cv_off, P->cv_end, P->cv_j, P->cv_s, d_pos, d_Z, d_Zc, ca.use_mask ? 1 : 0, lat, es, want_grad ? 1 : 0,
                    d_le, want_grad ? d_force : nullptr, want_grad ? P->b_vir.as<double>() + 9 * slot : nullptr, want_grad ? d_lv : nullptr, st,
                    &launches, &nb);
      if (want_grad) slot += nb;
      if (ca.do_var && cd.var_mat)
        launch_pair2b_var(cd.p2, cd.var_mat, cd.delta * cd.delta + cd.f0 * cd.f0 + ca.var_reg * ca.var_reg, first, last, d_Zc, P->cv_off, P->cv_end,
                          P->cv_j, P->cv_s, d_pos, d_Z, lat, P->b_lgv.as<double>(), want_grad ? P->b_gvg.as<double>() : nullptr,
                          P->b_varflag.as<int>(), st, &launches);
      mark(P, st, ST_PAIR2B);
    }
    if (ca.do_epc) {  // running total of local_e (e0 not yet added): the increments are energy_per_coordinate * E_scale (:462)
      k_sum_range<<<1, 1024, 0, st>>>(d_le, N, P->b_epc.as<double>() + ic);
      launches += 1;
    }
    P->launches += launches;
  }
  if (ca.do_epc) P->epc_valid = true;
  if (ca.do_var) { P->var_N = N; P->var_grad = want_grad; }
  // totals
  P->b_fin.ensure(sizeof(double) * 10 * FIN_BLOCKS);
  const size_t fin_work = std::max<size_t>((size_t)N, want_grad ? slot : 0);
  const int fin_blocks = (int)std::min<size_t>(FIN_BLOCKS, std::max<size_t>(1, (fin_work + FIN_THREADS - 1) / FIN_THREADS));
  launch_pdl(k_finalize, dim3(fin_blocks), dim3(FIN_THREADS), 0, st, d_Zc, N, first, last, (const double*)P->d_e0, es, d_le,
             (const double*)(want_grad ? P->b_vir.as<double>() : nullptr), want_grad ? (int)slot : 0, P->b_fin.as<double>(), P->d_fin_counter, d_packed,
             P->pending_check ? P->b_off.as<int>() + N : (int*)nullptr, P->pending_cap, P->d_hpin);
  if (P->pending_check) P->stat_clean = true;
  P->launches += 1;
  mark(P, st, ST_OTHER);
  CUDA_OK(cudaGetLastError());
}

// This rank's block of centres and, when a communicator is set, the sum of the partials over the ranks (IPModel_GAP.f95:538-556):
// d_packed (and d_le_user / d_lv if given) hold the TOTALS on every rank once the stream has drained.
void calc_reduced(gap_potential* P, int N, const double* d_pos, const int* d_Z, const double* lattice, const int* pbc, const char* args_str,
                  bool want_grad, double* d_packed, double* d_le_user, double* d_lv, cudaStream_t st) {
  if (!P->comm || comm_size(P->comm) < 2) {
    calc_device_impl(P, N, d_pos, d_Z, lattice, pbc, args_str, want_grad, d_packed, d_le_user, d_lv, st);
    return;
  }
  const size_t count = 10 + 3 * (size_t)N;
  double* partial = comm_partial_buffer(P->comm, count, d_packed, st);
  calc_device_impl(P, N, d_pos, d_Z, lattice, pbc, args_str, want_grad, partial, d_le_user, d_lv, st);
  comm_allreduce_packed(P->comm, count, d_packed, st);
  if (d_le_user) comm_allreduce_inplace(P->comm, d_le_user, (size_t)N, st);
  if (d_lv) comm_allreduce_inplace(P->comm, d_lv, 9 * (size_t)N, st);
  mark(P, st, ST_OTHER);
}

bool host_pointer_is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int guard(const std::function<void()>& fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  } catch (...) {
    g_last_error = "unknown error";
    return 2;
  }
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* gap_last_error(void) { return g_last_error.c_str(); }

int gap_potential_initialise(gap_potential** pot, const char* args_str, const char* param_str, const char* base_dir, int device) {
  return guard([&] {
    if (!pot) throw GapError("gap_potential_initialise: pot is NULL");
    *pot = nullptr;
    if (!param_str) throw GapError("gap_potential_initialise: param_str is NULL");
    *pot = create_potential(args_str ? args_str : "", param_str, base_dir && *base_dir ? base_dir : ".", device);
  });
}

int gap_potential_filename_initialise(gap_potential** pot, const char* args_str, const char* param_filename, int device) {
  return guard([&] {
    if (!pot) throw GapError("gap_potential_filename_initialise: pot is NULL");
    *pot = nullptr;
    if (!param_filename) throw GapError("gap_potential_filename_initialise: param_filename is NULL");
    std::ifstream f(param_filename, std::ios::binary);
    if (!f) throw GapError(std::string("Potential_Filename_Initialise: cannot open ") + param_filename);
    std::ostringstream ss;
    ss << f.rdbuf();
    std::string path(param_filename);
    size_t sl = path.find_last_of('/');
    std::string dir = sl == std::string::npos ? "." : (sl == 0 ? "/" : path.substr(0, sl));
    *pot = create_potential(args_str ? args_str : "", ss.str(), dir, device);
  });
}

void gap_potential_finalise(gap_potential* P) {
  if (!P) return;
  cudaSetDevice(P->device);
  if (P->stream) cudaStreamSynchronize(P->stream);
  if (P->comm) comm_destroy(P->comm);
  P->comm = nullptr;
  for (CoordDev& cd : P->cd) {
    cudaFree(cd.gen_blob);
    for (SoapDev* q : cd.d_sp_grid) cudaFree(q);
    cudaFree(cd.gen_global);
    cudaFree(cd.d_sp); cudaFree(cd.sp_rows); cudaFree(cd.st_rows); cudaFree(cd.alpha); cudaFree(cd.scut);
    cudaFree(cd.x2); cudaFree(cd.a2); cudaFree(cd.c2); cudaFree(cd.t3); cudaFree(cd.var_mat);
  }
  cudaFree(P->d_e0);
  cudaFree(P->d_fin_counter);
  if (P->h_pin) cudaFreeHost(P->h_pin);
  if (P->h_stage) cudaFreeHost(P->h_stage);
  if (P->h_disp) cudaFreeHost(P->h_disp);
  DevBuf* bufs[] = {&P->b_cell_of, &P->b_mshift, &P->b_keys, &P->b_idx, &P->b_slot, &P->b_iota, &P->b_cstart, &P->b_ccount, &P->b_end, &P->b_mask, &P->b_epc, &P->b_lgv, &P->b_gvg, &P->b_varflag, &P->b_vc, &P->b_vq, &P->b_vk, &P->b_spos, &P->b_smshift,
                    &P->b_nn, &P->b_cub, &P->b_minmax, &P->b_off, &P->b_j, &P->b_s, &P->b_d, &P->b_pos, &P->b_Z, &P->b_packed, &P->b_le,
                    &P->b_lv, &P->b_flags, &P->b_scan, &P->b_centres, &P->b_x, &P->b_xlm, &P->b_pnorm, &P->b_acoef, &P->b_gvec, &P->b_epart,
                    &P->b_vir, &P->b_fin, &P->b_xoff, &P->b_xj, &P->b_xs, &P->b_zc, &P->b_velo, &P->b_velo2, &P->b_acc, &P->b_mass, &P->b_ke, &P->b_lastpos, &P->b_disp, &P->b_resid,
                    &P->b_fpair, &P->b_fself, &P->b_dkeys, &P->b_dkeys2, &P->b_dvals, &P->b_dvals2, &P->b_joff, &P->b_dcub, &P->b_a3idx, &P->b_lambda, &P->b_xt_pass, &P->b_lam_pass};
  for (DevBuf* b : bufs) b->release();
  for (cudaEvent_t e : P->ev) cudaEventDestroy(e);
  if (P->stream) cudaStreamDestroy(P->stream);
  delete P;
}

double gap_potential_cutoff(const gap_potential* P) { return P ? P->model.cutoff : 0.0; }
int gap_potential_n_coordinate(const gap_potential* P) { return P ? (int)P->model.coord.size() : 0; }
long gap_potential_launch_count(const gap_potential* P) { return P ? P->launches : 0; }

int gap_potential_print(const gap_potential* P, char* buf, size_t n) {
  return guard([&] {
    if (!P || !buf || n == 0) throw GapError("gap_potential_print: bad arguments");
    std::ostringstream os;
    os << "IPModel_GAP : Gaussian Approximation Potential (B200 path)\n";
    os << "IPModel_GAP : label = " << P->model.label << "\n";
    os << "IPModel_GAP : cutoff = " << P->model.cutoff << "\n";
    os << "IPModel_GAP : E_scale = " << P->model.E_scale << "\n";
    os << "IPModel_GAP : gap_version = " << P->model.xml_version << "\n";
    for (size_t i = 0; i < P->model.coord.size(); i++) {
      const Coordinate& c = P->model.coord[i];
      os << "IPModel_GAP : coordinate " << (i + 1) << " : " << c.descriptor_str << " | dimensions=" << c.d << " n_sparseX=" << c.M
         << " covariance_type=" << c.covariance_type << " delta=" << c.delta << " zeta=" << c.zeta << "\n";
    }
    std::string s = os.str();
    size_t k = s.size() < n - 1 ? s.size() : n - 1;
    memcpy(buf, s.data(), k);
    buf[k] = 0;
  });
}

int gap_potential_set_partition(gap_potential* P, int rank, int n_ranks) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_partition: pot is NULL");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw GapError("gap_potential_set_partition: need 0 <= rank < n_ranks");
    P->rank = rank;
    P->n_ranks = n_ranks;
  });
}

int gap_comm_get_unique_id(char* id) {
  return guard([&] {
    if (!id) throw GapError("gap_comm_get_unique_id: id is NULL");
    comm_get_unique_id(id);
  });
}

int gap_potential_set_comm(gap_potential* P, const char* id, int rank, int n_ranks) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_comm: pot is NULL");
    if (P->comm) { comm_destroy(P->comm); P->comm = nullptr; }
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw GapError("gap_potential_set_comm: need 0 <= rank < n_ranks");
    if (n_ranks > 1) {
      if (!id) throw GapError("gap_potential_set_comm: id is NULL");
      P->comm = comm_create(id, rank, n_ranks, P->device);
    }
    P->rank = rank;
    P->n_ranks = n_ranks;
  });
}

int gap_potential_comm_timing(const gap_potential* P, double* wait_us, double* sum_us) {
  if (!P || !wait_us || !sum_us) return 1;
  cudaSetDevice(P->device);
  comm_last_stamps(P->comm, wait_us, sum_us);
  return 0;
}

int gap_potential_comm_info(const gap_potential* P, int* rank, int* n_ranks, char* transport, size_t n) {
  if (!P) return 1;
  if (rank) *rank = P->rank;
  if (n_ranks) *n_ranks = P->n_ranks;
  if (transport && n) {
    const char* t = P->comm ? comm_last_transport(P->comm) : "none";
    strncpy(transport, t, n - 1);
    transport[n - 1] = 0;
  }
  return 0;
}

int gap_potential_set_deterministic(gap_potential* P, int on) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_deterministic: pot is NULL");
    P->deterministic = on != 0;
  });
}

int gap_potential_set_cutoff_skin(gap_potential* P, double cutoff_skin) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_cutoff_skin: pot is NULL");
    if (!(cutoff_skin >= 0.0)) throw GapError("gap_potential_set_cutoff_skin: cutoff_skin must be >= 0");
    P->cutoff_skin = cutoff_skin;
    P->list_valid = false;
  });
}

int gap_potential_connect_stats(const gap_potential* P, long* n_rebuilds, long* n_reuses) {
  if (!P) return 1;
  if (n_rebuilds) *n_rebuilds = P->n_rebuilds;
  if (n_reuses) *n_reuses = P->n_reuses;
  return 0;
}

int gap_potential_set_atom_mask(gap_potential* P, int N, const int* mask) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_atom_mask: pot is NULL");
    if (!mask) { P->mask_N = -1; return; }
    if (N < 0) throw GapError("gap_potential_set_atom_mask: N < 0");
    CUDA_OK(cudaSetDevice(P->device));
    P->b_mask.ensure(sizeof(int) * (size_t)(N + 1));
    if (N > 0) CUDA_OK(cudaMemcpy(P->b_mask.p, mask, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    P->mask_N = N;
  });
}

int gap_potential_set_resid(gap_potential* P, int N, const int* resid) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_resid: pot is NULL");
    if (!resid) { P->resid_N = -1; return; }
    if (N < 0) throw GapError("gap_potential_set_resid: N < 0");
    CUDA_OK(cudaSetDevice(P->device));
    P->b_resid.ensure(sizeof(int) * (size_t)(N + 1));
    if (N > 0) CUDA_OK(cudaMemcpy(P->b_resid.p, resid, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    P->resid_N = N;
  });
}

int gap_potential_get_energy_per_coordinate(gap_potential* P, double* out) {
  return guard([&] {
    if (!P || !out) throw GapError("gap_potential_get_energy_per_coordinate: bad arguments");
    if (!P->epc_valid) throw GapError("gap_potential_get_energy_per_coordinate: the last calc was not asked for energy_per_coordinate=NAME");
    CUDA_OK(cudaSetDevice(P->device));
    const size_t n = P->cd.size();
    std::vector<double> run(n);
    CUDA_OK(cudaStreamSynchronize(P->last_stream));
    if (n) CUDA_OK(cudaMemcpy(run.data(), P->b_epc.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    const double es = P->model.E_scale;
    for (size_t k = 0; k < n; k++) out[k] = es != 0.0 ? (run[k] - (k ? run[k - 1] : 0.0)) / es : 0.0;
  });
}

int gap_potential_get_local_gap_variance(gap_potential* P, int N, double* local_gap_variance, double* gap_variance_gradient) {
  return guard([&] {
    if (!P || !local_gap_variance) throw GapError("gap_potential_get_local_gap_variance: bad arguments");
    if (P->var_N < 0 || P->var_N != N) throw GapError("gap_potential_get_local_gap_variance: the last calc was not asked for local_gap_variance=NAME (or N differs)");
    if (gap_variance_gradient && !P->var_grad)
      throw GapError("gap_potential_get_local_gap_variance: gap_variance_gradient needs a calc with forces or virials (IPModel_GAP.f95:560-564)");
    CUDA_OK(cudaSetDevice(P->device));
    CUDA_OK(cudaStreamSynchronize(P->last_stream));
    int flag = 0;
    CUDA_OK(cudaMemcpy(&flag, P->b_varflag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) throw GapError("gpCoordinates_Predict: variance_estimate: negative variance predicted");  // gp_predict.f95:3877
    if (N > 0) CUDA_OK(cudaMemcpy(local_gap_variance, P->b_lgv.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
    if (gap_variance_gradient && N > 0) CUDA_OK(cudaMemcpy(gap_variance_gradient, P->b_gvg.p, sizeof(double) * 3 * (size_t)N, cudaMemcpyDeviceToHost));
  });
}

int gap_potential_calc_device(gap_potential* P, int N, const double* d_pos, const int* d_Z, const double* lattice, const int* pbc,
                              const char* args_str, int want_grad, double* d_packed, double* d_local_e, double* d_local_virial, void* stream) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_calc_device: pot is NULL");
    if (N < 0) throw GapError("gap_potential_calc_device: N < 0");
    if (!d_packed) throw GapError("gap_potential_calc_device: d_packed is NULL");
    cudaStream_t st = stream ? (cudaStream_t)stream : P->stream;
    const bool sharded = P->comm && comm_size(P->comm) > 1;
    for (int attempt = 0; attempt < 2; attempt++) {
      calc_reduced(P, N, d_pos, d_Z, lattice, pbc, args_str, want_grad != 0 || d_local_virial != nullptr, d_packed, d_local_e, d_local_virial, st);
      if (!sharded && !P->pending_check) break;  // exact list on one rank: nothing to verify, the work is simply enqueued
      // a rank whose neighbour rows overflowed NaN-poisons its energy (k_finalize): after the reduction every rank sees it and all
      // ranks repeat together
      double e = 0.0;
      if (sharded) CUDA_OK(cudaMemcpyAsync(&e, d_packed, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      comm_check(P->comm);
      const bool ok_local = verify_connect(P);
      if (ok_local && e == e) break;
      if (ok_local) P->row_hint = -1;  // another rank overflowed: take the exact path together with it
    }
  });
}

int gap_potential_calc_device_enqueue(gap_potential* P, int N, const double* d_pos, const int* d_Z, const double* lattice, const int* pbc,
                                      const char* args_str, int want_grad, double* d_packed, double* d_local_e, double* d_local_virial, void* stream) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_calc_device_enqueue: pot is NULL");
    if (N < 0) throw GapError("gap_potential_calc_device_enqueue: N < 0");
    if (!d_packed) throw GapError("gap_potential_calc_device_enqueue: d_packed is NULL");
    cudaStream_t st = stream ? (cudaStream_t)stream : P->stream;
    calc_reduced(P, N, d_pos, d_Z, lattice, pbc, args_str, want_grad != 0 || d_local_virial != nullptr, d_packed, d_local_e, d_local_virial, st);
  });
}

int gap_potential_calc_device_verify(gap_potential* P, int* repeat) {
  return guard([&] {
    if (!P || !repeat) throw GapError("gap_potential_calc_device_verify: bad arguments");
    comm_check(P->comm);
    *repeat = verify_connect(P) ? 0 : 1;
  });
}

int gap_potential_calc(gap_potential* P, int N, const double* pos, const int* Z, const double* lattice, const int* pbc, const char* args_str,
                       double* energy, double* local_e, double* force, double* virial, double* local_virial) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_calc: pot is NULL");
    if (N < 0) throw GapError("gap_potential_calc: N < 0");
    if (N > 0 && (!pos || !Z)) throw GapError("gap_potential_calc: pos/Z are NULL");
    CUDA_OK(cudaSetDevice(P->device));
    cudaStream_t st = P->stream;
    const size_t n3 = 3 * (size_t)N;
    // inputs: pos and Z go into one device buffer [pos (3N f64) | Z (N i32)].  Host arrays that are already page-locked (cudaHostAlloc /
    // cudaHostRegister by the caller) are copied from where they lie; pageable ones are staged through ONE pinned buffer and one H2D
    // copy.  Results: [E | virial | F] (+ local_e, local_virial) come back the same way; one synchronisation per call.
    const size_t in_bytes = sizeof(double) * n3 + sizeof(int) * (size_t)N;
    const size_t out_doubles = 16 + n3 + (local_e ? (size_t)N : 0) + (local_virial ? 9 * (size_t)N : 0);
    P->b_pos.ensure(in_bytes + 64);
    P->b_packed.ensure(sizeof(double) * (10 + n3));
    P->b_le.ensure(sizeof(double) * (size_t)(N + 1));
    if (local_virial) P->b_lv.ensure(sizeof(double) * 9 * (size_t)(N + 1));
    if (P->h_stage_cap < std::max(in_bytes, sizeof(double) * out_doubles) + 64) {
      if (P->h_stage) cudaFreeHost(P->h_stage);
      P->h_stage = nullptr;
      P->h_stage_cap = 2 * std::max(in_bytes, sizeof(double) * out_doubles) + 4096;
      CUDA_OK(cudaHostAlloc((void**)&P->h_stage, P->h_stage_cap, cudaHostAllocDefault));
    }
    double* d_pos = P->b_pos.as<double>();
    int* d_Z = (int*)(P->b_pos.as<char>() + sizeof(double) * n3);
    const bool pinned_in = N > 0 && host_pointer_is_pinned(pos) && host_pointer_is_pinned(Z);
    if (N > 0) {
      if (pinned_in) {
        CUDA_OK(cudaMemcpyAsync(d_pos, pos, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(d_Z, Z, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
      } else {
        memcpy(P->h_stage, pos, sizeof(double) * n3);
        memcpy(P->h_stage + sizeof(double) * n3, Z, sizeof(int) * (size_t)N);
        CUDA_OK(cudaMemcpyAsync(P->b_pos.p, P->h_stage, in_bytes, cudaMemcpyHostToDevice, st));
      }
    }
    bool want_grad = force || virial || local_virial;  // IPModel_GAP.f95:416-424
    const bool sharded = P->comm && comm_size(P->comm) > 1;
    const bool pinned_f = force && N > 0 && host_pointer_is_pinned(force);
    double* h_out = (double*)P->h_stage;  // [head (16) | F (3N) | local_e (N) | local_virial (9N)]
    for (int attempt = 0; attempt < 2; attempt++) {
      if (attempt == 1 && N > 0 && !pinned_in) CUDA_OK(cudaStreamSynchronize(st));  // (the staging buffer is shared by both directions)
      // with a communicator the totals arrive on every rank; local_e / local_virial are reduced only when the caller wants them
      calc_reduced(P, N, d_pos, d_Z, lattice, pbc, args_str, want_grad, P->b_packed.as<double>(), (!sharded || local_e) ? P->b_le.as<double>() : nullptr,
                   local_virial ? P->b_lv.as<double>() : nullptr, st);
      size_t o = 16;
      CUDA_OK(cudaMemcpyAsync(h_out, P->b_packed.p, sizeof(double) * 10, cudaMemcpyDeviceToHost, st));
      if (force && N > 0) CUDA_OK(cudaMemcpyAsync(pinned_f ? force : h_out + o, P->b_packed.as<double>() + 10, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
      o += n3;
      if (local_e && N > 0) { CUDA_OK(cudaMemcpyAsync(h_out + o, P->b_le.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st)); o += N; }
      if (local_virial && N > 0) CUDA_OK(cudaMemcpyAsync(h_out + o, P->b_lv.p, sizeof(double) * 9 * (size_t)N, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      comm_check(P->comm);
      // false: the speculatively sized neighbour list overflowed; repeat with the exact size.  A rank whose rows overflowed has
      // NaN-poisoned its energy (k_finalize), so after the reduction every rank of a sharded run sees it and all repeat together.
      const bool ok_local = verify_connect(P);
      if (ok_local && (!sharded || h_out[0] == h_out[0])) break;
      if (ok_local) P->row_hint = -1;
    }
    double head[10];
    memcpy(head, h_out, sizeof(head));
    {
      size_t o = 16;
      if (force && N > 0 && !pinned_f) memcpy(force, h_out + o, sizeof(double) * n3);
      o += n3;
      if (local_e && N > 0) { memcpy(local_e, h_out + o, sizeof(double) * (size_t)N); o += N; }
      if (local_virial && N > 0) memcpy(local_virial, h_out + o, sizeof(double) * 9 * (size_t)N);
    }
    if (energy) *energy = head[0];
    if (virial)
      for (int k = 0; k < 9; k++) virial[k] = head[1 + k];
    collect_timings(P);
  });
}

namespace {
// DynamicalSystem_run core on device-resident state.  reduce (may be NULL) is called after every force evaluation has been
// enqueued and must enqueue, on the same stream, the reduction of d_packed over the ranks of a partitioned run.
void md_run_impl(gap_potential* P, int N, double* d_pos, double* d_velo, const int* d_Z, const double* d_mass, const double* lattice, const int* pbc,
                 double dt, int n_steps, const char* args_str, double* d_packed, gap_reduce_fn reduce, void* reduce_ctx, double* epot, double* ekin,
                 cudaStream_t st) {
  const size_t n3 = 3 * (size_t)N;
  P->b_velo2.ensure(sizeof(double) * n3);
  P->b_acc.ensure(sizeof(double) * n3);
  P->b_ke.ensure(sizeof(double) * 128);
  P->b_le.ensure(sizeof(double) * (size_t)(N + 1));
  double* velo_cur = d_velo;
  double* velo_new = P->b_velo2.as<double>();
  const int nb = (int)((n3 + 255) / 256);
  double ke_part[128];
  auto evaluate = [&](int step) {  // calc(pot, atoms, "energy force") with the per-step neighbour-list rebuild (Potential.f95:2340-2365)
    for (int attempt = 0; attempt < 2; attempt++) {
      // partial [E | virial | F] -> totals on every rank (IPModel_GAP.f95:538-556): by the caller's hook if one is given, else by the
      // handle's communicator (gap_potential_set_comm)
      if (reduce) {
        calc_device_impl(P, N, d_pos, d_Z, lattice, pbc, args_str, true, d_packed, nullptr, nullptr, st);
        reduce(reduce_ctx, (void*)st);
      } else {
        calc_reduced(P, N, d_pos, d_Z, lattice, pbc, args_str, true, d_packed, nullptr, nullptr, st);
      }
      // the new velocities go to a second buffer: if the speculatively sized neighbour list overflowed, the evaluation is simply repeated
      k_verlet2<<<nb, 256, 0, st>>>(N, dt, d_packed + 10, d_mass, velo_cur, velo_new, P->b_acc.as<double>(), step > 0 ? 1 : 0);
      P->launches += 1;
      if (ekin) {
        k_kinetic<<<128, 256, 0, st>>>(N, d_mass, velo_new, P->b_ke.as<double>());
        P->launches += 1;
        CUDA_OK(cudaMemcpyAsync(ke_part, P->b_ke.p, sizeof(ke_part), cudaMemcpyDeviceToHost, st));
      }
      double e = 0.0;
      CUDA_OK(cudaMemcpyAsync(&e, d_packed, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      comm_check(P->comm);
      // a rank whose neighbour rows overflowed has NaN-poisoned its energy word (k_finalize): after the reduction every rank sees
      // it, so all ranks repeat together (the local verification also refreshes the row-capacity hint)
      const bool ok_local = verify_connect(P);
      if (!ok_local || e != e) {
        if (attempt == 1) throw GapError("gap_md_run: the neighbour list overflowed twice (non-finite positions or energies?)");
        if (ok_local) P->row_hint = -1;  // another rank overflowed: take the exact path together with it
        continue;
      }
      std::swap(velo_cur, velo_new);
      if (epot) epot[step] = e;
      if (ekin) {
        double t = 0.0;
        for (int k = 0; k < 128; k++) t += ke_part[k];
        ekin[step] = t;
      }
      break;
    }
  };
  evaluate(0);  // initial forces -> accelerations (Potential.f95:2348-2351)
  for (int n = 1; n <= n_steps; n++) {
    k_verlet1<<<nb, 256, 0, st>>>(N, dt, d_pos, velo_cur, P->b_acc.as<double>());
    P->launches += 1;
    evaluate(n);
  }
  if (velo_cur != d_velo) CUDA_OK(cudaMemcpyAsync(d_velo, velo_cur, sizeof(double) * n3, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
}
}  // namespace

int gap_md_run(gap_potential* P, int N, double* pos, double* velo, const int* Z, const double* mass, const double* lattice, const int* pbc,
               double dt, int n_steps, const char* args_str, double* epot, double* ekin) {
  return guard([&] {
    if (!P) throw GapError("gap_md_run: pot is NULL");
    if (N <= 0 || !pos || !velo || !Z || !mass) throw GapError("gap_md_run: N, pos, velo, Z and mass are required");
    if (n_steps < 0) throw GapError("gap_md_run: n_steps < 0");
    if (P->n_ranks != 1 && !P->comm) throw GapError("gap_md_run: a partitioned run needs a communicator (gap_potential_set_comm) or the reduction hook of gap_md_run_device");
    CUDA_OK(cudaSetDevice(P->device));
    cudaStream_t st = P->stream;
    const size_t n3 = 3 * (size_t)N;
    P->b_pos.ensure(sizeof(double) * (n3 + 3));
    P->b_Z.ensure(sizeof(int) * (size_t)(N + 1));
    P->b_velo.ensure(sizeof(double) * n3);
    P->b_mass.ensure(sizeof(double) * (size_t)N);
    P->b_packed.ensure(sizeof(double) * (10 + n3));
    CUDA_OK(cudaMemcpyAsync(P->b_pos.p, pos, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->b_velo.p, velo, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->b_Z.p, Z, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->b_mass.p, mass, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
    md_run_impl(P, N, P->b_pos.as<double>(), P->b_velo.as<double>(), P->b_Z.as<int>(), P->b_mass.as<double>(), lattice, pbc, dt, n_steps, args_str,
                P->b_packed.as<double>(), nullptr, nullptr, epot, ekin, st);
    CUDA_OK(cudaMemcpyAsync(pos, P->b_pos.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(velo, P->b_velo.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  });
}

int gap_md_run_device(gap_potential* P, int N, double* d_pos, double* d_velo, const int* d_Z, const double* d_mass, const double* lattice,
                      const int* pbc, double dt, int n_steps, const char* args_str, double* d_packed, gap_reduce_fn reduce, void* reduce_ctx,
                      double* epot, double* ekin, void* stream) {
  return guard([&] {
    if (!P) throw GapError("gap_md_run_device: pot is NULL");
    if (N <= 0 || !d_pos || !d_velo || !d_Z || !d_mass || !d_packed) throw GapError("gap_md_run_device: N, d_pos, d_velo, d_Z, d_mass and d_packed are required");
    if (n_steps < 0) throw GapError("gap_md_run_device: n_steps < 0");
    if (P->n_ranks != 1 && !reduce && !P->comm) throw GapError("gap_md_run_device: a partitioned run needs a communicator (gap_potential_set_comm) or the reduction hook");
    CUDA_OK(cudaSetDevice(P->device));
    md_run_impl(P, N, d_pos, d_velo, d_Z, d_mass, lattice, pbc, dt, n_steps, args_str, d_packed, reduce, reduce_ctx, epot, ekin,
                stream ? (cudaStream_t)stream : P->stream);
  });
}

// -----------------------------------------------------------------------------------------------------
// LAMMPS `pair_style quip` ABI (src/Potentials/quip_lammps_wrapper.f95:24-197): same symbols, same argument lists
// (everything by reference, as Fortran bind(c) passes it), so pair_quip.cpp links against libgapb200.so unchanged.
// -----------------------------------------------------------------------------------------------------
int quip_lammps_api_version(void) { return 1; }  // quip_lammps_wrapper.f95:24-28

static gap_potential* g_lammps_pot = nullptr;  // the reference keeps ONE saved Potential as well (:171)

// CUDA device of a LAMMPS rank: GAP_B200_DEVICE if set, else the node-local MPI rank (Open MPI / MVAPICH / Slurm) modulo the
// number of visible devices, so that the ranks of a multi-GPU node do not all land on device 0
static int lammps_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return 0;
  for (const char* k : {"GAP_B200_DEVICE", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "LOCAL_RANK"}) {
    const char* v = getenv(k);
    if (v && *v) {
      char* end = nullptr;
      long r = strtol(v, &end, 10);
      if (end != v && r >= 0) return (int)(r % ndev);
    }
  }
  return 0;
}

void quip_lammps_potential_initialise(int* quip_potential, int* n_quip_potential, double* quip_cutoff, char* quip_file, int* n_quip_file,
                                      char* quip_string, int* n_quip_string) {
  // two-call protocol (:176-192): first call (n == 0) builds the potential and returns the handle size in ints, the second
  // call stores the handle in the caller's integer array
  static_assert(sizeof(gap_potential*) <= 2 * sizeof(int), "handle does not fit two ints");
  if (*n_quip_potential == 0) {
    std::string file(quip_file, (size_t)*n_quip_file), args(quip_string, (size_t)*n_quip_string);
    if (g_lammps_pot) gap_potential_finalise(g_lammps_pot);
    g_lammps_pot = nullptr;
    if (gap_potential_filename_initialise(&g_lammps_pot, args.c_str(), file.c_str(), lammps_device()) != 0) {
      fprintf(stderr, "SYSTEM ABORT: quip_lammps_potential_initialise: %s\n", gap_last_error());  // the reference system_aborts here
      abort();
    }
    *n_quip_potential = 2;
  } else {
    memcpy(quip_potential, &g_lammps_pot, sizeof(gap_potential*));
  }
  *quip_cutoff = g_lammps_pot ? gap_potential_cutoff(g_lammps_pot) : 0.0;
}

void quip_lammps_wrapper(int* nlocal, int* nghost, int* atomic_numbers, int* lmptag, int* inum, int* sum_num_neigh, int* ilist, int* quip_num_neigh,
                         int* quip_neigh, double* lattice, int* quip_potential, int* n_quip_potential, double* quip_x, double* quip_e,
                         double* quip_local_e, double* quip_virial, double* quip_local_virial, double* quip_force) {
  (void)lmptag;
  const int N = *nlocal + *nghost;
  int rc = guard([&] {
    if (*n_quip_potential == 0) throw GapError("quip_lammps_wrapper: quip_potential not initialised");  // :70-72
    gap_potential* P = nullptr;
    memcpy(&P, quip_potential, sizeof(gap_potential*));
    if (!P) throw GapError("quip_lammps_wrapper: quip_potential not initialised");
    *quip_e = 0.0;
    for (int k = 0; k < 9; k++) quip_virial[k] = 0.0;
    std::fill(quip_local_e, quip_local_e + N, 0.0);
    std::fill(quip_local_virial, quip_local_virial + 9 * (size_t)N, 0.0);
    std::fill(quip_force, quip_force + 3 * (size_t)N, 0.0);
    if (*nlocal <= 0) return;  // vacuum region: nothing to do (:78, :148-154)
    CUDA_OK(cudaSetDevice(P->device));
    cudaStream_t st = P->stream;
    // LAMMPS' full list -> CSR over all atoms (rows of ghosts and of unlisted atoms stay empty); neighbour indices arrive
    // 1-based (:107-109), shifts are zero because periodic images are explicit ghosts (:112-116)
    std::vector<int> off((size_t)N + 1, 0), zc((size_t)N, -1);
    for (int ni = 0; ni < *inum; ni++) {
      int i = ilist[ni];
      if (i < 0 || i >= *nlocal) throw GapError("quip_lammps_wrapper: ilist entry outside the local atoms");
      off[(size_t)i + 1] = quip_num_neigh[ni];
      zc[i] = atomic_numbers[i];
    }
    for (int i = 0; i < N; i++) off[(size_t)i + 1] += off[i];
    if (off[N] != *sum_num_neigh) throw GapError("quip_lammps_wrapper: sum_num_neigh does not match the neighbour counts");
    std::vector<int> nj((size_t)std::max(*sum_num_neigh, 1));
    {
      size_t nn = 0;
      for (int ni = 0; ni < *inum; ni++) {
        int i = ilist[ni], w = off[i];
        for (int n = 0; n < quip_num_neigh[ni]; n++, nn++) {
          int j = quip_neigh[nn] - 1;
          if (j < 0 || j >= N) throw GapError("quip_lammps_wrapper: neighbour index out of range");
          nj[(size_t)w + n] = j;
        }
      }
    }
    const size_t nnz = (size_t)*sum_num_neigh;
    P->b_pos.ensure(sizeof(double) * 3 * (size_t)(N + 1));
    P->b_Z.ensure(sizeof(int) * (size_t)(N + 1));
    P->b_zc.ensure(sizeof(int) * (size_t)(N + 1));
    P->b_xoff.ensure(sizeof(int) * (size_t)(N + 2));
    P->b_xj.ensure(sizeof(int) * (nnz + 1));
    P->b_xs.ensure(sizeof(int) * (nnz + 1));
    P->b_packed.ensure(sizeof(double) * (10 + 3 * (size_t)N));
    P->b_le.ensure(sizeof(double) * (size_t)(N + 1));
    P->b_lv.ensure(sizeof(double) * 9 * (size_t)(N + 1));
    CUDA_OK(cudaMemcpyAsync(P->b_pos.p, quip_x, sizeof(double) * 3 * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->b_Z.p, atomic_numbers, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->b_zc.p, zc.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->b_xoff.p, off.data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
    if (nnz) CUDA_OK(cudaMemcpyAsync(P->b_xj.p, nj.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(P->b_xs.p, 0, sizeof(int) * (nnz + 1), st));
    ExtList ext{P->b_xoff.as<int>(), P->b_xj.as<int>(), P->b_xs.as<int>(), P->b_zc.as<int>(), *nlocal, (long)nnz};
    const int pbc[3] = {0, 0, 0};
    const int save_rank = P->rank, save_n = P->n_ranks;
    P->rank = 0; P->n_ranks = 1;
    try {
      calc_device_impl(P, N, P->b_pos.as<double>(), P->b_Z.as<int>(), lattice, pbc, "", true, P->b_packed.as<double>(), P->b_le.as<double>(),
                       P->b_lv.as<double>(), st, &ext);
    } catch (...) {
      P->rank = save_rank; P->n_ranks = save_n;
      throw;
    }
    P->rank = save_rank; P->n_ranks = save_n;
    double head[10];
    CUDA_OK(cudaMemcpyAsync(head, P->b_packed.p, sizeof(head), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(quip_force, P->b_packed.as<double>() + 10, sizeof(double) * 3 * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(quip_local_e, P->b_le.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(quip_local_virial, P->b_lv.p, sizeof(double) * 9 * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    *quip_e = head[0];
    for (int k = 0; k < 9; k++) quip_virial[k] = head[1 + k];
  });
  if (rc != 0) {  // the Fortran wrapper has no error argument: it system_aborts
    fprintf(stderr, "SYSTEM ABORT: %s\n", gap_last_error());
    abort();
  }
}

int gap_potential_set_timing(gap_potential* P, int on) {
  return guard([&] {
    if (!P) throw GapError("gap_potential_set_timing: pot is NULL");
    P->timing = on == 2 ? 2 : (on != 0 ? 1 : 0);
    P->ev_used = 0;
  });
}

int gap_potential_last_timings(gap_potential* P, double* ms8) {
  if (!P || !ms8) return 1;
  if (P->ev_used >= 2) {  // events of the last calc (host- or device-pointer entry): wait for the last one, then read
    cudaSetDevice(P->device);
    cudaEventSynchronize(P->ev[P->ev_used - 1]);
    collect_timings(P);
  }
  for (int k = 0; k < 8; k++) ms8[k] = P->last_ms[k];
  return 0;
}

int gap_b200_wrapper_simple(const char* param_filename, const int* N, const double* lattice, const int* Z, const double* pos, double* energy,
                            double* force, double* virial) {
  gap_potential* P = nullptr;
  int rc = gap_potential_filename_initialise(&P, "", param_filename, 0);
  if (rc) return rc;
  int pbc[3] = {1, 1, 1};
  rc = gap_potential_calc(P, *N, pos, Z, lattice, pbc, "", energy, nullptr, force, virial, nullptr);
  gap_potential_finalise(P);
  return rc;
}

int gap_calc_connect(gap_potential* P, int N, const double* pos, const double* lattice, const int* pbc, double cutoff, int* n_entries) {
  return guard([&] {
    if (!P) throw GapError("gap_calc_connect: pot is NULL");
    CUDA_OK(cudaSetDevice(P->device));
    P->b_pos.ensure(sizeof(double) * 3 * (size_t)(N + 1));
    if (N > 0) CUDA_OK(cudaMemcpyAsync(P->b_pos.p, pos, sizeof(double) * 3 * (size_t)N, cudaMemcpyHostToDevice, P->stream));
    build_connect(P, N, 0, N, P->b_pos.as<double>(), lattice, pbc, cutoff, true, false, P->stream);
    CUDA_OK(cudaStreamSynchronize(P->stream));
    if (n_entries) *n_entries = P->conn_nnz;
  });
}

int gap_get_connect(gap_potential* P, int* offsets, int* j, int* shift, double* distance) {
  return guard([&] {
    if (!P) throw GapError("gap_get_connect: pot is NULL");
    CUDA_OK(cudaSetDevice(P->device));
    int N = P->conn_N, nnz = P->conn_nnz;
    if (offsets) CUDA_OK(cudaMemcpy(offsets, P->b_off.p, sizeof(int) * (N + 1), cudaMemcpyDeviceToHost));
    if (nnz > 0) {
      if (j) CUDA_OK(cudaMemcpy(j, P->b_j.p, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
      if (shift) {
        std::vector<int> packed(nnz);
        CUDA_OK(cudaMemcpy(packed.data(), P->b_s.p, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
        for (int k = 0; k < nnz; k++) unpack_shift(packed[k], shift[3 * k], shift[3 * k + 1], shift[3 * k + 2]);
      }
      if (distance) {
        if (P->b_d.cap < sizeof(double) * (size_t)nnz) throw GapError("gap_get_connect: distances were not stored (call gap_calc_connect first)");
        CUDA_OK(cudaMemcpy(distance, P->b_d.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost));
      }
    }
  });
}

int gap_descriptor_calc(gap_potential* P, int i_coord, int N, const double* pos, const int* Z, const double* lattice, const int* pbc, int* n_desc,
                        int* d_out, double* x, int* ci) {
  return guard([&] {
    if (!P) throw GapError("gap_descriptor_calc: pot is NULL");
    if (i_coord < 0 || i_coord >= (int)P->cd.size()) throw GapError("gap_descriptor_calc: coordinate index out of range");
    const CoordDev& cd = P->cd[i_coord];
    if (cd.kind != DESC_SOAP) throw GapError("gap_descriptor_calc: only soap coordinates have a descriptor-level entry point");
    CUDA_OK(cudaSetDevice(P->device));
    cudaStream_t st = P->stream;
    P->b_pos.ensure(sizeof(double) * 3 * (size_t)(N + 1));
    P->b_Z.ensure(sizeof(int) * (size_t)(N + 1));
    if (N > 0) {
      CUDA_OK(cudaMemcpyAsync(P->b_pos.p, pos, sizeof(double) * 3 * (size_t)N, cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaMemcpyAsync(P->b_Z.p, Z, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    }
    int n_ub = select_centres(P, cd, P->b_Z.as<int>(), 0, N, st);
    int nc = 0;
    if (n_ub > 0) {
      CUDA_OK(cudaMemcpyAsync(&nc, nc_dev(P, n_ub), sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
    }
    const bool glob = cd.general && cd.gen.global_mode;
    if (n_desc) *n_desc = glob ? (nc > 0 ? 1 : 0) : nc;  // average=T: one descriptor per configuration (ci returns its first centre)
    if (d_out) *d_out = cd.h.d;
    if (!x) return;
    build_connect(P, N, 0, N, P->b_pos.as<double>(), lattice, pbc, cd.h.cutoff, false, false, st);
    Lattice9 lat;
    for (int k = 0; k < 9; k++) lat.v[k] = lattice[k];
    soap_forward_stage(P, cd, n_ub, P->b_pos.as<double>(), P->b_Z.as<int>(), lat, st);
    if (nc > 0) {
      const int nrow = glob ? 1 : nc;
      CUDA_OK(cudaMemcpy2DAsync(x, sizeof(double) * cd.h.d, P->b_x.p, sizeof(double) * cd.d_pad, sizeof(double) * cd.h.d, nrow, cudaMemcpyDeviceToHost, st));
      if (ci) CUDA_OK(cudaMemcpyAsync(ci, P->b_centres.p, sizeof(int) * nrow, cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(cudaStreamSynchronize(st));
  });
}

int gap_gp_predict(gap_potential* P, int i_coord, int n, const double* x, double* e, double* grad) {
  return guard([&] {
    if (!P) throw GapError("gap_gp_predict: pot is NULL");
    if (i_coord < 0 || i_coord >= (int)P->cd.size()) throw GapError("gap_gp_predict: coordinate index out of range");
    const CoordDev& cd = P->cd[i_coord];
    if (cd.kind != DESC_SOAP) throw GapError("gap_gp_predict: dot_product (soap) coordinates only");
    if (n <= 0) return;
    CUDA_OK(cudaSetDevice(P->device));
    cudaStream_t st = P->stream;
    P->ev_used = 0;
    mark(P, st, -1);
    int n_pad = round_up(n, COV_BM);
    P->b_x.ensure(sizeof(double) * (size_t)n_pad * cd.d_pad);
    CUDA_OK(cudaMemsetAsync(P->b_x.p, 0, sizeof(double) * (size_t)n_pad * cd.d_pad, st));
    CUDA_OK(cudaMemcpy2DAsync(P->b_x.p, sizeof(double) * cd.d_pad, x, sizeof(double) * cd.h.d, sizeof(double) * cd.h.d, n, cudaMemcpyHostToDevice, st));
    P->b_centres.ensure(sizeof(int) * (size_t)(n + 1));
    k_iota<<<(n + 255) / 256, 256, 0, st>>>(P->b_centres.as<int>(), n);
    P->b_le.ensure(sizeof(double) * (size_t)(n + 1));
    CUDA_OK(cudaMemsetAsync(P->b_le.p, 0, sizeof(double) * (size_t)(n + 1), st));
    covariance_stage(P, cd, n, nullptr, grad != nullptr, false, st);
    int launches = 1;
    launch_energy_rows(P->b_epart.as<double>(), P->g_tiles_n, P->b_centres.as<int>(), nullptr, n, 1.0, P->b_le.as<double>(), st, &launches);
    P->launches += launches;
    if (e) CUDA_OK(cudaMemcpyAsync(e, P->b_le.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (grad)
      CUDA_OK(cudaMemcpy2DAsync(grad, sizeof(double) * cd.h.d, P->b_gvec.p, sizeof(double) * cd.dn_pad, sizeof(double) * cd.h.d, n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
  });
}

int gap_model_describe(const char* args_str, const char* param_str, const char* base_dir, char* buf, size_t n) {
  return guard([&] {
    if (!param_str || !buf || n == 0) throw GapError("gap_model_describe: bad arguments");
    GapModel m = load_gap_model(args_str ? args_str : "", param_str, base_dir && *base_dir ? base_dir : ".");
    std::ostringstream os;
    os.precision(17);
    os << "label " << m.label << "\nxml_version " << m.xml_version << "\ncutoff " << m.cutoff << "\nE_scale " << m.E_scale << "\nn_coordinate "
       << m.coord.size() << "\n";
    os << "e0";
    for (int z = 0; z < 128; z++)
      if (m.e0[z] != 0.0) os << " " << z << ":" << m.e0[z];
    os << "\n";
    for (size_t i = 0; i < m.coord.size(); i++) {
      const Coordinate& c = m.coord[i];
      double sx = 0, sa = 0, sc = 0;
      for (size_t k = 0; k < c.sparseX.size(); k++) sx += c.sparseX[k] * (double)((k % 7) + 1);
      for (double v : c.alpha) sa += v;
      for (double v : c.sparseCutoff) sc += v;
      os << "coordinate " << i << " kind " << c.kind << " covariance_type " << c.covariance_type << " d " << c.d << " M " << c.M << " delta " << c.delta
         << " f0 " << c.f0 << " zeta " << c.zeta << " theta0 " << (c.theta.empty() ? 0.0 : c.theta[0]) << " cutoff " << c.cutoff() << " sparseX_wsum " << sx
         << " alpha_sum " << sa << " sparseCutoff_sum " << sc << "\n";
      if (c.kind == DESC_SOAP) {
        const SoapSpec& s = c.soap;
        os << "soap " << i << " l_max " << s.l_max << " n_max " << s.n_max << " n_species " << s.n_species << " n_Z " << s.n_Z << " cras "
           << (int)s.central_reference_all_species << " normalise " << (int)s.normalise << " two_lp1 " << (int)s.do_two_l_plus_one << " alpha " << s.alpha
           << " ctw " << s.cutoff_transition_width << " central_weight " << s.central_weight << " sigma0 " << s.covariance_sigma0 << "\n";
        os << "species_Z";
        for (int v : s.species_Z) os << " " << v;
        os << "\nZ";
        for (int v : s.Z) os << " " << v;
        os << "\nr_basis";
        for (double v : s.r_basis) os << " " << v;
        os << "\ntransform_basis";
        for (double v : s.transform_basis) os << " " << v;
        os << "\ncholesky_overlap";
        for (double v : s.cholesky_overlap) os << " " << v;
        os << "\n";
        if (s.general) {  // the general path's derived tables (compression modes, GTO / POLY radial maps)
          os << "soap_general " << i << " radial_basis " << s.radial_basis << " n_grid " << s.n_grid << " Ka " << s.Ka << " Kb " << s.Kb << " n_pairs "
             << s.pair_ia.size() << "\nr_grid";
          for (double v : s.r_grid) os << " " << v;
          os << "\nP";
          for (double v : s.P) os << " " << v;
          os << "\nc0";
          for (double v : s.c0) os << " " << v;
          os << "\nW1";
          for (double v : s.W1) os << " " << v;
          os << "\nW2";
          for (double v : s.W2) os << " " << v;
          os << "\npairs";
          for (size_t k = 0; k < s.pair_ia.size(); k++) os << " " << s.pair_ia[k] << ":" << s.pair_jb[k] << ":" << s.pair_fac[k];
          os << "\n";
        }
      } else if (c.kind == DESC_ANGLE_3B) {
        os << "angle_3b " << i << " Z " << c.a3b.Zc << " Z1 " << c.a3b.Z1 << " Z2 " << c.a3b.Z2 << " ctw " << c.a3b.cutoff_transition_width << "\n";
      } else {
        os << "distance_2b " << i << " Z1 " << c.d2b.Z1 << " Z2 " << c.d2b.Z2 << " ctw " << c.d2b.cutoff_transition_width << "\n";
      }
    }
    std::string s = os.str();
    if (s.size() + 1 > n) throw GapError("gap_model_describe: buffer too small (need " + std::to_string(s.size() + 1) + ")");
    memcpy(buf, s.c_str(), s.size() + 1);
  });
}

}  // extern "C"
