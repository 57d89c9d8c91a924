// angle3b.cu -- angle_3b descriptor + ARD squared-exponential covariance + its scatter, fused.
//
// Replaces angle_3b_calc (src/GAP/descriptors.f95:4932-5112), the ARD_SE branch of gpCoordinates_Predict
// (src/GAP/gp_predict.f95:3692-3697, 3787-3817, n_permutations = 1: descriptors.f95:12452-12466) and the scatter of those
// instances in IPModel_GAP_Calc (src/Potentials/IPModel_GAP.f95:452-499).
//
// The reference creates one instance per centre i and ORDERED pair (n, m /= n) of its neighbours j, k inside the cutoff:
//     x = (r_ij + r_ik, (r_ij - r_ik)^2, r_jk),  covariance_cutoff = fc(r_ij) fc(r_ik),  ci = (i)
// x and the cutoff are symmetric in (j, k), so instance (m, n) repeats instance (n, m) with the roles of j and k swapped: the energy of
// the two is twice one of them and so is the force each of them puts on every one of the three atoms.
//   default       one warp per centre, lanes over UNORDERED pairs n < m, everything doubled; forces on j and k by FP64 atomics
//   deterministic one lane per neighbour n, loop over all m /= n (the reference's own double loop); the lane keeps twice the force that
//                 instance (n, m) puts on j_n (= its share as j of (n, m) plus its share as k of (m, n)) and stores the sum at the list
//                 slot of n (fpair); no atomics on forces, fixed summation order
// The row is first compacted to the entries inside this descriptor's cutoff with a usable species (warp ballot) into an int scratch
// parallel to the list; distances and cutoff factors are recomputed per pair (a few dozen flops against M exponentials).
// The sparse points (divided by theta, with alpha_s sparseCutoff_s delta^2) sit in shared memory; every lane walks them for its own pair.
#include "gap_device.cuh"

namespace gapb200 {

namespace {
constexpr double PI_D = 3.14159265358979323846264338327950288;
constexpr int WPB = 4;  // warps (= centres) per block
constexpr int A3_SMEM_MAX_M = 5120;  // 4 doubles per sparse point in shared memory (160 KiB); larger models read the table through L1

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Nb {
  double d[3], r, fc, dfc;
  int j;
  bool z1, z2;
};

__device__ __forceinline__ void load_nb(const Angle3bDev& p, int q, int i, const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                        const double* __restrict__ pos, const int* __restrict__ Z, const Lattice9& lat, Nb& b) {
  b.j = nbr_j[q];
  int s0, s1, s2;
  unpack_shift(nbr_s[q], s0, s1, s2);
  image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)b.j, lat.v, s0, s1, s2, b.d);
  b.r = norm_nofma(b.d);
  const int Zj = Z[b.j];
  b.z1 = (p.Z1 == 0) || (Zj == p.Z1);
  b.z2 = (p.Z2 == 0) || (Zj == p.Z2);
  if (b.r > p.cutoff - p.ctw) {  // coordination_function, linearalgebra.f95:7488-7516
    double sn, cn;
    sincos(PI_D * (b.r - p.cutoff + p.ctw) / p.ctw, &sn, &cn);
    b.fc = 0.5 * (cn + 1.0);
    b.dfc = -0.5 * PI_D * sn / p.ctw;
  } else { b.fc = 1.0; b.dfc = 0.0; }
}

// one instance (j = a, k = b): energy e cc and the forces f_gp on j and k (IPModel_GAP.f95:479-480); f_gp on i is minus their sum
__device__ __forceinline__ void eval_instance(const Angle3bDev& p, const double* __restrict__ T, const Nb& a, const Nb& b, int do_grad, double& ecc,
                                              double* fj, double* fk) {
  double djk[3] = {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]};
  const double rjk = norm_nofma(djk);
  const double dr = a.r - b.r;
  const double x0 = (a.r + b.r) * p.inv_theta[0], x1 = dr * dr * p.inv_theta[1], x2 = rjk * p.inv_theta[2];  // descriptors.f95:5066-5068
  double e = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
  for (int s = 0; s < p.M; s++) {  // gp_predict.f95:3795-3816
    const double t0 = T[4 * s] - x0, t1 = T[4 * s + 1] - x1, t2 = T[4 * s + 2] - x2;
    const double ce = T[4 * s + 3] * exp(-0.5 * (t0 * t0 + t1 * t1 + t2 * t2));
    e += ce;
    g0 += ce * t0; g1 += ce * t1; g2 += ce * t2;
  }
  e += p.e_f0;
  const double cc = a.fc * b.fc;  // :5072
  ecc = e * cc;
  if (!do_grad) return;
  g0 *= p.inv_theta[0]; g1 *= p.inv_theta[1]; g2 *= p.inv_theta[2];
  // grad_data(:,:,1) = (u_ij, 2 (r_ij - r_ik) u_ij, u_jk), grad_data(:,:,2) = (u_ik, -2 (r_ij - r_ik) u_ik, -u_jk) (:5090-5104),
  // grad_covariance_cutoff = dfc_j fc_k u_ij and dfc_k fc_j u_ik
  const double Aj = (g0 + 2.0 * dr * g1) * cc + e * a.dfc * b.fc;
  const double Ak = (g0 - 2.0 * dr * g1) * cc + e * b.dfc * a.fc;
  const double B = g2 * cc / rjk;
  const double ia = Aj / a.r, ib = Ak / b.r;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    fj[k] = ia * a.d[k] + B * djk[k];
    fk[k] = ib * b.d[k] - B * djk[k];
  }
}

template <bool DET>
__global__ void __launch_bounds__(WPB * 32) k_angle3b(Angle3bDev p, int use_smem, int first, int last, const int* __restrict__ nbr_off,
                                                      const int* __restrict__ nbr_end, const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                      const double* __restrict__ pos, const int* __restrict__ Z, const int* __restrict__ Zc, Lattice9 lat,
                                                      double e_scale, int do_grad, double* __restrict__ local_e, double* __restrict__ force,
                                                      double* __restrict__ fpair, double* __restrict__ vir_part, double* __restrict__ local_virial,
                                                      int* __restrict__ cidx) {
  extern __shared__ double sT[];
  __shared__ double svir[WPB][9];
  const double* T = p.table;
  if (use_smem) {
    for (int k = threadIdx.x; k < 4 * p.M; k += blockDim.x) sT[k] = p.table[k];
    T = sT;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = first + blockIdx.x * WPB + w;
  double e_acc = 0, fi[3] = {0, 0, 0}, v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (i < last && Zc[i] >= 0 && (p.Zc == 0 || Z[i] == p.Zc)) {  // descriptors.f95:5025-5029
    // compact the row: entries inside the cutoff (:5034, 5047) whose species can take either role
    const int q0 = nbr_off[i], q1 = nbr_end[i];
    int nc = 0;
    for (int qb = q0; qb < q1; qb += 32) {
      const int q = qb + lane;
      bool keep = false;
      if (q < q1) {
        const int j = nbr_j[q];
        int s0, s1, s2;
        unpack_shift(nbr_s[q], s0, s1, s2);
        double dd[3];
        image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)j, lat.v, s0, s1, s2, dd);
        const int Zj = Z[j];
        keep = norm_nofma(dd) < p.cutoff && ((p.Z1 == 0) || (Zj == p.Z1) || (p.Z2 == 0) || (Zj == p.Z2));
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) cidx[q0 + nc + __popc(m & ((1u << lane) - 1u))] = q;
      nc += __popc(m);
    }
    __syncwarp();
    if (!DET) {
      const int npair = nc * (nc - 1) / 2;
      for (int t = lane; t < npair; t += 32) {
        int m = (int)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
        while (m * (m - 1) / 2 > t) m--;
        while ((m + 1) * m / 2 <= t) m++;
        const int n = t - m * (m - 1) / 2;  // 0 <= n < m < nc
        const int qn = cidx[q0 + n], qm = cidx[q0 + m];
        Nb a, b;
        load_nb(p, qn, i, nbr_j, nbr_s, pos, Z, lat, a);
        load_nb(p, qm, i, nbr_j, nbr_s, pos, Z, lat, b);
        if (!((b.z1 && a.z2) || (b.z2 && a.z1))) continue;  // :5052
        double ecc, fj[3], fk[3];
        eval_instance(p, T, a, b, do_grad, ecc, fj, fk);
        e_acc += 2.0 * ecc;
        if (do_grad) {
          const double sc = 2.0 * e_scale;
#pragma unroll
          for (int k = 0; k < 3; k++) { fj[k] *= sc; fk[k] *= sc; fi[k] += fj[k] + fk[k]; }
          if (force) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
              atomicAdd(&force[3 * (size_t)a.j + k], -fj[k]);
              atomicAdd(&force[3 * (size_t)b.j + k], -fk[k]);
            }
          }
          // virial_in(:,:,j) -= (pos_j - pos_i) (x) f_gp, column-major (alpha + 3 beta)
#pragma unroll
          for (int bb = 0; bb < 3; bb++)
#pragma unroll
            for (int aa = 0; aa < 3; aa++) {
              const double wj = a.d[aa] * fj[bb], wk = b.d[aa] * fk[bb];
              v[aa + 3 * bb] -= wj + wk;
              if (local_virial) {
                atomicAdd(&local_virial[9 * (size_t)a.j + aa + 3 * bb], -wj);
                atomicAdd(&local_virial[9 * (size_t)b.j + aa + 3 * bb], -wk);
              }
            }
        }
      }
    } else {
      for (int n = lane; n < nc; n += 32) {
        const int qn = cidx[q0 + n];
        Nb a;
        load_nb(p, qn, i, nbr_j, nbr_s, pos, Z, lat, a);
        double fn[3] = {0, 0, 0};
        for (int m = 0; m < nc; m++) {
          if (m == n) continue;  // :5044
          Nb b;
          load_nb(p, cidx[q0 + m], i, nbr_j, nbr_s, pos, Z, lat, b);
          if (!((b.z1 && a.z2) || (b.z2 && a.z1))) continue;
          double ecc, fj[3], fk[3];
          eval_instance(p, T, a, b, do_grad, ecc, fj, fk);
          e_acc += ecc;
          if (do_grad)
#pragma unroll
            for (int k = 0; k < 3; k++) fn[k] += 2.0 * e_scale * fj[k];
        }
        if (do_grad) {
#pragma unroll
          for (int k = 0; k < 3; k++) {
            fi[k] += fn[k];
            if (fpair) fpair[3 * (size_t)qn + k] = -fn[k];
          }
#pragma unroll
          for (int bb = 0; bb < 3; bb++)
#pragma unroll
            for (int aa = 0; aa < 3; aa++) {
              const double wj = a.d[aa] * fn[bb];
              v[aa + 3 * bb] -= wj;
              if (local_virial) atomicAdd(&local_virial[9 * (size_t)a.j + aa + 3 * bb], -wj);
            }
        }
      }
    }
  }
  e_acc = wsum(e_acc);
  if (do_grad) {
#pragma unroll
    for (int k = 0; k < 3; k++) fi[k] = wsum(fi[k]);
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = wsum(v[k]);
  }
  if (lane == 0) {
    if (i < last) {
      if (local_e) local_e[i] += e_scale * e_acc;  // ci = (i): the whole instance energy stays on the centre (:5069, IPModel_GAP.f95:454-459)
      if (do_grad && force) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (DET) force[3 * (size_t)i + k] = fi[k];  // per-atom buffer of the centres' own sums, added by k_det_gather
          else atomicAdd(&force[3 * (size_t)i + k], fi[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) svir[w][k] = (i < last && do_grad) ? v[k] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 9 && vir_part) {
    const int k = threadIdx.x;
    vir_part[9 * (size_t)blockIdx.x + k] = (svir[0][k] + svir[1][k]) + (svir[2][k] + svir[3][k]);
  }
}
}  // namespace

void launch_angle3b(Angle3bDev p, int first, int last, const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos,
                    const int* Z, const int* Zc, Lattice9 lat, double e_scale, int do_grad, double* local_e, double* force, double* fpair, double* vir_part,
                    double* local_virial, int* cidx, cudaStream_t st, int* launches, int* n_blocks_out) {
  const int n = last - first;
  const int nb = (n + WPB - 1) / WPB;
  *n_blocks_out = nb;
  if (nb <= 0) return;
  const int use_smem = p.M <= A3_SMEM_MAX_M ? 1 : 0;
  const size_t smem = use_smem ? sizeof(double) * 4 * (size_t)(p.M > 0 ? p.M : 1) : 0;
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(k_angle3b<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 4 * A3_SMEM_MAX_M));
    cudaFuncSetAttribute(k_angle3b<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 4 * A3_SMEM_MAX_M));
  }
  if (fpair)
    k_angle3b<true><<<nb, WPB * 32, smem, st>>>(p, use_smem, first, last, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, Zc, lat, e_scale, do_grad, local_e, force, fpair,
                                                vir_part, local_virial, cidx);
  else
    k_angle3b<false><<<nb, WPB * 32, smem, st>>>(p, use_smem, first, last, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, Zc, lat, e_scale, do_grad, local_e, force,
                                                 fpair, vir_part, local_virial, cidx);
  *launches += 1;
}

}  // namespace gapb200
