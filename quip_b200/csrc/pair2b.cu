// pair2b.cu -- distance_2b descriptor + ARD squared-exponential covariance + its scatter, fused.
//
// Replaces distance_2b_calc (src/GAP/descriptors.f95:4615-4783), the ARD_SE branch of gpCoordinates_Predict
// (src/GAP/gp_predict.f95:3692-3697, 3787-3817, n_permutations = 1) and the scatter of those instances in
// IPModel_GAP_Calc (src/Potentials/IPModel_GAP.f95:452-499).
//
// The reference creates one descriptor instance per ORDERED pair (i, n) and scatters half of its energy to each
// end.  Summed per RECEIVING atom the contributions of instance (i->j) and of its mirror (j->i) are
//     local_e(i) += e(r) fc(r)            F_i += 2 phi u_ij           W_i -= phi d_ij (x) u_ij
// with phi = g(r) fc(r) + e(r) fc'(r), e = sum_s alpha_s k_s, g = de/dr, so one thread per (atom, neighbour)
// needs no atomics and no descriptor storage at all.  One warp per atom, lanes over neighbours, fixed-order
// warp reductions (deterministic).
#include "gap_device.cuh"

namespace gapb200 {

namespace {
constexpr double PI_D = 3.14159265358979323846264338327950288;
constexpr int WPB = 4;  // warps (= atoms) per block

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(WPB * 32) k_pair2b(Pair2bDev p, int first, int last, const int* __restrict__ nbr_off, const int* __restrict__ nbr_end,
                                                     const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                     const double* __restrict__ pos, const int* __restrict__ Z, const int* __restrict__ Zc,
                                                     int scatter, Lattice9 lat, double e_scale,
                                                     int do_grad, double* __restrict__ local_e, double* __restrict__ force,
                                                     double* __restrict__ vir_part, double* __restrict__ local_virial) {
  __shared__ double sX[64 * PAIR2B_MAX_EXP], sA[64], sC[64];
  __shared__ double svir[WPB][9];
  const int ne = p.n_exp;
  for (int k = threadIdx.x; k < p.M && k < 64; k += blockDim.x) {
    for (int q = 0; q < ne; q++) sX[k * ne + q] = p.sparseX[(size_t)k * ne + q];
    sA[k] = p.alpha[k];
    sC[k] = p.scut[k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = first + blockIdx.x * WPB + w;
  double e_acc = 0, f0 = 0, f1 = 0, f2 = 0, v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  // scatter = 0: every atom of the configuration is a centre somewhere (whole system, or a partition whose partial
  //   results are summed): the instance (j,i) mirrors (i,j), so atom i takes both halves and nothing is scattered.
  // scatter = 1 (atom mask, IPModel_GAP.f95:344-346): only masked atoms are centres; each instance (i,j) gives atom j its own
  //   half of the energy, the opposite force and the virial row, exactly as the reference's scatter loop (:454-491).
  if (i < last && Zc[i] >= 0) {
    const int Zi = Z[i];
    const bool Zi1 = (p.Z1 == 0) || (Zi == p.Z1), Zi2 = (p.Z2 == 0) || (Zi == p.Z2);
    if (Zi1 || Zi2) {
      for (int q = nbr_off[i] + lane; q < nbr_end[i]; q += 32) {
        int j = nbr_j[q], s0, s1, s2;
        unpack_shift(nbr_s[q], s0, s1, s2);
        double dd[3];
        image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)j, lat.v, s0, s1, s2, dd);
        double r = norm_nofma(dd);
        if (r >= p.cutoff) continue;  // descriptors.f95:4729
        int Zj = Z[j];
        bool Zj1 = (p.Z1 == 0) || (Zj == p.Z1), Zj2 = (p.Z2 == 0) || (Zj == p.Z2);
        if (!((Zi1 && Zj2) || (Zi2 && Zj1))) continue;  // :4733
        if (p.intra_mode && p.resid) {  // :4735-4738
          const bool same = p.resid[i] == p.resid[j];
          if ((p.intra_mode == 1 && !same) || (p.intra_mode == 2 && same)) continue;
        }
        // descriptor data_k = r^exponents_k and d data_k / dr (:4750, 4768); ARD_SE over the n_exp components (gp_predict.f95:3795-3816)
        double xk[PAIR2B_MAX_EXP], dxk[PAIR2B_MAX_EXP];
#pragma unroll
        for (int q = 0; q < PAIR2B_MAX_EXP; q++) {
          xk[q] = dxk[q] = 0.0;
          if (q < ne) {
            if (p.exponents[q] == 1.0) { xk[q] = r; dxk[q] = 1.0; }
            else { xk[q] = pow(r, p.exponents[q]); dxk[q] = p.exponents[q] * pow(r, p.exponents[q] - 1.0); }
          }
        }
        double e = 0.0, g = 0.0;
        for (int s = 0; s < p.M; s++) {
          const double as = s < 64 ? sA[s] : p.alpha[s], cs = s < 64 ? sC[s] : p.scut[s];
          double r2 = 0.0, gs = 0.0;
#pragma unroll
          for (int q = 0; q < PAIR2B_MAX_EXP; q++)
            if (q < ne) {
              const double xs = s < 64 ? sX[s * ne + q] : p.sparseX[(size_t)s * ne + q];
              const double t = (xs - xk[q]) * p.inv_theta[q];
              r2 += t * t;
              gs += t * p.inv_theta[q] * dxk[q];
            }
          const double ce = p.delta2 * exp(-0.5 * r2);
          e += as * (ce + p.f02) * cs;
          g += as * ce * gs * cs;
        }
        double fc, dfc;  // coordination_function, linearalgebra.f95:7488-7516
        if (r > p.cutoff - p.ctw) {
          double sn, cn;
          sincos(PI_D * (r - p.cutoff + p.ctw) / p.ctw, &sn, &cn);
          fc = 0.5 * (cn + 1.0);
          dfc = -0.5 * PI_D * sn / p.ctw;
        } else { fc = 1.0; dfc = 0.0; }
        if (p.tail_exponent != 0) {  // covariance_cutoff = f_cut tail (:4743-4747, 4759-4764)
          const double ef = erf(p.tail_range * r);
          const double tail = pow(ef / r, (double)p.tail_exponent);
          const double dtail = tail * p.tail_exponent * (2.0 * p.tail_range * exp(-p.tail_range * p.tail_range * r * r) / sqrt(PI_D) / ef - 1.0 / r);
          dfc = dfc * tail + fc * dtail;
          fc = fc * tail;
        }
        if (scatter) {
          e_acc += 0.5 * e * fc;
          if (local_e) atomicAdd(&local_e[j], 0.5 * e_scale * e * fc);
        } else {
          e_acc += e * fc;
        }
        if (do_grad) {
          double phi = (g * fc + e * dfc) * e_scale, rinv = 1.0 / r;
          double u0 = dd[0] * rinv, u1 = dd[1] * rinv, u2 = dd[2] * rinv;
          double wv[9] = {dd[0] * u0, dd[1] * u0, dd[2] * u0, dd[0] * u1, dd[1] * u1, dd[2] * u1, dd[0] * u2, dd[1] * u2, dd[2] * u2};
          if (scatter) {
            f0 += phi * u0; f1 += phi * u1; f2 += phi * u2;
            if (force) {
              atomicAdd(&force[3 * (size_t)j + 0], -phi * u0);
              atomicAdd(&force[3 * (size_t)j + 1], -phi * u1);
              atomicAdd(&force[3 * (size_t)j + 2], -phi * u2);
            }
            if (local_virial)
#pragma unroll
              for (int k = 0; k < 9; k++) atomicAdd(&local_virial[9 * (size_t)j + k], -phi * wv[k]);
          } else {
            f0 += 2.0 * phi * u0; f1 += 2.0 * phi * u1; f2 += 2.0 * phi * u2;
          }
#pragma unroll
          for (int k = 0; k < 9; k++) v[k] -= phi * wv[k];
        }
      }
    }
  }
  e_acc = wsum(e_acc);
  if (do_grad) {
    f0 = wsum(f0); f1 = wsum(f1); f2 = wsum(f2);
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = wsum(v[k]);
  }
  if (lane == 0) {
    if (i < last) {
      if (scatter) {  // other warps scatter into row i concurrently
        if (local_e) atomicAdd(&local_e[i], e_scale * e_acc);
        if (do_grad && force) {
          atomicAdd(&force[3 * (size_t)i + 0], f0); atomicAdd(&force[3 * (size_t)i + 1], f1); atomicAdd(&force[3 * (size_t)i + 2], f2);
        }
        // the virial of an instance belongs to its row n = 1 (atom j, scattered above); row n = 0 has zero displacement
      } else {
        if (local_e) local_e[i] += e_scale * e_acc;
        if (do_grad && force) {
          // other kernels scatter into force[] with atomics on the same stream; plain RMW is safe because kernels
          // of one calc are stream-ordered and this kernel owns row i exclusively
          force[3 * (size_t)i + 0] += f0; force[3 * (size_t)i + 1] += f1; force[3 * (size_t)i + 2] += f2;
        }
        if (do_grad && local_virial)
#pragma unroll
          for (int k = 0; k < 9; k++) local_virial[9 * (size_t)i + k] += v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) svir[w][k] = (i < last && do_grad) ? v[k] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 9 && vir_part) {
    int k = threadIdx.x;
    vir_part[9 * (size_t)blockIdx.x + k] = (svir[0][k] + svir[1][k]) + (svir[2][k] + svir[3][k]);
  }
}
}  // namespace

void launch_pair2b(Pair2bDev p, int first, int last, const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z,
                   const int* Zc, int scatter, Lattice9 lat, double e_scale, int do_grad, double* local_e, double* force, double* vir_part, double* local_virial,
                   cudaStream_t st, int* launches, int* n_blocks_out) {
  int n = last - first;
  int nb = (n + WPB - 1) / WPB;
  *n_blocks_out = nb;
  if (nb <= 0) return;
  k_pair2b<<<nb, WPB * 32, 0, st>>>(p, first, last, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, Zc, scatter, lat, e_scale, do_grad, local_e, force, vir_part,
                                    local_virial);
  *launches += 1;
}

}  // namespace gapb200
