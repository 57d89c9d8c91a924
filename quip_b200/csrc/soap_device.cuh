// soap_device.cuh -- device helpers shared by the SOAP kernels (soap.cu: the specialised / DMMA kernels of the reference's "original"
// power spectrum; soap_general.cu: the general path for the compression modes and the GTO / POLY radial bases): cutoff function,
// the radial recursion, real spherical harmonics and their gradients.  Everything has internal linkage.
#pragma once
#include "gap_device.cuh"

namespace gapb200 {

namespace {

constexpr int NT = 128;    // threads per CTA
constexpr int NBCAP = 128; // CSR entries examined per pass (= compacted list capacity)
constexpr int NW = NT / 32;
constexpr int TNF = 16;    // neighbours per tile, forward (multiple of 4: K steps of the DMMA)
constexpr int TNA = 8;     // neighbours per tile, adjoint (= one DMMA N tile)
constexpr int LC = SOAP_LMAX_CAP;
constexpr double PI_D = 3.14159265358979323846264338327950288;

__constant__ double c_dblfact[LC + 1] = {1., 1., 3., 15., 105., 945., 10395., 135135., 2027025., 34459425., 654729075., 13749310575., 316234143225.};
__constant__ double c_invint[LC + 2] = {0., 1., 1. / 2, 1. / 3, 1. / 4, 1. / 5, 1. / 6, 1. / 7, 1. / 8, 1. / 9, 1. / 10, 1. / 11, 1. / 12, 1. / 13};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// fixed-order block sum (NT = 128 = 4 warps); result broadcast to all threads
__device__ __forceinline__ double block_sum(double v, double* red /* >= 4 doubles */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}

__device__ __forceinline__ int species_of(const SoapDev* sp, int Zj) {  // species_map, descriptors.f95:7783-7790 ; -1 = ignored
  int r = -1;
  for (int k = 0; k < sp->n_species; k++) {
    if (sp->species_Z[k] == 0) return 0;
    if (sp->species_Z[k] == Zj) r = k;
  }
  return r;
}

// coordination_function / dcoordination_function (linearalgebra.f95:7488-7516) times the optional radial decay
// (descriptors.f95:8204-8216)
__device__ __forceinline__ void cutoff_fn(const SoapDev* sp, double r, double& f, double& df) {
  double fc, dfc;
  if (r > sp->cutoff) { fc = 0.0; dfc = 0.0; }
  else if (r > sp->cutoff - sp->ctw) {
    double s, c;
    sincos(PI_D * (r - sp->cutoff + sp->ctw) / sp->ctw, &s, &c);
    fc = 0.5 * (c + 1.0);
    dfc = -0.5 * PI_D * s / sp->ctw;
  } else { fc = 1.0; dfc = 0.0; }
  if (sp->cutoff_dexp > 0) {
    double rp = pow(r / sp->cutoff_scale, (double)sp->cutoff_dexp);
    double rd = sp->norm_radial_decay * (1.0 + sp->cutoff_rate) / (sp->cutoff_rate + rp);
    double drd = -sp->norm_radial_decay * sp->cutoff_dexp * (1.0 + sp->cutoff_rate) * rp / (r * (sp->cutoff_rate + rp) * (sp->cutoff_rate + rp));
    df = dfc * rd + fc * drd;
    f = fc * rd;
  } else { f = fc; df = dfc; }
}

// Radial item (neighbour q, basis point a): Phi_l(a) = f phi_l(a) (and R_l(a) = f phi_l'(a) + f' phi_l(a)) for l = 0..L,
// written at out[l * ld] (descriptors.f95:8218-8258 -- the upward recursion exactly as the reference runs it).
template <bool GRAD>
__device__ __forceinline__ void radial_item(double alpha, double r, double rb, double f, double df, int L, double* out, double* dout, int ld) {
  double arg = 2.0 * alpha * r * rb;
  if (arg == 0.0) {
    double bl = exp(-alpha * (rb * rb + r * r));
    out[0] = f * bl;
    if (GRAD) dout[0] = f * (-2.0 * alpha * r * bl) + df * bl;
    for (int l = 1; l <= L; l++) {
      out[l * ld] = 0.0;
      if (GRAD) dout[l * ld] = 0.0;
    }
    return;
  }
  double exp_p = exp(-alpha * (r + rb) * (r + rb));
  double exp_m = exp(-alpha * (r - rb) * (r - rb));
  double inv = 1.0 / arg, rinv = 1.0 / r;
  double blm = 0.5 * (exp_m + exp_p) * inv;
  double bl = 0.5 * (exp_m - exp_p) * inv;
  double blp = blm - bl * inv;
  out[0] = f * bl;
  if (GRAD) dout[0] = f * (-2.0 * alpha * r * bl + blp * 2.0 * alpha * rb) + df * bl;
  for (int l = 1; l <= L; l++) {
    blm = bl;
    bl = blp;
    blp = blm - (double)(2 * l + 1) * bl * inv;
    out[l * ld] = f * bl;
    if (GRAD) dout[l * ld] = f * (-2.0 * alpha * r * bl + (double)l * bl * rinv + blp * 2.0 * alpha * rb) + df * bl;
  }
}

// (x + i y)^m by repeated multiplication; also returns the (m-1)th power (needed by the gradient)
__device__ __forceinline__ void cs_power(double ux, double uy, int m, double& Cm, double& Sm, double& Cm1, double& Sm1) {
  Cm = 1.0; Sm = 0.0; Cm1 = 0.0; Sm1 = 0.0;
  for (int k = 0; k < m; k++) {
    Cm1 = Cm; Sm1 = Sm;
    Cm = ux * Cm1 - uy * Sm1;
    Sm = ux * Sm1 + uy * Cm1;
  }
}

// Harmonic item (neighbour q, order m): real orthonormal Y_{l,+m} (cos type) and Y_{l,-m} (sin type), l = m..L, index
// lm = l*l + l +- m.  Y_lm = N_lm Q_l^m(z) {C_m, S_m}(x, y) with Q_l^m = d^m P_l / dz^m (upward recursion in l) and
// C_m + i S_m = (x + i y)^m; N_lm carries sqrt(2) for m > 0.  (The reference uses complex Y_lm,
// angular_functions.f95:120-136; the power spectrum is invariant under this unitary change of basis.)
// GRAD: also the gradient of the polynomial extension N_lm Q_l^m(z) {C_m,S_m}(x,y) (three tables, stride gstride); the
// consumer projects it: grad Y = (g - u (u.g)) / r  (GradSphericalYCartesian_all, angular_functions.f95:205-278).
template <bool GRAD>
__device__ __forceinline__ void ylm_order(const double* __restrict__ ynorm, int L, int m, double ux, double uy, double uz, double* Yq, double* Gq,
                                          int gstride) {
  double Cm, Sm, Cm1, Sm1;
  cs_power(ux, uy, m, Cm, Sm, Cm1, Sm1);
  const double dm = (double)m;
  double p2 = 0.0, p1 = c_dblfact[m];  // Q^m recursion, starts at Q_m^m = (2m-1)!!
  double z2 = 0.0, z1 = 0.0;           // Q^{m+1} recursion (= dQ^m/dz), zero at l = m
  for (int l = m; l <= L; l++) {
    double pl, zl = 0.0;
    if (l == m) pl = p1;
    else {
      pl = ((double)(2 * l - 1) * uz * p1 - (double)(l + m - 1) * p2) * c_invint[l - m];
      p2 = p1; p1 = pl;
      if (GRAD) {
        if (l == m + 1) zl = c_dblfact[m + 1];
        else zl = ((double)(2 * l - 1) * uz * z1 - (double)(l + m) * z2) * c_invint[l - m - 1];
        z2 = z1; z1 = zl;
      }
    }
    const double nrm = ynorm[l * (l + 1) / 2 + m];
    const double q = pl * nrm, qz = zl * nrm;
    const int base = l * l + l;
    if (m == 0) {
      Yq[base] = q;
      if (GRAD) { Gq[base] = 0.0; Gq[gstride + base] = 0.0; Gq[2 * gstride + base] = qz; }
    } else {
      Yq[base + m] = q * Cm;
      Yq[base - m] = q * Sm;
      if (GRAD) {
        const double qm = q * dm;
        Gq[base + m] = qm * Cm1; Gq[gstride + base + m] = -qm * Sm1; Gq[2 * gstride + base + m] = qz * Cm;
        Gq[base - m] = qm * Sm1; Gq[gstride + base - m] = qm * Cm1;  Gq[2 * gstride + base - m] = qz * Sm;
      }
    }
  }
}

template <bool GRAD>
__device__ __forceinline__ void ylm_item(const double* __restrict__ ynorm, int L, int item, double ux, double uy, double uz, double* Yq, double* Gq,
                                         int gstride) {
  ylm_order<GRAD>(ynorm, L, item, ux, uy, uz, Yq, Gq, gstride);
  const int m2 = L + 1 - item;
  if (item > 0 && m2 > item) ylm_order<GRAD>(ynorm, L, m2, ux, uy, uz, Yq, Gq, gstride);
}

}  // namespace

}  // namespace gapb200
