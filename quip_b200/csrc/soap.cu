// soap.cu -- SOAP power-spectrum descriptor (forward) and its reverse-mode gradient + force/virial scatter.
//
// Replaces soap_calc's atomic path (src/GAP/descriptors.f95:8100-8611), SphericalYCartesian_all /
// GradSphericalYCartesian_all (src/libAtoms/angular_functions.f95:120-136, 205-278), the cutoff function
// (src/libAtoms/linearalgebra.f95:7488-7516) and the SOAP part of the scatter loop of IPModel_GAP_Calc
// (src/Potentials/IPModel_GAP.f95:472-499).
//
// Not a port.  Two deliberate re-designs, both algebraically identical to the reference:
//  * REAL spherical harmonics (2l+1 reals per l) instead of complex ones stored as separate real and imaginary
//    (2l+1) arrays: the power spectrum sum_m conj(X_lm(a)) X_lm(b) is invariant under the unitary change of basis,
//    and half of the reference's storage/work is redundant (Y_l,-m = (-1)^m conj(Y_lm)).
//  * REVERSE-mode gradients: the reference materialises grad_data(d,3,0:nn) per centre (226 KB/atom at d=325,
//    forward mode, descriptors.f95:8462-8608) and contracts it with gradPredict afterwards
//    (IPModel_GAP.f95:479).  Here dE_i/dx is pulled back through normalisation and power spectrum once per
//    centre (Lambda = dE_i/dX_lm, a nlm x K1 block in shared memory) and each neighbour costs
//    O(n_max (l_max+1)^2) to turn Lambda into the 3-vector f_gp: grad_data never exists.
//
// One CTA (128 threads) per centre.  Neighbour shells are staged in shared memory in tiles; warp 0 evaluates the
// harmonics (one neighbour per lane) while warps 1-3 evaluate the radial functions; the contraction over the
// tile is done by all threads with register accumulation; all reductions are fixed-order (deterministic) except
// the final force scatter, which uses FP64 atomics on HBM.
#include "gap_device.cuh"

namespace gapb200 {

namespace {

constexpr int NT = 128;    // threads per CTA
constexpr int NBCAP = 128; // CSR entries examined per pass (= compacted list capacity)
constexpr int TNF = 32;    // neighbours per tile, forward
constexpr int TNA = 16;    // neighbours per tile, adjoint
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

// Radial functions of one (neighbour, basis point) item for l = 0..l_max (descriptors.f95:8218-8258):
// out[l*stride] = exp(-alpha (r^2 + r_a^2)) i_l(2 alpha r r_a), upward recursion exactly as the reference,
// dout[l*stride] = d/dr of it.
template <bool GRAD>
__device__ __forceinline__ void radial_item(double alpha, double r, double rb, int l_max, double* out, double* dout, int stride) {
  double arg = 2.0 * alpha * r * rb;
  if (arg == 0.0) {
    double bl = exp(-alpha * (rb * rb + r * r));
    out[0] = bl;
    if (GRAD) dout[0] = -2.0 * alpha * r * bl;
    for (int l = 1; l <= l_max; l++) {
      out[l * stride] = 0.0;
      if (GRAD) dout[l * stride] = 0.0;
    }
    return;
  }
  double exp_p = exp(-alpha * (r + rb) * (r + rb));
  double exp_m = exp(-alpha * (r - rb) * (r - rb));
  double inv = 1.0 / arg, rinv = 1.0 / r;
  double blm = 0.5 * (exp_m + exp_p) * inv;
  double bl = 0.5 * (exp_m - exp_p) * inv;
  double blp = blm - bl * inv;
  out[0] = bl;
  if (GRAD) dout[0] = -2.0 * alpha * r * bl + blp * 2.0 * alpha * rb;
  for (int l = 1; l <= l_max; l++) {
    blm = bl;
    bl = blp;
    blp = blm - (double)(2 * l + 1) * bl * inv;
    out[l * stride] = bl;
    if (GRAD) dout[l * stride] = -2.0 * alpha * r * bl + (double)l * bl * rinv + blp * 2.0 * alpha * rb;
  }
}

// Real orthonormal spherical harmonics of the unit vector u for all l <= l_max, index lm = l*l + l + m
// (m > 0: cos-type, m < 0: sin-type), and -- if GRAD -- their gradient with respect to the UNNORMALISED
// displacement (u = dvec/r): grad = (g - u (u.g)) / r with g the gradient of the polynomial extension
// N_lm Q_l^m(z) {C_m,S_m}(x,y), Q_l^m = d^m P_l/dz^m, C_m + i S_m = (x + i y)^m.
template <bool GRAD>
__device__ __forceinline__ void real_ylm(const double* __restrict__ ynorm, int L, double ux, double uy, double uz, double rinv, double* Y,
                                         double* dY, int comp_stride) {
  double Cm[LC + 1], Sm[LC + 1], Qn[LC + 2], Qc[LC + 2];
  Cm[0] = 1.0;
  Sm[0] = 0.0;
  for (int m = 1; m <= L; m++) {
    Cm[m] = ux * Cm[m - 1] - uy * Sm[m - 1];
    Sm[m] = ux * Sm[m - 1] + uy * Cm[m - 1];
  }
  for (int l = 0; l <= L + 1; l++) Qn[l] = 0.0;
  for (int m = L; m >= 0; m--) {
    Qc[m] = c_dblfact[m];
    if (m + 1 <= L) Qc[m + 1] = (double)(2 * m + 1) * uz * Qc[m];
    for (int l = m + 2; l <= L; l++) Qc[l] = ((double)(2 * l - 1) * uz * Qc[l - 1] - (double)(l + m - 1) * Qc[l - 2]) * c_invint[l - m];
    for (int l = m; l <= L; l++) {
      double nrm = ynorm[l * (l + 1) / 2 + m];
      double q = Qc[l] * nrm;
      int base = l * l + l;
      if (m == 0) {
        Y[base] = q;
        if (GRAD) {
          double gz = Qn[l] * nrm;
          double dot = uz * gz;
          dY[base] = (-ux * dot) * rinv;
          dY[comp_stride + base] = (-uy * dot) * rinv;
          dY[2 * comp_stride + base] = (gz - uz * dot) * rinv;
        }
      } else {
        Y[base + m] = q * Cm[m];
        Y[base - m] = q * Sm[m];
        if (GRAD) {
          double qz = Qn[l] * nrm, qm = q * (double)m;
          {
            double gx = qm * Cm[m - 1], gy = -qm * Sm[m - 1], gz = qz * Cm[m];
            double dot = ux * gx + uy * gy + uz * gz;
            dY[base + m] = (gx - ux * dot) * rinv;
            dY[comp_stride + base + m] = (gy - uy * dot) * rinv;
            dY[2 * comp_stride + base + m] = (gz - uz * dot) * rinv;
          }
          {
            double gx = qm * Sm[m - 1], gy = qm * Cm[m - 1], gz = qz * Sm[m];
            double dot = ux * gx + uy * gy + uz * gz;
            dY[base - m] = (gx - ux * dot) * rinv;
            dY[comp_stride + base - m] = (gy - uy * dot) * rinv;
            dY[2 * comp_stride + base - m] = (gz - uz * dot) * rinv;
          }
        }
      }
    }
    for (int l = m; l <= L; l++) Qn[l] = Qc[l];
    if (m >= 1) Qn[m - 1] = 0.0;
  }
}

struct Smem {
  double* T;      // n_max*n_max
  double* rb;     // n_max
  double* ynorm;  // (L+1)(L+2)/2
  double* X;      // nlm*K1   (forward: accumulates X ; adjoint: Lambda)
  double* nbd;    // NBCAP*3 displacement
  double* nbr;    // NBCAP distance
  double* red;    // 64
  double* rf;     // TN*(L+1)*n_max
  double* drf;    // adjoint only
  double* Y;      // TN*nlm
  double* dY;     // adjoint only: 3 * TN*nlm
  double* p;      // d_pad  (aliases rf.. region in forward; own region in adjoint)
  int* nbs;       // NBCAP species
  int* nbj;       // NBCAP neighbour atom
  int* lof;       // nlm: l of lm
  int* wcount;    // 8
};

__host__ __device__ inline size_t carve(const SoapDev& h, bool adjoint, Smem* s, unsigned char* base) {
  const int n = h.n_max, L1 = h.l_max + 1, nlm = h.nlm, TN = adjoint ? TNA : TNF;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o += cnt * sizeof(double); return r; };
  size_t oT = take(n * n), orb = take(n), oyn = take((size_t)L1 * (L1 + 1) / 2), oX = take((size_t)nlm * h.K1), onbd = take(NBCAP * 3),
         onbr = take(NBCAP), ored = take(64);
  size_t orf = take((size_t)TN * L1 * n), odrf = adjoint ? take((size_t)TN * L1 * n) : 0, oY = take((size_t)TN * nlm),
         odY = adjoint ? take((size_t)3 * TN * nlm) : 0;
  size_t stage_bytes = o - orf;
  size_t op;
  if (adjoint) {
    // the staging area doubles as scratch for X_lm while Lambda is formed
    if ((size_t)nlm * h.K1 * sizeof(double) > stage_bytes) o = orf + (size_t)nlm * h.K1 * sizeof(double);
    op = take(h.d_pad);
  } else {
    op = orf;  // forward: the power spectrum reuses the staging area after the neighbour loop
    if ((size_t)h.d_pad * sizeof(double) > stage_bytes) o = orf + (size_t)h.d_pad * sizeof(double);
  }
  size_t oi = o;
  o += sizeof(int) * (NBCAP * 2 + nlm + 8);
  if (s) {
    s->T = (double*)(base + oT); s->rb = (double*)(base + orb); s->ynorm = (double*)(base + oyn); s->X = (double*)(base + oX);
    s->nbd = (double*)(base + onbd); s->nbr = (double*)(base + onbr); s->red = (double*)(base + ored); s->rf = (double*)(base + orf);
    s->drf = (double*)(base + odrf); s->Y = (double*)(base + oY); s->dY = (double*)(base + odY); s->p = (double*)(base + op);
    s->nbs = (int*)(base + oi); s->nbj = s->nbs + NBCAP; s->lof = s->nbj + NBCAP; s->wcount = s->lof + nlm;
  }
  return o;
}

__device__ __forceinline__ void load_tables(const SoapDev* sp, const Smem& s) {
  const int n = sp->n_max, L1 = sp->l_max + 1;
  for (int k = threadIdx.x; k < n * n; k += NT) s.T[k] = sp->T[k];
  for (int k = threadIdx.x; k < n; k += NT) s.rb[k] = sp->r_basis[k];
  for (int k = threadIdx.x; k < L1 * (L1 + 1) / 2; k += NT) s.ynorm[k] = sp->ynorm[k];
  for (int k = threadIdx.x; k < sp->nlm; k += NT) {
    int l = 0;
    while ((l + 1) * (l + 1) <= k) l++;
    s.lof[k] = l;
  }
}

// Ordered (deterministic) block compaction of up to NBCAP CSR entries of centre i into shared memory.
__device__ __forceinline__ int gather_neighbours(const SoapDev* sp, const Smem& s, int i, int pbeg, int pend, const int* __restrict__ nbr_j,
                                                 const int* __restrict__ nbr_s, const double* __restrict__ pos, const int* __restrict__ Z,
                                                 const Lattice9& lat) {
  int p = pbeg + threadIdx.x;
  bool valid = false;
  double dd[3], r = 0.0;
  int spc = -1, j = -1;
  if (p < pend) {
    j = nbr_j[p];
    int s0, s1, s2;
    unpack_shift(nbr_s[p], s0, s1, s2);
    image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)j, lat.v, s0, s1, s2, dd);
    r = norm_nofma(dd);
    spc = species_of(sp, Z[j]);
    valid = (r < sp->cutoff) && (spc >= 0);  // descriptors.f95:8190, 8194-8195
  }
  unsigned bal = __ballot_sync(0xffffffffu, valid);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s.wcount[w] = __popc(bal);
  __syncthreads();
  int off = 0;
  for (int k = 0; k < w; k++) off += s.wcount[k];
  int total = s.wcount[0] + s.wcount[1] + s.wcount[2] + s.wcount[3];
  if (valid) {
    int q = off + __popc(bal & ((1u << lane) - 1u));
    s.nbd[3 * q] = dd[0];
    s.nbd[3 * q + 1] = dd[1];
    s.nbd[3 * q + 2] = dd[2];
    s.nbr[q] = r;
    s.nbs[q] = spc;
    s.nbj[q] = j;
  }
  __syncthreads();
  return total;
}

// stage one tile: warp 0 -> harmonics, warps 1..3 -> radial functions; then the radial transform in place.
template <bool GRAD>
__device__ __forceinline__ void stage_tile(const SoapDev* sp, const Smem& s, int t0, int tn) {
  const int n = sp->n_max, L = sp->l_max, L1 = L + 1, nlm = sp->nlm, TN = GRAD ? TNA : TNF;
  if (threadIdx.x < 32) {
    int q = threadIdx.x;
    if (q < tn) {
      double r = s.nbr[t0 + q], rinv = 1.0 / r;
      real_ylm<GRAD>(s.ynorm, L, s.nbd[3 * (t0 + q)] * rinv, s.nbd[3 * (t0 + q) + 1] * rinv, s.nbd[3 * (t0 + q) + 2] * rinv, rinv,
                     s.Y + (size_t)q * nlm, GRAD ? s.dY + (size_t)q * nlm : nullptr, TN * nlm);
    }
  } else {
    for (int it = threadIdx.x - 32; it < tn * n; it += NT - 32) {
      int q = it / n, a = it - q * n;
      radial_item<GRAD>(sp->alpha, s.nbr[t0 + q], s.rb[a], L, s.rf + (size_t)q * L1 * n + a, GRAD ? s.drf + (size_t)q * L1 * n + a : nullptr, n);
    }
  }
  __syncthreads();
  // radial_coefficient = matmul(radial_fun, transform_basis) * f_cut  (descriptors.f95:8261-8263), row (q,l) in place
  for (int row = threadIdx.x; row < tn * L1; row += NT) {
    int q = row / L1;
    double f, df;
    cutoff_fn(sp, s.nbr[t0 + q], f, df);
    double v[SOAP_NMAX_CAP], dv[SOAP_NMAX_CAP];
    double* rr = s.rf + (size_t)row * n;
    double* dr = GRAD ? s.drf + (size_t)row * n : nullptr;
    for (int a = 0; a < n; a++) {
      v[a] = rr[a];
      if (GRAD) dv[a] = dr[a];
    }
    for (int b = 0; b < n; b++) {
      double t = 0.0, tg = 0.0;
      for (int a = 0; a < n; a++) {
        t += v[a] * s.T[a + n * b];
        if (GRAD) tg += dv[a] * s.T[a + n * b];
      }
      rr[b] = t * f;
      if (GRAD) dr[b] = tg * f + t * df;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward: x (normalised power spectrum), X_lm (kept for the adjoint), |p|
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_soap_forward(const SoapDev* __restrict__ sp, const int* __restrict__ centres, int n_centres,
                                                     const int* __restrict__ nbr_off, const int* __restrict__ nbr_j,
                                                     const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                     const int* __restrict__ Z, Lattice9 lat, double* __restrict__ x,
                                                     double* __restrict__ xlm, double* __restrict__ pnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem s;
  carve(*sp, false, &s, smem_raw);
  const int c = blockIdx.x;
  if (c >= n_centres) return;
  const int i = centres[c];
  const int n = sp->n_max, L1 = sp->l_max + 1, nlm = sp->nlm, K1 = sp->K1, d = sp->d, d_pad = sp->d_pad;
  load_tables(sp, s);
  for (int k = threadIdx.x; k < nlm * K1; k += NT) s.X[k] = 0.0;
  __syncthreads();

  const int pbeg = nbr_off[i], pend = nbr_off[i + 1];
  for (int pb = pbeg; pb < pend; pb += NBCAP) {
    int nv = gather_neighbours(sp, s, i, pb, min(pb + NBCAP, pend), nbr_j, nbr_s, pos, Z, lat);
    for (int t0 = 0; t0 < nv; t0 += TNF) {
      int tn = min(TNF, nv - t0);
      stage_tile<false>(sp, s, t0, tn);
      // X_lm(species, a) += sum_q c_l(a; q) Y_lm(q)   (descriptors.f95:8289-8295)
      for (int idx = threadIdx.x; idx < nlm * n; idx += NT) {
        int lm = idx / n, a = idx - lm * n, l = s.lof[lm];
        double acc = 0.0;
        int cur = s.nbs[t0];
        for (int q = 0; q < tn; q++) {
          int spq = s.nbs[t0 + q];
          if (spq != cur) {
            s.X[lm * K1 + cur * n + a] += acc;
            acc = 0.0;
            cur = spq;
          }
          acc += s.Y[(size_t)q * nlm + lm] * s.rf[((size_t)q * L1 + l) * n + a];
        }
        s.X[lm * K1 + cur * n + a] += acc;
      }
      __syncthreads();
    }
  }
  // central atom term (descriptors.f95:8151-8182): only a = 1 is non-zero because the Cholesky factor is lower triangular
  if (threadIdx.x < sp->n_species) {
    int k = threadIdx.x;
    if (sp->cras || sp->species_Z[k] == Z[i] || sp->species_Z[k] == 0) s.X[0 * K1 + k * n + 0] += sp->central_weight * sp->chol00 * 0.28209479177387814347;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nlm * K1; k += NT) xlm[(size_t)c * nlm * K1 + k] = s.X[k];

  // power spectrum (descriptors.f95:8370-8418): element q = l + (l_max+1) * pair(ia, jb<=ia)
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += NT) {
    int l = q % L1, pr = q / L1;
    int ia = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
    while ((ia + 1) * (ia + 2) / 2 <= pr) ia++;
    while (ia * (ia + 1) / 2 > pr) ia--;
    int jb = pr - ia * (ia + 1) / 2;
    double t = 0.0;
    for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) t += s.X[lm * K1 + ia] * s.X[lm * K1 + jb];
    t *= sp->tlpo[l];
    if (ia != jb) t *= 1.41421356237309504880;
    s.p[q] = t;
    loc += t * t;
  }
  double nrm = sqrt(block_sum(loc, s.red));  // :8450-8451
  double inv = sp->normalise ? 1.0 / nrm : 1.0;
  double* xr = x + (size_t)c * d_pad;
  for (int q = threadIdx.x; q < d_pad; q += NT) xr[q] = q < d - 1 ? s.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (threadIdx.x == 0) pnorm[c] = nrm;
}

// ------------------------------------------------------------------------------------------------
// adjoint: gvec = dE_i/dx  ->  forces / virial
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_soap_adjoint(const SoapDev* __restrict__ sp, const int* __restrict__ centres, int n_centres,
                                                     const int* __restrict__ nbr_off, const int* __restrict__ nbr_j,
                                                     const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                     const int* __restrict__ Z, Lattice9 lat, const double* __restrict__ x,
                                                     const double* __restrict__ xlm, const double* __restrict__ pnorm,
                                                     const double* __restrict__ gvec, int ldg, double e_scale, double* __restrict__ force,
                                                     double* __restrict__ vir_part, double* __restrict__ local_virial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem s;
  carve(*sp, true, &s, smem_raw);
  const int c = blockIdx.x;
  if (c >= n_centres) return;
  const int i = centres[c];
  const int n = sp->n_max, L1 = sp->l_max + 1, nlm = sp->nlm, K1 = sp->K1, d = sp->d, d_pad = sp->d_pad;
  load_tables(sp, s);
  // u = dE/dp: pull gradPredict back through x = p/|p| (reference forward form: descriptors.f95:8595-8600)
  const double* xr = x + (size_t)c * d_pad;
  const double* gr = gvec + (size_t)c * ldg;
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += NT) loc += xr[q] * gr[q];
  double sdot = block_sum(loc, s.red);
  double nrm = pnorm[c];
  for (int q = threadIdx.x; q < d - 1; q += NT) s.p[q] = sp->normalise ? (gr[q] - xr[q] * sdot) / nrm : gr[q];
  // stage X in the (not yet used) rf area, then Lambda = dE/dX_lm into s.X
  double* Xs = s.rf;  // nlm*K1 doubles fit: checked on the host
  for (int k = threadIdx.x; k < nlm * K1; k += NT) Xs[k] = xlm[(size_t)c * nlm * K1 + k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < nlm * K1; idx += NT) {
    int lm = idx / K1, ia = idx - lm * K1, l = s.lof[lm];
    double t = 0.0;
    for (int jb = 0; jb < K1; jb++) {
      int hi = ia > jb ? ia : jb, lo = ia > jb ? jb : ia;
      double u = s.p[l + L1 * (hi * (hi + 1) / 2 + lo)];
      t += (ia == jb ? 2.0 * u : 1.41421356237309504880 * u) * Xs[lm * K1 + jb];
    }
    s.X[idx] = t * sp->tlpo[l];
  }
  __syncthreads();

  double fi[3] = {0, 0, 0}, vir[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pbeg = nbr_off[i], pend = nbr_off[i + 1];
  for (int pb = pbeg; pb < pend; pb += NBCAP) {
    int nv = gather_neighbours(sp, s, i, pb, min(pb + NBCAP, pend), nbr_j, nbr_s, pos, Z, lat);
    for (int t0 = 0; t0 < nv; t0 += TNA) {
      int tn = min(TNA, nv - t0);
      stage_tile<true>(sp, s, t0, tn);
      // one warp per neighbour: f_gp,k = sum_{l,a} [ c'_l(a) u_k (sum_m Lambda_lm Y_lm) + c_l(a) (sum_m Lambda_lm dY_lm,k) ]
      for (int q = w; q < tn; q += NT / 32) {
        const double r = s.nbr[t0 + q], rinv = 1.0 / r;
        const double ux = s.nbd[3 * (t0 + q)] * rinv, uy = s.nbd[3 * (t0 + q) + 1] * rinv, uz = s.nbd[3 * (t0 + q) + 2] * rinv;
        const int spq = s.nbs[t0 + q];
        const double* Yq = s.Y + (size_t)q * nlm;
        const double* dYq = s.dY + (size_t)q * nlm;
        const int cs = TNA * nlm;
        double f0 = 0, f1 = 0, f2 = 0;
        for (int it = lane; it < L1 * n; it += 32) {
          int l = it / n, a = it - l * n;
          double B = 0, D0 = 0, D1 = 0, D2 = 0;
          for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) {
            double lam = s.X[lm * K1 + spq * n + a];
            B += lam * Yq[lm];
            D0 += lam * dYq[lm];
            D1 += lam * dYq[cs + lm];
            D2 += lam * dYq[2 * cs + lm];
          }
          double cl = s.rf[((size_t)q * L1 + l) * n + a], dcl = s.drf[((size_t)q * L1 + l) * n + a];
          f0 += dcl * ux * B + cl * D0;
          f1 += dcl * uy * B + cl * D1;
          f2 += dcl * uz * B + cl * D2;
        }
        f0 = warp_sum(f0) * e_scale;
        f1 = warp_sum(f1) * e_scale;
        f2 = warp_sum(f2) * e_scale;
        if (lane == 0) {
          // IPModel_GAP.f95:479-491: F_j -= f_gp ; centre row is minus the sum ; W_j -= (pos_j - pos_i) (x) f_gp
          int j = s.nbj[t0 + q];
          if (force) {
            atomicAdd(&force[3 * (size_t)j + 0], -f0);
            atomicAdd(&force[3 * (size_t)j + 1], -f1);
            atomicAdd(&force[3 * (size_t)j + 2], -f2);
            fi[0] += f0; fi[1] += f1; fi[2] += f2;
          }
          const double dx = s.nbd[3 * (t0 + q)], dy = s.nbd[3 * (t0 + q) + 1], dz = s.nbd[3 * (t0 + q) + 2];
          double wv[9] = {dx * f0, dy * f0, dz * f0, dx * f1, dy * f1, dz * f1, dx * f2, dy * f2, dz * f2};  // column-major (a + 3b)
#pragma unroll
          for (int k = 0; k < 9; k++) vir[k] -= wv[k];
          if (local_virial)
#pragma unroll
            for (int k = 0; k < 9; k++) atomicAdd(&local_virial[9 * (size_t)j + k], -wv[k]);
        }
      }
      __syncthreads();
    }
  }
  // combine the 4 warp leaders in fixed order
  __syncthreads();
  if (lane == 0) {
    for (int k = 0; k < 3; k++) s.red[w * 12 + k] = fi[k];
    for (int k = 0; k < 9; k++) s.red[w * 12 + 3 + k] = vir[k];
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    int k = threadIdx.x;
    double t = (s.red[k] + s.red[12 + k]) + (s.red[24 + k] + s.red[36 + k]);
    if (k < 3) {
      if (force) atomicAdd(&force[3 * (size_t)i + k], t);
    } else if (vir_part) vir_part[9 * (size_t)c + (k - 3)] = t;
  }
}

__global__ void k_select_centres(const int* __restrict__ Z, int first, int last, const SoapDev* __restrict__ sp, int* __restrict__ flags) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > last - first) return;
  int f = 0;
  if (t < last - first) {
    int Zi = Z[first + t];
    for (int k = 0; k < sp->n_Z; k++)
      if (sp->centre_Z[k] == Zi || sp->centre_Z[k] == 0) f = 1;  // descriptors.f95:7962
  }
  flags[t] = f;
}
__global__ void k_compact(const int* __restrict__ scan, const int* __restrict__ flags, int first, int n, int* __restrict__ centres) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && flags[t]) centres[scan[t]] = first + t;
}

}  // namespace

size_t soap_forward_smem(const SoapDev& h) { return carve(h, false, nullptr, nullptr); }
size_t soap_adjoint_smem(const SoapDev& h) { return carve(h, true, nullptr, nullptr); }

void launch_select_centres(const int* Z, int first, int last, const SoapDev* sp, int* flags, cudaStream_t st, int* launches) {
  int n = last - first + 1;
  k_select_centres<<<(n + 255) / 256, 256, 0, st>>>(Z, first, last, sp, flags);
  *launches += 1;
}
void launch_compact(const int* flags_scan, const int* flags, int first, int n, int* centres, cudaStream_t st, int* launches) {
  if (n <= 0) return;
  k_compact<<<(n + 255) / 256, 256, 0, st>>>(flags_scan, flags, first, n, centres);
  *launches += 1;
}

void launch_soap_forward(const SoapDev* sp, const SoapDev& h, const int* centres, int n_centres, const int* nbr_off, const int* nbr_j,
                         const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, double* x, double* xlm, double* pnorm,
                         cudaStream_t st, int* launches) {
  if (n_centres <= 0) return;
  size_t sm = soap_forward_smem(h);
  cudaFuncSetAttribute(k_soap_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_forward<<<n_centres, NT, sm, st>>>(sp, centres, n_centres, nbr_off, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm);
  *launches += 1;
}

void launch_soap_adjoint(const SoapDev* sp, const SoapDev& h, const int* centres, int n_centres, const int* nbr_off, const int* nbr_j,
                         const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, const double* x, const double* xlm,
                         const double* pnorm, const double* gvec, int ldg, double e_scale, double* force, double* vir_part,
                         double* local_virial, cudaStream_t st, int* launches) {
  if (n_centres <= 0) return;
  size_t sm = soap_adjoint_smem(h);
  cudaFuncSetAttribute(k_soap_adjoint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_adjoint<<<n_centres, NT, sm, st>>>(sp, centres, n_centres, nbr_off, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg, e_scale,
                                            force, vir_part, local_virial);
  *launches += 1;
}

}  // namespace gapb200
