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
constexpr int TNA = 32;    // neighbours per tile, adjoint
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

// ------------------------------------------------------------------------------------------------------------
// Work decomposition (one CTA of NT threads per centre, several CTAs resident per SM):
//   gather   : the centre's CSR row -> shared memory (displacement, distance, species, cutoff function), compacted
//              in list order (deterministic);
//   tile loop over TN neighbours at a time:
//     stage  : one item per (neighbour, radial basis point a): Phi_l(a) = f_cut * phi_l(a) for all l (and, for the
//              adjoint, R_l(a) = f_cut * phi_l'(a) + f_cut' * phi_l(a));  forward only: one item per (neighbour, m):
//              Y_{l,+-m} for all l >= m;
//     forward accumulate: one item per (lm, group of 4 basis points): Xt_lm(s,a) += sum_q Phi_l(a;q) Y_lm(q) held in
//              registers across the tile (the item owns its shared-memory row: no atomics, no extra barrier);
//     adjoint contract : one item per (neighbour, m-pair): harmonics and their gradients are generated on the fly
//              and contracted with Lambda~ and the radial tables straight away -- Y and grad Y never touch memory;
//   The transform_basis product is hoisted out of the neighbour loop: X = Xt . T once per centre (forward) and
//   Lambda~ = T . Lambda once per centre (adjoint), instead of the reference's per-neighbour matmul
//   (descriptors.f95:8261-8263): (l_max+1) n_max^2 flops per NEIGHBOUR become (l_max+1)^2 n_max^2 per CENTRE.
// ------------------------------------------------------------------------------------------------------------
constexpr int AG = 4;  // radial basis points per accumulate item

__host__ __device__ inline int ceil4(int v) { return (v + 3) & ~3; }
// row stride (doubles) of the per-neighbour radial tables: = 2 (mod 16), so that 8 consecutive neighbours read
// 16-byte words from 8 distinct bank quads
__host__ __device__ inline int rf_stride(int L1, int n4) {
  int s = L1 * n4;
  while ((s & 15) != 2) s += 2;
  return s;
}
__host__ __device__ inline int y_stride(int nlm) { return nlm | 1; }

struct Smem {
  double* T;       // n*n transform_basis, T[a + n*a']
  double* rb;      // n
  double* ynorm;   // (L+1)(L+2)/2
  double* X;       // nlm * K1p : forward Xt -> X ; adjoint Lambda~   (row lm, column s*n4 + a)
  double* nbd;     // NBCAP*3 displacement
  double* nbr;     // NBCAP distance
  double* nbf;     // NBCAP f_cut
  double* nbdf;    // NBCAP f_cut'
  double* red;     // 64
  double* rf;      // TN * RFS
  double* drf;     // adjoint: TN * RFS
  double* Y;       // forward: TN * YS
  double* part;    // adjoint: TN * MP * 3 partial forces
  double* p;       // forward: d_pad power spectrum (aliases the staging area) ; adjoint: d_pad, own region
  int* nbs;        // NBCAP species
  int* nbj;        // NBCAP neighbour atom
  int* lof;        // nlm: l of lm
  int* wcount;     // 8
};

__host__ __device__ inline int m_pairs(int L) { return L / 2 + 1 + (L & 1); }  // items per neighbour in the adjoint contraction

__host__ __device__ inline size_t carve(const SoapDev& h, bool adjoint, Smem* s, unsigned char* base) {
  const int n = h.n_max, L1 = h.l_max + 1, nlm = h.nlm, TN = adjoint ? TNA : TNF, n4 = ceil4(n), K1p = h.n_species * n4;
  const int RFS = rf_stride(L1, n4), YS = y_stride(nlm), MP = m_pairs(h.l_max);
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o += ((cnt + 1) & ~(size_t)1) * sizeof(double); return r; };
  size_t oT = take(n * n), orb = take(n), oyn = take((size_t)L1 * (L1 + 1) / 2), oX = take((size_t)nlm * K1p), onbd = take(NBCAP * 3),
         onbr = take(NBCAP), onbf = take(NBCAP), onbdf = take(NBCAP), ored = take(64);
  size_t orf = take((size_t)TN * RFS), odrf = adjoint ? take((size_t)TN * RFS) : 0, oY = adjoint ? 0 : take((size_t)TN * YS),
         opart = adjoint ? take((size_t)TN * MP * 3) : 0;
  size_t stage_bytes = o - orf;
  size_t op;
  if (adjoint) {
    // the staging area doubles as scratch for X_lm / Lambda while Lambda~ is formed (2 * nlm * K1 doubles)
    if ((size_t)2 * nlm * h.K1 * sizeof(double) > stage_bytes) o = orf + (size_t)2 * nlm * h.K1 * sizeof(double);
    op = take(h.d_pad);
  } else {
    op = orf;  // forward: the power spectrum reuses the staging area after the neighbour loop
    if ((size_t)h.d_pad * sizeof(double) > stage_bytes) o = orf + (size_t)h.d_pad * sizeof(double);
  }
  size_t oi = o;
  o += sizeof(int) * (NBCAP * 2 + nlm + 8);
  o = (o + 15) & ~(size_t)15;
  if (s) {
    s->T = (double*)(base + oT); s->rb = (double*)(base + orb); s->ynorm = (double*)(base + oyn); s->X = (double*)(base + oX);
    s->nbd = (double*)(base + onbd); s->nbr = (double*)(base + onbr); s->nbf = (double*)(base + onbf); s->nbdf = (double*)(base + onbdf);
    s->red = (double*)(base + ored); s->rf = (double*)(base + orf); s->drf = (double*)(base + odrf); s->Y = (double*)(base + oY);
    s->part = (double*)(base + opart); s->p = (double*)(base + op);
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
    double f, df;
    cutoff_fn(sp, r, f, df);
    s.nbd[3 * q] = dd[0];
    s.nbd[3 * q + 1] = dd[1];
    s.nbd[3 * q + 2] = dd[2];
    s.nbr[q] = r;
    s.nbf[q] = f;
    s.nbdf[q] = df;
    s.nbs[q] = spc;
    s.nbj[q] = j;
  }
  __syncthreads();
  return total;
}

// Radial item (neighbour q, basis point a): Phi_l(a) = f phi_l(a) (and R_l(a) = f phi_l'(a) + f' phi_l(a)) for l = 0..L,
// written at out[l * n4] (descriptors.f95:8218-8258 -- the upward recursion exactly as the reference runs it).
template <bool GRAD>
__device__ __forceinline__ void radial_item(double alpha, double r, double rb, double f, double df, int L, double* out, double* dout, int n4) {
  double arg = 2.0 * alpha * r * rb;
  if (arg == 0.0) {
    double bl = exp(-alpha * (rb * rb + r * r));
    out[0] = f * bl;
    if (GRAD) dout[0] = f * (-2.0 * alpha * r * bl) + df * bl;
    for (int l = 1; l <= L; l++) {
      out[l * n4] = 0.0;
      if (GRAD) dout[l * n4] = 0.0;
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
    out[l * n4] = f * bl;
    if (GRAD) dout[l * n4] = f * (-2.0 * alpha * r * bl + (double)l * bl * rinv + blp * 2.0 * alpha * rb) + df * bl;
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

// Forward harmonic item (neighbour q, order m): real orthonormal Y_{l,+m} (cos type) and Y_{l,-m} (sin type), l = m..L,
// index lm = l*l + l +- m.  Y_lm = N_lm Q_l^m(z) {C_m, S_m}(x, y) with Q_l^m = d^m P_l / dz^m (upward recursion in l) and
// C_m + i S_m = (x + i y)^m; N_lm carries sqrt(2) for m > 0.  (The reference uses complex Y_lm,
// angular_functions.f95:120-136; the power spectrum is invariant under this unitary change of basis.)
__device__ __forceinline__ void ylm_item(const double* __restrict__ ynorm, int L, int m, double ux, double uy, double uz, double* Yq) {
  double Cm, Sm, Cm1, Sm1;
  cs_power(ux, uy, m, Cm, Sm, Cm1, Sm1);
  double p2 = 0.0, p1 = c_dblfact[m];  // Q_{l-2}^m, Q_{l-1}^m while stepping; starts at Q_m^m = (2m-1)!!
  for (int l = m; l <= L; l++) {
    double pl;
    if (l == m) pl = p1;
    else {
      pl = ((double)(2 * l - 1) * uz * p1 - (double)(l + m - 1) * p2) * c_invint[l - m];
      p2 = p1;
      p1 = pl;
    }
    double q = pl * ynorm[l * (l + 1) / 2 + m];
    int base = l * l + l;
    if (m == 0) Yq[base] = q;
    else {
      Yq[base + m] = q * Cm;
      Yq[base - m] = q * Sm;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward: x (normalised power spectrum), X_lm (kept for the adjoint), |p|
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 4) k_soap_forward(const SoapDev* __restrict__ sp, const int* __restrict__ centres,
                                                        const int* __restrict__ n_centres_dev,
                                                        const int* __restrict__ nbr_off, const int* __restrict__ nbr_j,
                                                        const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                        const int* __restrict__ Z, Lattice9 lat, double* __restrict__ x,
                                                        double* __restrict__ xlm, double* __restrict__ pnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem s;
  carve(*sp, false, &s, smem_raw);
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;  // the grid is sized by an upper bound; the centre count never leaves the device
  const int i = centres[c];
  const int n = sp->n_max, L = sp->l_max, L1 = L + 1, nlm = sp->nlm, K1 = sp->K1, d = sp->d, d_pad = sp->d_pad, ns = sp->n_species;
  const int n4 = ceil4(n), K1p = ns * n4, RFS = rf_stride(L1, n4), YS = y_stride(nlm), NG = n4 / AG;
  const double alpha = sp->alpha;
  load_tables(sp, s);
  for (int k = threadIdx.x; k < nlm * K1p; k += NT) s.X[k] = 0.0;
  __syncthreads();

  const int pbeg = nbr_off[i], pend = nbr_off[i + 1];
  for (int pb = pbeg; pb < pend; pb += NBCAP) {
    int nv = gather_neighbours(sp, s, i, pb, min(pb + NBCAP, pend), nbr_j, nbr_s, pos, Z, lat);
    for (int t0 = 0; t0 < nv; t0 += TNF) {
      const int tn = min(TNF, nv - t0);
      // ---- stage: radial items then harmonic items ----
      const int n_rad = tn * n, n_items = n_rad + tn * L1;
      for (int it = threadIdx.x; it < n_items; it += NT) {
        if (it < n_rad) {
          int q = it / n, a = it - q * n;
          radial_item<false>(alpha, s.nbr[t0 + q], s.rb[a], s.nbf[t0 + q], 0.0, L, s.rf + (size_t)q * RFS + a, nullptr, n4);
        } else {
          int r2 = it - n_rad;
          int m = r2 / tn, q = r2 - m * tn;
          double rinv = 1.0 / s.nbr[t0 + q];
          ylm_item(s.ynorm, L, m, s.nbd[3 * (t0 + q)] * rinv, s.nbd[3 * (t0 + q) + 1] * rinv, s.nbd[3 * (t0 + q) + 2] * rinv,
                   s.Y + (size_t)q * YS);
        }
      }
      if (n4 != n)  // zero the padding columns once per tile (read by the 4-wide accumulate)
        for (int it = threadIdx.x; it < tn * L1 * (n4 - n); it += NT) {
          int row = it / (n4 - n), a = n + it % (n4 - n);
          int q = row / L1, l = row - q * L1;
          s.rf[(size_t)q * RFS + l * n4 + a] = 0.0;
        }
      __syncthreads();
      // ---- accumulate: Xt_lm(s, a) += sum_q Phi_l(a; q) Y_lm(q)   (descriptors.f95:8289-8295, before the basis transform) ----
      for (int it = threadIdx.x; it < nlm * NG; it += NT) {
        const int lm = it / NG, g = it - lm * NG, l = s.lof[lm];
        const double* rfp = s.rf + l * n4 + g * AG;
        const double* yp = s.Y + lm;
        double* xrow = s.X + (size_t)lm * K1p + g * AG;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int cur = s.nbs[t0];
        for (int q = 0; q < tn; q++) {
          int spq = s.nbs[t0 + q];
          if (spq != cur) {
            double* xr = xrow + cur * n4;
            xr[0] += a0; xr[1] += a1; xr[2] += a2; xr[3] += a3;
            a0 = a1 = a2 = a3 = 0.0;
            cur = spq;
          }
          const double y = yp[(size_t)q * YS];
          const double2 r01 = *reinterpret_cast<const double2*>(rfp + (size_t)q * RFS);
          const double2 r23 = *reinterpret_cast<const double2*>(rfp + (size_t)q * RFS + 2);
          a0 += y * r01.x; a1 += y * r01.y; a2 += y * r23.x; a3 += y * r23.y;
        }
        double* xr = xrow + cur * n4;
        xr[0] += a0; xr[1] += a1; xr[2] += a2; xr[3] += a3;
      }
      __syncthreads();
    }
  }
  // ---- basis transform, once per centre: X_lm(s, a') = sum_a Xt_lm(s, a) T(a, a')   (in place: the item owns its row) ----
  for (int it = threadIdx.x; it < nlm * ns; it += NT) {
    double* row = s.X + (size_t)(it / ns) * K1p + (it % ns) * n4;
    double v[SOAP_NMAX_CAP];
    for (int a = 0; a < n; a++) v[a] = row[a];
    for (int b = 0; b < n; b++) {
      double t = 0.0;
      for (int a = 0; a < n; a++) t += v[a] * s.T[a + n * b];
      row[b] = t;
    }
  }
  __syncthreads();
  // central atom term (descriptors.f95:8151-8182): only a = 1 is non-zero because the Cholesky factor is lower triangular
  if (threadIdx.x < ns) {
    int k = threadIdx.x;
    if (sp->cras || sp->species_Z[k] == Z[i] || sp->species_Z[k] == 0) s.X[k * n4] += sp->central_weight * sp->chol00 * 0.28209479177387814347;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nlm * K1; k += NT) {
    int lm = k / K1, ic = k - lm * K1, sk = ic / n;
    xlm[(size_t)c * nlm * K1 + k] = s.X[(size_t)lm * K1p + sk * n4 + (ic - sk * n)];
  }

  // power spectrum (descriptors.f95:8370-8418): element q = l + (l_max+1) * pair(ia, jb<=ia)
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += NT) {
    int l = q % L1, pr = q / L1;
    int ia = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
    while ((ia + 1) * (ia + 2) / 2 <= pr) ia++;
    while (ia * (ia + 1) / 2 > pr) ia--;
    int jb = pr - ia * (ia + 1) / 2;
    const int ca = (ia / n) * n4 + ia % n, cb = (jb / n) * n4 + jb % n;
    double t = 0.0;
    for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) t += s.X[lm * K1p + ca] * s.X[lm * K1p + cb];
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
// Contribution of the harmonics of order m (both the cos and the sin type, all l >= m) of one neighbour to
//   SA = sum_lm A_lm Y_lm           with A_lm = sum_a Lambda~_lm(a) R_l(a)
//   G  = sum_lm B_lm grad_poly Y_lm with B_lm = sum_a Lambda~_lm(a) Phi_l(a)
// where grad_poly is the gradient of the polynomial extension N_lm Q_l^m(z) {C_m,S_m}(x,y); the caller projects it:
//   f_k = SA u_k + (G_k - u_k (u.G)) / r
__device__ __forceinline__ void adjoint_order(const double* __restrict__ ynorm, int L, int m, double ux, double uy, double uz,
                                              const double* __restrict__ lam /* Lambda~ + s*n4 */, int K1p, const double* __restrict__ rf,
                                              const double* __restrict__ drf, int n4, double& SA, double& G0, double& G1, double& G2) {
  double Cm, Sm, Cm1, Sm1;
  cs_power(ux, uy, m, Cm, Sm, Cm1, Sm1);
  const double dm = (double)m;
  double p2 = 0.0, p1 = c_dblfact[m];        // Q^m recursion
  double z2 = 0.0, z1 = 0.0;                 // Q^{m+1} recursion (dQ^m/dz): zero at l = m
  for (int l = m; l <= L; l++) {
    double pl, zl;
    if (l == m) { pl = p1; zl = 0.0; }
    else {
      pl = ((double)(2 * l - 1) * uz * p1 - (double)(l + m - 1) * p2) * c_invint[l - m];
      p2 = p1; p1 = pl;
      if (l == m + 1) zl = c_dblfact[m + 1];
      else zl = ((double)(2 * l - 1) * uz * z1 - (double)(l + m) * z2) * c_invint[l - m - 1];
      z2 = z1; z1 = zl;
    }
    const double nrm = ynorm[l * (l + 1) / 2 + m];
    const double q = pl * nrm, qz = zl * nrm;
    const double* rl = rf + l * n4;
    const double* dl = drf + l * n4;
    const int base = l * l + l;
    {  // cos type (or m = 0)
      const double* lp = lam + (size_t)(base + m) * K1p;
      double A = 0.0, B = 0.0;
      for (int a = 0; a < n4; a += 2) {
        const double2 lv = *reinterpret_cast<const double2*>(lp + a);
        const double2 rv = *reinterpret_cast<const double2*>(rl + a);
        const double2 dv = *reinterpret_cast<const double2*>(dl + a);
        A += lv.x * dv.x + lv.y * dv.y;
        B += lv.x * rv.x + lv.y * rv.y;
      }
      SA += A * (q * Cm);
      G0 += B * (q * dm * Cm1);
      G1 -= B * (q * dm * Sm1);
      G2 += B * (qz * Cm);
    }
    if (m > 0) {  // sin type
      const double* lp = lam + (size_t)(base - m) * K1p;
      double A = 0.0, B = 0.0;
      for (int a = 0; a < n4; a += 2) {
        const double2 lv = *reinterpret_cast<const double2*>(lp + a);
        const double2 rv = *reinterpret_cast<const double2*>(rl + a);
        const double2 dv = *reinterpret_cast<const double2*>(dl + a);
        A += lv.x * dv.x + lv.y * dv.y;
        B += lv.x * rv.x + lv.y * rv.y;
      }
      SA += A * (q * Sm);
      G0 += B * (q * dm * Sm1);
      G1 += B * (q * dm * Cm1);
      G2 += B * (qz * Sm);
    }
  }
}

__global__ void __launch_bounds__(NT, 4) k_soap_adjoint(const SoapDev* __restrict__ sp, const int* __restrict__ centres,
                                                        const int* __restrict__ n_centres_dev,
                                                        const int* __restrict__ nbr_off, const int* __restrict__ nbr_j,
                                                        const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                        const int* __restrict__ Z, Lattice9 lat, const double* __restrict__ x,
                                                        const double* __restrict__ xlm, const double* __restrict__ pnorm,
                                                        const double* __restrict__ gvec, int ldg, int g_splits, size_t g_split_stride,
                                                        const double* __restrict__ epart, int n_tiles_n, double* __restrict__ local_e,
                                                        double e_scale, double* __restrict__ force, double* __restrict__ vir_part,
                                                        double* __restrict__ local_virial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem s;
  carve(*sp, true, &s, smem_raw);
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) {
    if (vir_part && threadIdx.x < 9) vir_part[9 * (size_t)c + threadIdx.x] = 0.0;  // unused slot of the upper-bound grid
    return;
  }
  const int i = centres[c];
  // E_i = sum over the column tiles of GEMM-1 (fixed order) ; local_e(centre) += E_i  (IPModel_GAP.f95:454-459, cc = 1, |ci| = 1)
  if (epart && threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < n_tiles_n; k++) t += epart[(size_t)c * n_tiles_n + k];
    local_e[i] += e_scale * t;
  }
  const int n = sp->n_max, L = sp->l_max, L1 = L + 1, nlm = sp->nlm, K1 = sp->K1, d = sp->d, d_pad = sp->d_pad, ns = sp->n_species;
  const int n4 = ceil4(n), K1p = ns * n4, RFS = rf_stride(L1, n4), MP = m_pairs(L);
  const double alpha = sp->alpha;
  (void)d_pad;
  load_tables(sp, s);
  // u = dE/dp: pull gradPredict back through x = p/|p| (reference forward form: descriptors.f95:8595-8600)
  const double* xr = x + (size_t)c * sp->d_pad;
  const double* gr = gvec + (size_t)c * ldg;
  // gradPredict arrives as g_splits partial sums (the K splits of GEMM-2), added here in a fixed order
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += NT) {
    double g = gr[q];
    for (int k = 1; k < g_splits; k++) g += gr[(size_t)k * g_split_stride + q];
    s.p[q] = g;
    loc += xr[q] * g;
  }
  double sdot = block_sum(loc, s.red);
  double nrm = pnorm[c];
  if (sp->normalise)
    for (int q = threadIdx.x; q < d - 1; q += NT) s.p[q] = (s.p[q] - xr[q] * sdot) / nrm;
  // X_lm and Lambda = dE/dX_lm are staged in the (not yet used) radial-table area
  double* Xs = s.rf;                       // nlm*K1
  double* Ls = s.rf + (size_t)nlm * K1;    // nlm*K1
  for (int k = threadIdx.x; k < nlm * K1; k += NT) Xs[k] = xlm[(size_t)c * nlm * K1 + k];
  for (int k = threadIdx.x; k < nlm * K1p; k += NT) s.X[k] = 0.0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < nlm * K1; idx += NT) {
    int lm = idx / K1, ia = idx - lm * K1, l = s.lof[lm];
    double t = 0.0;
    for (int jb = 0; jb < K1; jb++) {
      int hi = ia > jb ? ia : jb, lo = ia > jb ? jb : ia;
      double u = s.p[l + L1 * (hi * (hi + 1) / 2 + lo)];
      t += (ia == jb ? 2.0 * u : 1.41421356237309504880 * u) * Xs[lm * K1 + jb];
    }
    Ls[idx] = t * sp->tlpo[l];
  }
  __syncthreads();
  // Lambda~_lm(s, a) = sum_a' T(a, a') Lambda_lm(s, a')  : the basis transform pulled back once per centre
  for (int idx = threadIdx.x; idx < nlm * K1; idx += NT) {
    int lm = idx / K1, ic = idx - lm * K1, sk = ic / n, a = ic - sk * n;
    const double* lrow = Ls + (size_t)lm * K1 + sk * n;
    double t = 0.0;
    for (int b = 0; b < n; b++) t += s.T[a + n * b] * lrow[b];
    s.X[(size_t)lm * K1p + sk * n4 + a] = t;
  }
  __syncthreads();

  double fi[3] = {0, 0, 0}, vir[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int pbeg = nbr_off[i], pend = nbr_off[i + 1];
  for (int pb = pbeg; pb < pend; pb += NBCAP) {
    int nv = gather_neighbours(sp, s, i, pb, min(pb + NBCAP, pend), nbr_j, nbr_s, pos, Z, lat);
    for (int t0 = 0; t0 < nv; t0 += TNA) {
      const int tn = min(TNA, nv - t0);
      // ---- stage: radial tables with derivative ----
      for (int it = threadIdx.x; it < tn * n4; it += NT) {
        int q = it / n4, a = it - q * n4;
        if (a < n)
          radial_item<true>(alpha, s.nbr[t0 + q], s.rb[a], s.nbf[t0 + q], s.nbdf[t0 + q], L, s.rf + (size_t)q * RFS + a,
                            s.drf + (size_t)q * RFS + a, n4);
        else
          for (int l = 0; l < L1; l++) s.rf[(size_t)q * RFS + l * n4 + a] = s.drf[(size_t)q * RFS + l * n4 + a] = 0.0;
      }
      __syncthreads();
      // ---- contract: item (m-pair, neighbour); orders are paired (m, L+1-m) so that every item has about L+2 (l,m) terms ----
      for (int it = threadIdx.x; it < tn * MP; it += NT) {
        const int mp = it / tn, q = it - mp * tn;
        const double r = s.nbr[t0 + q], rinv = 1.0 / r;
        const double ux = s.nbd[3 * (t0 + q)] * rinv, uy = s.nbd[3 * (t0 + q) + 1] * rinv, uz = s.nbd[3 * (t0 + q) + 2] * rinv;
        const double* lam = s.X + s.nbs[t0 + q] * n4;
        const double* rfq = s.rf + (size_t)q * RFS;
        const double* drq = s.drf + (size_t)q * RFS;
        double SA = 0, G0 = 0, G1 = 0, G2 = 0;
        adjoint_order(s.ynorm, L, mp, ux, uy, uz, lam, K1p, rfq, drq, n4, SA, G0, G1, G2);
        const int m2 = L + 1 - mp;
        if (mp > 0 && m2 > mp) adjoint_order(s.ynorm, L, m2, ux, uy, uz, lam, K1p, rfq, drq, n4, SA, G0, G1, G2);
        const double ug = ux * G0 + uy * G1 + uz * G2;
        double* pp = s.part + ((size_t)q * MP + mp) * 3;
        pp[0] = SA * ux + (G0 - ux * ug) * rinv;
        pp[1] = SA * uy + (G1 - uy * ug) * rinv;
        pp[2] = SA * uz + (G2 - uz * ug) * rinv;
      }
      __syncthreads();
      // ---- scatter: one thread per neighbour ----
      if (threadIdx.x < tn) {
        const int q = threadIdx.x;
        double f0 = 0, f1 = 0, f2 = 0;
        for (int mp = 0; mp < MP; mp++) {
          const double* pp = s.part + ((size_t)q * MP + mp) * 3;
          f0 += pp[0]; f1 += pp[1]; f2 += pp[2];
        }
        f0 *= e_scale; f1 *= e_scale; f2 *= e_scale;
        // IPModel_GAP.f95:479-491: F_j -= f_gp ; centre row is minus the sum ; W_j -= (pos_j - pos_i) (x) f_gp
        const int j = s.nbj[t0 + q];
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
      // the next tile's stage phase overwrites rf/drf only after the barrier at its end; part is rewritten after that barrier too
    }
  }
  // combine the per-thread centre force / virial partials in fixed order (threads 0..TNA-1 hold them)
  for (int k = 0; k < 3; k++) fi[k] = warp_sum(fi[k]);
  for (int k = 0; k < 9; k++) vir[k] = warp_sum(vir[k]);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
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

void launch_soap_forward(const SoapDev* sp, const SoapDev& h, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off,
                         const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, double* x, double* xlm, double* pnorm,
                         cudaStream_t st, int* launches) {
  const int n_centres = n_centres_ub;
  if (n_centres <= 0) return;
  size_t sm = soap_forward_smem(h);
  cudaFuncSetAttribute(k_soap_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_forward<<<n_centres, NT, sm, st>>>(sp, centres, n_centres_dev, nbr_off, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm);
  *launches += 1;
}

void launch_soap_adjoint(const SoapDev* sp, const SoapDev& h, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off,
                         const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, const double* x, const double* xlm,
                         const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride, const double* epart, int n_tiles_n,
                         double* local_e, double e_scale, double* force, double* vir_part, double* local_virial, cudaStream_t st,
                         int* launches) {
  const int n_centres = n_centres_ub;
  if (n_centres <= 0) return;
  size_t sm = soap_adjoint_smem(h);
  cudaFuncSetAttribute(k_soap_adjoint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_adjoint<<<n_centres, NT, sm, st>>>(sp, centres, n_centres_dev, nbr_off, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg, g_splits,
                                            g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial);
  *launches += 1;
}

}  // namespace gapb200
