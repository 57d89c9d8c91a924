// soap_general.cu -- the GENERAL form of the SOAP power spectrum: channel compression (Z_mix / R_mix / sym_mix with QUIP's random
// weights, coupling=F, Z_map, nu_R / nu_S, diagonal_radial; src/GAP/descriptors.f95:7274-7670) and the GTO / POLY radial bases
// (:2643-2770, :8264-8278).  soap.cu keeps the reference's "original" power spectrum (:7772-7775) on its DMMA kernels; everything
// else comes here.
//
// One formulation covers every variant (the host builds the tables, gap_model.cpp soap_general_setup):
//     Xt_lm(s, g)  = sum_neighbours  Y_lm  f_cut Phi_l(r; r_g)                 radial functions on n_grid points r_g
//     X_lm(s, a)   = sum_g  Xt_lm(s, g) P_l(g, a)  (+ central term)              P_l: transform_basis, or the least-squares map of GTO / POLY
//     Y1 = X W1 ,  Y2 = X W2                                                     channel mixing
//     p(l, k)      = tlpo_l fac_k sum_m Y1_lm(ia_k) Y2_lm(jb_k)                  element list (ia, jb, fac): index l + (l_max + 1) k
// in real spherical harmonics (the power spectrum is invariant under the unitary change from the reference's complex ones).
// Gradients in REVERSE mode, as in soap.cu: dE/dp -> dE/dY1, dE/dY2 -> Lambda = dE/dX -> Lambda~ = Lambda P^T on the radial grid,
// then per neighbour the 3-vector f_gp.  The reference's forward-mode dY / grad_data (:8311-8336, :8470-8555) never exists.
//
// These are research / compression options, not the headline shapes: one CTA per centre, plain FP64 FMAs,
// fixed-order reductions except the shared-memory accumulation of dE/dY and the final force scatter (FP64 atomics).
#include "gap_device.cuh"
#include "soap_device.cuh"

namespace gapb200 {

namespace {

struct GSmem {
  double *ynorm, *Xt, *X, *Y1, *Y2, *dY1, *dY2, *p, *Phi, *Rr, *Yq, *Gq, *red, *nbd, *nbr, *nbf, *nbdf, *tlpo;
  int *nbs, *nbj, *nbv, *l_of;
};
__host__ __device__ inline size_t gcarve(int L, int n, int ns, int d_pad, const SoapGenDev& g, bool adjoint, GSmem* s, unsigned char* base) {
  const int nlm = (L + 1) * (L + 1), K1 = ns * n, Kg = ns * g.n_grid;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o += ((cnt + 1) & ~(size_t)1) * sizeof(double); return r; };
  const size_t oyn = take((size_t)(L + 1) * (L + 2) / 2), oXt = take((size_t)nlm * Kg), oX = take((size_t)nlm * K1), oY1 = take((size_t)nlm * g.Ka),
               oY2 = take((size_t)nlm * g.Kb), odY1 = take(adjoint ? (size_t)nlm * g.Ka : 0), odY2 = take(adjoint ? (size_t)nlm * g.Kb : 0),
               op = take(d_pad), oPhi = take((size_t)(L + 1) * g.n_grid), oRr = take(adjoint ? (size_t)(L + 1) * g.n_grid : 0), oYq = take(nlm),
               oGq = take(adjoint ? 3 * (size_t)nlm : 0), ored = take(64), onbd = take(3 * NT), onbr = take(NT), onbf = take(NT), onbdf = take(NT),
               otl = take(L + 1);
  const size_t oi = o;
  o += sizeof(int) * (3 * NT + nlm);
  o = (o + 15) & ~(size_t)15;
  if (s) {
    s->ynorm = (double*)(base + oyn); s->Xt = (double*)(base + oXt); s->X = (double*)(base + oX); s->Y1 = (double*)(base + oY1);
    s->Y2 = (double*)(base + oY2); s->dY1 = (double*)(base + odY1); s->dY2 = (double*)(base + odY2); s->p = (double*)(base + op);
    s->Phi = (double*)(base + oPhi); s->Rr = (double*)(base + oRr); s->Yq = (double*)(base + oYq); s->Gq = (double*)(base + oGq);
    s->red = (double*)(base + ored); s->nbd = (double*)(base + onbd); s->nbr = (double*)(base + onbr); s->nbf = (double*)(base + onbf);
    s->nbdf = (double*)(base + onbdf); s->tlpo = (double*)(base + otl);
    s->nbs = (int*)(base + oi); s->nbj = s->nbs + NT; s->nbv = s->nbj + NT; s->l_of = s->nbv + NT;
  }
  return o;
}

__device__ __forceinline__ void g_tables(const SoapDev* sp, const GSmem& s, int L, int nlm) {
  for (int k = threadIdx.x; k < (L + 1) * (L + 2) / 2; k += NT) s.ynorm[k] = sp->ynorm[k];
  for (int k = threadIdx.x; k <= L; k += NT) s.tlpo[k] = sp->tlpo[k];
  for (int lm = threadIdx.x; lm < nlm; lm += NT) {
    int l = 0;
    while ((l + 1) * (l + 1) <= lm) l++;
    s.l_of[lm] = l;
  }
}

// up to NT CSR entries of centre i -> shared arrays indexed by the thread (no compaction; nbv marks the accepted ones)
__device__ __forceinline__ void g_load_chunk(const SoapDev* sp, const GSmem& s, int i, int p0, int pend, const int* __restrict__ nbr_j,
                                             const int* __restrict__ nbr_s, const double* __restrict__ pos, const int* __restrict__ Z, const Lattice9& lat) {
  const int p = p0 + threadIdx.x;
  int valid = 0;
  if (p < pend) {
    const int j = nbr_j[p];
    int s0, s1, s2;
    unpack_shift(nbr_s[p], s0, s1, s2);
    double dd[3];
    image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)j, lat.v, s0, s1, s2, dd);
    const double r = norm_nofma(dd);
    const int spc = species_of(sp, Z[j]);
    if (r < sp->cutoff && spc >= 0) {  // descriptors.f95:8190, 8194-8195
      double f, df;
      cutoff_fn(sp, r, f, df);
      s.nbd[3 * threadIdx.x] = dd[0]; s.nbd[3 * threadIdx.x + 1] = dd[1]; s.nbd[3 * threadIdx.x + 2] = dd[2];
      s.nbr[threadIdx.x] = r; s.nbf[threadIdx.x] = f; s.nbdf[threadIdx.x] = df; s.nbs[threadIdx.x] = spc; s.nbj[threadIdx.x] = j;
      valid = 1;
    }
  }
  s.nbv[threadIdx.x] = valid;
}

// Y1 = X W1, Y2 = X W2 (descriptors.f95:8384-8394)
__device__ __forceinline__ void g_mix(const GSmem& s, const SoapGenDev& g, int nlm, int K1) {
  for (int idx = threadIdx.x; idx < nlm * (g.Ka + g.Kb); idx += NT) {
    const bool second = idx >= nlm * g.Ka;
    const int t = second ? idx - nlm * g.Ka : idx, Kw = second ? g.Kb : g.Ka;
    const int lm = t / Kw, k = t - lm * Kw;
    const double* W = second ? g.W2 : g.W1;
    double acc = 0.0;
    for (int ic = 0; ic < K1; ic++) acc += s.X[lm * K1 + ic] * W[(size_t)ic * Kw + k];
    (second ? s.Y2 : s.Y1)[t] = acc;
  }
}

// dE/dY1(lm, ia) = sum over the elements k with ia_k = ia of  w_k Y2(lm, jb_k),  dE/dY2(lm, jb) likewise with Y1,  w_k = tlpo_l fac_k dE/dp(l, k):
// one thread per output, the elements visited in list order (deterministic; no shared-memory atomics)
__device__ __forceinline__ void g_dY(const GSmem& s, const SoapGenDev& g, int nlm, int L1, int np) {
  for (int idx = threadIdx.x; idx < nlm * (g.Ka + g.Kb); idx += NT) {
    const bool second = idx >= nlm * g.Ka;
    const int t = second ? idx - nlm * g.Ka : idx, Kw = second ? g.Kb : g.Ka;
    const int lm = t / Kw, k = t - lm * Kw, l = s.l_of[lm];
    const double tl = s.tlpo[l];
    double acc = 0.0;
    for (int e = 0; e < np; e++) {
      const int ia = g.pair_ia[e], jb = g.pair_jb[e];
      if ((second ? jb : ia) != k) continue;
      const double w = s.p[l + L1 * e] * tl * g.pair_fac[e];
      acc += w * (second ? s.Y1[lm * g.Ka + ia] : s.Y2[lm * g.Kb + jb]);
    }
    (second ? s.dY2 : s.dY1)[t] = acc;
  }
}

__global__ void __launch_bounds__(NT) k_soap_forward_gen(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ centres,
                                                         const int* __restrict__ n_centres_dev, const int* __restrict__ nbr_off,
                                                         const int* __restrict__ nbr_end, const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                         const double* __restrict__ pos, const int* __restrict__ Z, Lattice9 lat, double* __restrict__ x,
                                                         double* __restrict__ xlm, double* __restrict__ pnorm, int global_mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng, d = sp->d,
            d_pad = sp->d_pad, L1 = L + 1, MP = L / 2 + 1 + (L & 1), np = g.n_pairs;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, false, &s, smem_raw);
  const int i = centres[c];
  const double alpha = sp->alpha;
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * Kg; k += NT) s.Xt[k] = 0.0;
  __syncthreads();
  const int pbeg = nbr_off[i], pend = nbr_end[i];
  for (int pb = pbeg; pb < pend; pb += NT) {
    g_load_chunk(sp, s, i, pb, pend, nbr_j, nbr_s, pos, Z, lat);
    __syncthreads();
    const int cnt = min(NT, pend - pb);
    for (int q = 0; q < cnt; q++) {
      if (!s.nbv[q]) continue;
      const double r = s.nbr[q], f = s.nbf[q], rinv = 1.0 / r;
      const int sq = s.nbs[q];
      for (int it = threadIdx.x; it < ng + MP; it += NT) {
        if (it < ng) radial_item<false>(alpha, r, g.r_grid[it], f, 0.0, L, s.Phi + it, nullptr, ng);  // :8218-8258
        else ylm_item<false>(s.ynorm, L, it - ng, s.nbd[3 * q] * rinv, s.nbd[3 * q + 1] * rinv, s.nbd[3 * q + 2] * rinv, s.Yq, nullptr, 0);
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < nlm * ng; idx += NT) {  // :8289-8295 before the radial map
        const int lm = idx / ng, gg = idx - lm * ng;
        s.Xt[lm * Kg + sq * ng + gg] += s.Yq[lm] * s.Phi[s.l_of[lm] * ng + gg];
      }
      __syncthreads();
    }
  }
  // radial_coefficient = radial_fun . P_l (:8261-8270); the map is linear, so it is applied once per centre
  for (int idx = threadIdx.x; idx < nlm * K1; idx += NT) {
    const int lm = idx / K1, ic = idx - lm * K1, sk = ic / n, a = ic - sk * n, l = s.l_of[lm];
    double acc = 0.0;
    for (int gg = 0; gg < ng; gg++) acc += s.Xt[lm * Kg + sk * ng + gg] * g.P[((size_t)l * ng + gg) * n + a];
    s.X[idx] = acc;
  }
  __syncthreads();
  if (threadIdx.x < K1) {  // central atom term (:8151-8182)
    const int sk = threadIdx.x / n, a = threadIdx.x - sk * n;
    if (sp->cras || sp->species_Z[sk] == Z[i] || sp->species_Z[sk] == 0) s.X[threadIdx.x] += sp->central_weight * g.c0[a] * 0.28209479177387814347;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nlm * K1; k += NT) xlm[(size_t)c * nlm * K1 + k] = s.X[k];
  if (global_mode) return;  // average=T: the power spectrum is taken of the SUM over the centres (k_soap_global_power)
  g_mix(s, g, nlm, K1);
  __syncthreads();
  double loc = 0.0;
  for (int idx = threadIdx.x; idx < L1 * np; idx += NT) {  // element l + (l_max+1) k (:8396-8447)
    const int k = idx / L1, l = idx - k * L1, ia = g.pair_ia[k], jb = g.pair_jb[k];
    double acc = 0.0;
    for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) acc += s.Y1[lm * g.Ka + ia] * s.Y2[lm * g.Kb + jb];
    const double v = acc * s.tlpo[l] * g.pair_fac[k];
    s.p[idx] = v;
    loc += v * v;
  }
  const double nrm = sqrt(block_sum(loc, s.red));  // :8450-8451
  const double inv = sp->normalise ? 1.0 / nrm : 1.0;
  double* xr = x + (size_t)c * d_pad;
  for (int q = threadIdx.x; q < d_pad; q += NT) xr[q] = q < d - 1 ? s.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (threadIdx.x == 0) pnorm[c] = nrm;
}

__global__ void __launch_bounds__(NT) k_soap_adjoint_gen(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ centres,
                                                         const int* __restrict__ n_centres_dev, const int* __restrict__ nbr_off,
                                                         const int* __restrict__ nbr_end, const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                         const double* __restrict__ pos, const int* __restrict__ Z, Lattice9 lat,
                                                         const double* __restrict__ x, const double* __restrict__ xlm, const double* __restrict__ pnorm,
                                                         const double* __restrict__ gvec, int ldg, int g_splits, size_t g_split_stride,
                                                         const double* __restrict__ epart, int n_tiles_n, double* __restrict__ local_e, double e_scale,
                                                         double* __restrict__ force, double* __restrict__ vir_part, double* __restrict__ local_virial,
                                                         const double* __restrict__ Lt_global, double* __restrict__ fpair) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) {
    if (vir_part && threadIdx.x < 9) vir_part[9 * (size_t)c + threadIdx.x] = 0.0;  // unused slot of the upper-bound grid
    return;
  }
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng, d = sp->d,
            d_pad = sp->d_pad, L1 = L + 1, MP = L / 2 + 1 + (L & 1), np = g.n_pairs;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, true, &s, smem_raw);
  const int i = centres[c];
  const double alpha = sp->alpha;
  if (epart && threadIdx.x < 32) {  // E_i = sum over the column tiles of GEMM-1 (fixed order); local_e(centre) += E_i (IPModel_GAP.f95:454-459)
    double t = 0.0;
    for (int k = threadIdx.x; k < n_tiles_n; k += 32) t += epart[(size_t)c * n_tiles_n + k];
    t = warp_sum(t);
    if (threadIdx.x == 0) local_e[i] += e_scale * t;
  }
  g_tables(sp, s, L, nlm);
  if (Lt_global) {  // average=T: dE/dX on the radial grid is the same for every centre (k_soap_global_lambda)
    for (int k = threadIdx.x; k < nlm * Kg; k += NT) s.Xt[k] = Lt_global[k];
    __syncthreads();
  } else {
  for (int k = threadIdx.x; k < nlm * K1; k += NT) s.X[k] = xlm[(size_t)c * nlm * K1 + k];
  for (int k = threadIdx.x; k < nlm * g.Ka; k += NT) s.dY1[k] = 0.0;
  for (int k = threadIdx.x; k < nlm * g.Kb; k += NT) s.dY2[k] = 0.0;
  // u = dE/dp: gradPredict (the K-split partials of GEMM-2 added in a fixed order) pulled back through x = p / |p| (:8595-8600)
  const double* xr = x + (size_t)c * d_pad;
  const double* gr = gvec + (size_t)c * ldg;
  const double nrm = pnorm[c];
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += NT) {
    double gv = gr[q];
    for (int k = 1; k < g_splits; k++) gv += gr[(size_t)k * g_split_stride + q];
    s.p[q] = gv;
    loc += xr[q] * gv;
  }
  const double sdot = block_sum(loc, s.red);  // (synchronises: X and the zeroed dY are visible afterwards)
  if (sp->normalise)
    for (int q = threadIdx.x; q < d - 1; q += NT) s.p[q] = (s.p[q] - xr[q] * sdot) / nrm;
  g_mix(s, g, nlm, K1);
  __syncthreads();
  g_dY(s, g, nlm, L1, np);  // dE/dY1, dE/dY2 (product rule on the element list)
  __syncthreads();
  // Lambda = dE/dX = dE/dY1 W1^T + dE/dY2 W2^T  (into X)
  for (int idx = threadIdx.x; idx < nlm * K1; idx += NT) {
    const int lm = idx / K1, ic = idx - lm * K1;
    double acc = 0.0;
    for (int k = 0; k < g.Ka; k++) acc += s.dY1[lm * g.Ka + k] * g.W1[(size_t)ic * g.Ka + k];
    for (int k = 0; k < g.Kb; k++) acc += s.dY2[lm * g.Kb + k] * g.W2[(size_t)ic * g.Kb + k];
    s.X[idx] = acc;
  }
  __syncthreads();
  // Lambda~ = Lambda P_l^T on the radial grid (into Xt)
  for (int idx = threadIdx.x; idx < nlm * Kg; idx += NT) {
    const int lm = idx / Kg, t = idx - lm * Kg, sk = t / ng, gg = t - sk * ng, l = s.l_of[lm];
    double acc = 0.0;
    for (int a = 0; a < n; a++) acc += s.X[lm * K1 + sk * n + a] * g.P[((size_t)l * ng + gg) * n + a];
    s.Xt[idx] = acc;
  }
  __syncthreads();
  }  // !Lt_global
  // neighbour phase: f_gp,k = sum_lm sum_g Lambda~_lm(s, g) d/dr_k [ f Phi_l(g) Y_lm ]   (IPModel_GAP.f95:479 without grad_data)
  double acc12[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // thread 0: centre force (3), virial (9)
  const int pbeg = nbr_off[i], pend = nbr_end[i];
  for (int pb = pbeg; pb < pend; pb += NT) {
    g_load_chunk(sp, s, i, pb, pend, nbr_j, nbr_s, pos, Z, lat);
    __syncthreads();
    const int cnt = min(NT, pend - pb);
    for (int q = 0; q < cnt; q++) {
      if (!s.nbv[q]) continue;
      const double r = s.nbr[q], f = s.nbf[q], df = s.nbdf[q], rinv = 1.0 / r;
      const double dx = s.nbd[3 * q], dy = s.nbd[3 * q + 1], dz = s.nbd[3 * q + 2];
      const double ux = dx * rinv, uy = dy * rinv, uz = dz * rinv;
      const int sq = s.nbs[q];
      for (int it = threadIdx.x; it < ng + MP; it += NT) {
        if (it < ng) radial_item<true>(alpha, r, g.r_grid[it], f, df, L, s.Phi + it, s.Rr + it, ng);
        else ylm_item<true>(s.ynorm, L, it - ng, ux, uy, uz, s.Yq, s.Gq, nlm);
      }
      __syncthreads();
      double SA = 0.0, G0 = 0.0, G1 = 0.0, G2 = 0.0;
      for (int idx = threadIdx.x; idx < nlm * ng; idx += NT) {
        const int lm = idx / ng, gg = idx - lm * ng, l = s.l_of[lm];
        const double lam = s.Xt[lm * Kg + sq * ng + gg];
        SA += lam * s.Rr[l * ng + gg] * s.Yq[lm];
        const double t = lam * s.Phi[l * ng + gg];
        G0 += t * s.Gq[lm];
        G1 += t * s.Gq[nlm + lm];
        G2 += t * s.Gq[2 * nlm + lm];
      }
      SA = warp_sum(SA); G0 = warp_sum(G0); G1 = warp_sum(G1); G2 = warp_sum(G2);
      if ((threadIdx.x & 31) == 0) {
        double* rr = s.red + 4 * (threadIdx.x >> 5);
        rr[0] = SA; rr[1] = G0; rr[2] = G1; rr[3] = G2;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        SA = (s.red[0] + s.red[4]) + (s.red[8] + s.red[12]);
        G0 = (s.red[1] + s.red[5]) + (s.red[9] + s.red[13]);
        G1 = (s.red[2] + s.red[6]) + (s.red[10] + s.red[14]);
        G2 = (s.red[3] + s.red[7]) + (s.red[11] + s.red[15]);
        const double ug = ux * G0 + uy * G1 + uz * G2;
        // grad Y = (g - u (u.g)) / r  (GradSphericalYCartesian_all, angular_functions.f95:205-278, from the polynomial extension)
        const double f0 = (SA * ux + (G0 - ux * ug) * rinv) * e_scale;
        const double f1 = (SA * uy + (G1 - uy * ug) * rinv) * e_scale;
        const double f2 = (SA * uz + (G2 - uz * ug) * rinv) * e_scale;
        const int j = s.nbj[q];
        if (force) {  // IPModel_GAP.f95:479-491: F_j -= f_gp ; the centre row is minus the sum ; W_j -= (pos_j - pos_i) (x) f_gp
          if (fpair) {  // deterministic scatter (see gap_device.cuh): slot = position of the entry in the neighbour list
            const size_t pp = (size_t)(pb + q);
            fpair[3 * pp + 0] = -f0; fpair[3 * pp + 1] = -f1; fpair[3 * pp + 2] = -f2;
          } else {
            atomicAdd(&force[3 * (size_t)j + 0], -f0);
            atomicAdd(&force[3 * (size_t)j + 1], -f1);
            atomicAdd(&force[3 * (size_t)j + 2], -f2);
          }
          acc12[0] += f0; acc12[1] += f1; acc12[2] += f2;
        }
        const double wv[9] = {dx * f0, dy * f0, dz * f0, dx * f1, dy * f1, dz * f1, dx * f2, dy * f2, dz * f2};  // column-major (a + 3b)
#pragma unroll
        for (int k = 0; k < 9; k++) acc12[3 + k] -= wv[k];
        if (local_virial)
#pragma unroll
          for (int k = 0; k < 9; k++) atomicAdd(&local_virial[9 * (size_t)j + k], -wv[k]);
      }
    }
    __syncthreads();  // the next chunk overwrites the neighbour arrays
  }
  if (threadIdx.x == 0) {
    if (force && fpair)
      for (int k = 0; k < 3; k++) force[3 * (size_t)i + k] = acc12[k];
    else if (force)
      for (int k = 0; k < 3; k++) atomicAdd(&force[3 * (size_t)i + k], acc12[k]);
    if (vir_part)
      for (int k = 0; k < 9; k++) vir_part[9 * (size_t)c + k] = acc12[3 + k];
  }
}

// ---- average=T (global SOAP, descriptors.f95:8357-8367, 8738-9008): ONE descriptor per configuration from the sum of the density
//      expansions of all centres.  One CTA each: the sum + power spectrum, and the pull-back dE/dx -> Lambda~ shared by all centres.
__global__ void __launch_bounds__(NT) k_soap_global_power(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ n_centres_dev,
                                                          const double* __restrict__ xlm, double* __restrict__ Xg, double* __restrict__ x,
                                                          double* __restrict__ pnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), d = sp->d, d_pad = sp->d_pad, L1 = L + 1,
            np = g.n_pairs, nc = *n_centres_dev;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, false, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * K1; k += NT) {  // fixed order over the centres: deterministic
    double acc = 0.0;
    for (int c = 0; c < nc; c++) acc += xlm[(size_t)c * nlm * K1 + k];
    s.X[k] = acc;
    Xg[k] = acc;
  }
  __syncthreads();
  g_mix(s, g, nlm, K1);
  __syncthreads();
  double loc = 0.0;
  for (int idx = threadIdx.x; idx < L1 * np; idx += NT) {
    const int k = idx / L1, l = idx - k * L1, ia = g.pair_ia[k], jb = g.pair_jb[k];
    double acc = 0.0;
    for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) acc += s.Y1[lm * g.Ka + ia] * s.Y2[lm * g.Kb + jb];
    const double v = acc * s.tlpo[l] * g.pair_fac[k];
    s.p[idx] = v;
    loc += v * v;
  }
  double nrm = sqrt(block_sum(loc, s.red));
  if (nrm == 0.0) nrm = 2.2250738585072014e-308;  // tiny(1.0_dp), :8842
  const double inv = sp->normalise ? 1.0 / nrm : 1.0;
  for (int q = threadIdx.x; q < d_pad; q += NT) x[q] = q < d - 1 ? s.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (threadIdx.x == 0) pnorm[0] = nrm;
}

__global__ void __launch_bounds__(NT) k_soap_global_lambda(const SoapDev* __restrict__ sp, SoapGenDev g, const double* __restrict__ Xg,
                                                           const double* __restrict__ x, const double* __restrict__ pnorm,
                                                           const double* __restrict__ gvec, int g_splits, size_t g_split_stride,
                                                           double* __restrict__ Lt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng, d = sp->d,
            d_pad = sp->d_pad, L1 = L + 1, np = g.n_pairs;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, true, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * K1; k += NT) s.X[k] = Xg[k];
  for (int k = threadIdx.x; k < nlm * g.Ka; k += NT) s.dY1[k] = 0.0;
  for (int k = threadIdx.x; k < nlm * g.Kb; k += NT) s.dY2[k] = 0.0;
  const double nrm = pnorm[0];
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += NT) {
    double gv = gvec[q];
    for (int k = 1; k < g_splits; k++) gv += gvec[(size_t)k * g_split_stride + q];
    s.p[q] = gv;
    loc += x[q] * gv;
  }
  const double sdot = block_sum(loc, s.red);
  if (sp->normalise)
    for (int q = threadIdx.x; q < d - 1; q += NT) s.p[q] = (s.p[q] - x[q] * sdot) / nrm;
  g_mix(s, g, nlm, K1);
  __syncthreads();
  g_dY(s, g, nlm, L1, np);
  __syncthreads();
  for (int idx = threadIdx.x; idx < nlm * K1; idx += NT) {
    const int lm = idx / K1, ic = idx - lm * K1;
    double acc = 0.0;
    for (int k = 0; k < g.Ka; k++) acc += s.dY1[lm * g.Ka + k] * g.W1[(size_t)ic * g.Ka + k];
    for (int k = 0; k < g.Kb; k++) acc += s.dY2[lm * g.Kb + k] * g.W2[(size_t)ic * g.Kb + k];
    s.X[idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nlm * Kg; idx += NT) {
    const int lm = idx / Kg, t = idx - lm * Kg, sk = t / ng, gg = t - sk * ng, l = s.l_of[lm];
    double acc = 0.0;
    for (int a = 0; a < n; a++) acc += s.X[lm * K1 + sk * n + a] * g.P[((size_t)l * ng + gg) * n + a];
    Lt[idx] = acc;
  }
}

// local_e(ci(n)) += e_i / size(ci) for every centre (IPModel_GAP.f95:454-459 with the global descriptor's ci = all centres)
__global__ void k_global_energy(const double* __restrict__ epart, int n_tiles_n, const int* __restrict__ centres, const int* __restrict__ n_centres_dev,
                                double e_scale, double* __restrict__ local_e) {
  const int nc = *n_centres_dev, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double t = 0.0;
  for (int k = 0; k < n_tiles_n; k++) t += epart[k];
  local_e[centres[c]] += e_scale * t / (double)nc;
}

}  // namespace

size_t soap_general_smem(const SoapDev& h, const SoapGenDev& g) { return gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, true, nullptr, nullptr); }

void launch_soap_forward_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                                 const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                                 double* x, double* xlm, double* pnorm, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, false, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_forward_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_forward_gen<<<n_centres_ub, NT, sm, st>>>(sp, g, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm,
                                                   g.global_mode);
  *launches += 1;
  if (g.global_mode) {
    const size_t sm2 = soap_general_smem(h, g);
    cudaFuncSetAttribute(k_soap_global_power, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    k_soap_global_power<<<1, NT, sm2, st>>>(sp, g, n_centres_dev, xlm, g.Xg, x, pnorm);
    *launches += 1;
  }
}

void launch_soap_adjoint_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                                 const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                                 const double* x, const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride,
                                 const double* epart, int n_tiles_n, double* local_e, double e_scale, double* force, double* vir_part,
                                 double* local_virial, double* fpair, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const size_t sm = soap_general_smem(h, g);
  cudaFuncSetAttribute(k_soap_adjoint_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (g.global_mode) {  // one descriptor: its energy is shared by all centres, its dE/dX by all neighbour phases
    if (epart) {
      k_global_energy<<<(n_centres_ub + 255) / 256, 256, 0, st>>>(epart, n_tiles_n, centres, n_centres_dev, e_scale, local_e);
      *launches += 1;
    }
    if (!gvec) return;  // energy only
    cudaFuncSetAttribute(k_soap_global_lambda, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_soap_global_lambda<<<1, NT, sm, st>>>(sp, g, g.Xg, x, pnorm, gvec, g_splits, g_split_stride, g.Lt);
    *launches += 1;
    k_soap_adjoint_gen<<<n_centres_ub, NT, sm, st>>>(sp, g, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg,
                                                     g_splits, g_split_stride, nullptr, 0, local_e, e_scale, force, vir_part, local_virial, g.Lt, fpair);
    *launches += 1;
    return;
  }
  k_soap_adjoint_gen<<<n_centres_ub, NT, sm, st>>>(sp, g, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg,
                                                   g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, nullptr, fpair);
  *launches += 1;
}

}  // namespace gapb200
