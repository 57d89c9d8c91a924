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
// Two ways through this file:
//  * compression modes on the EQUISPACED_GAUSS basis: X_lm and the neighbour phase do not depend on the variant, so they run on soap.cu's
//    kernels (forward with skip_power, adjoint with lambda_in); k_soap_power_gen (X_lm -> descriptor) and k_soap_lambda_gen (dE/dx ->
//    Lambda) supply the variant-specific middle, one small CTA per centre;
//  * GTO / POLY (radial grid of 3 n_max points, per-l maps): the same, one run of soap.cu's kernels per slice of <= 16 grid points, with
//    k_soap_power_grid / k_soap_lambda_grid in between;
//  * average=T (and GTO / POLY in deterministic mode): k_soap_forward_gen / k_soap_adjoint_gen, one CTA per centre.
//    The neighbours of a centre are compacted and processed in BATCHES: the radial recursions and the harmonics of a whole batch run as
//    independent items across the CTA's threads (one barrier per batch instead of two per neighbour); the density accumulation and the
//    adjoint contraction of Lambda~ (stored with the lm index contiguous) against the batch are small FP64 tensor-core GEMMs per l.
// Fixed-order reductions throughout; only the final force scatter uses FP64 atomics (per-slot stores in deterministic mode).
#include "gap_device.cuh"
#include "soap_device.cuh"

namespace gapb200 {

namespace {

constexpr int GNT = 128;      // threads per CTA of the general kernels
constexpr int GNW = GNT / 32;

// fixed-order block sum over the GNW warps; result broadcast to all threads
__device__ __forceinline__ double gblock_sum(double v, double* red /* >= GNW doubles */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < GNW; k++) t += red[k];
  return t;
}

struct GSmem {
  double *ynorm, *Xt, *X, *Y1, *Y2, *dY1, *dY2, *p, *Phi, *Rr, *Yq, *Gq, *red, *part, *nbd, *nbr, *nbf, *nbdf, *tlpo;
  int *nbs, *nbj, *nbp, *l_of, *cnt;
};
// NB = neighbours per batch: Phi / Rr hold NB x (L+1) x n_grid radial values, Yq / Gq NB x nlm (x 3) harmonics.  NB = 0: no neighbour phase
// (the kernels that only turn X_lm into the descriptor or dE/dx into Lambda): the radial-grid, batch and neighbour arrays are left out
__host__ __device__ inline size_t gcarve(int L, int n, int ns, int d_pad, const SoapGenDev& g, bool adjoint, int NB, GSmem* s, unsigned char* base) {
  const int nlm = (L + 1) * (L + 1), K1 = ns * n, Kg = ns * g.n_grid;
  const size_t nbn = NB > 0 ? GNT : 0;
  const bool with_xt = NB != 0;  // NB = -1: the radial-grid array without the batch tables (k_soap_power_grid)
  if (NB < 0) NB = 0;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o += ((cnt + 1) & ~(size_t)1) * sizeof(double); return r; };
  const size_t oyn = take((size_t)(L + 1) * (L + 2) / 2), oXt = take(with_xt ? (size_t)nlm * Kg : 0), oX = take((size_t)nlm * K1),
               oY1 = take((size_t)nlm * g.Ka), oY2 = take((size_t)nlm * g.Kb), odY1 = take(adjoint ? (size_t)nlm * g.Ka : 0),
               odY2 = take(adjoint ? (size_t)nlm * g.Kb : 0), op = take(d_pad), oPhi = take((size_t)NB * (L + 1) * g.n_grid),
               oRr = take(adjoint ? (size_t)NB * (L + 1) * g.n_grid : 0), oYq = take((size_t)NB * nlm), oGq = take(adjoint ? 3 * (size_t)NB * nlm : 0),
               ored = take(64), opart = take(adjoint ? (size_t)NB * GNW * 4 : 0), onbd = take(3 * nbn), onbr = take(nbn), onbf = take(nbn),
               onbdf = take(nbn), otl = take(L + 1);
  const size_t oi = o;
  o += sizeof(int) * (3 * nbn + nlm + 8);
  o = (o + 15) & ~(size_t)15;
  if (s) {
    s->ynorm = (double*)(base + oyn); s->Xt = (double*)(base + oXt); s->X = (double*)(base + oX); s->Y1 = (double*)(base + oY1);
    s->Y2 = (double*)(base + oY2); s->dY1 = (double*)(base + odY1); s->dY2 = (double*)(base + odY2); s->p = (double*)(base + op);
    s->Phi = (double*)(base + oPhi); s->Rr = (double*)(base + oRr); s->Yq = (double*)(base + oYq); s->Gq = (double*)(base + oGq);
    s->red = (double*)(base + ored); s->part = (double*)(base + opart); s->nbd = (double*)(base + onbd); s->nbr = (double*)(base + onbr);
    s->nbf = (double*)(base + onbf); s->nbdf = (double*)(base + onbdf); s->tlpo = (double*)(base + otl);
    s->nbs = (int*)(base + oi); s->nbj = s->nbs + nbn; s->nbp = s->nbj + nbn; s->l_of = s->nbp + nbn; s->cnt = s->l_of + nlm;
  }
  return o;
}

__device__ __forceinline__ void g_tables(const SoapDev* sp, const GSmem& s, int L, int nlm) {
  for (int k = threadIdx.x; k < (L + 1) * (L + 2) / 2; k += GNT) s.ynorm[k] = sp->ynorm[k];
  for (int k = threadIdx.x; k <= L; k += GNT) s.tlpo[k] = sp->tlpo[k];
  for (int lm = threadIdx.x; lm < nlm; lm += GNT) {
    int l = 0;
    while ((l + 1) * (l + 1) <= lm) l++;
    s.l_of[lm] = l;
  }
}

// up to GNT CSR entries of centre i; the accepted ones (inside the cutoff, mapped species: descriptors.f95:8190, 8194-8195) are COMPACTED to
// the front of the shared arrays in list order (warp ballots + the four warp counts).  Returns their number.  Two barriers inside: the
// first also separates the previous chunk's readers from this chunk's writers.
__device__ __forceinline__ int g_load_chunk(const SoapDev* sp, const GSmem& s, int i, int p0, int pend, const int* __restrict__ nbr_j,
                                            const int* __restrict__ nbr_s, const double* __restrict__ pos, const int* __restrict__ Z, const Lattice9& lat) {
  const int p = p0 + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  bool valid = false;
  double dd[3] = {0.0, 0.0, 0.0}, r = 0.0, f = 0.0, df = 0.0;
  int spc = -1, j = 0;
  if (p < pend) {
    j = nbr_j[p];
    int s0, s1, s2;
    unpack_shift(nbr_s[p], s0, s1, s2);
    image_diff(pos + 3 * (size_t)i, pos + 3 * (size_t)j, lat.v, s0, s1, s2, dd);
    r = norm_nofma(dd);
    spc = species_of(sp, Z[j]);
    valid = r < sp->cutoff && spc >= 0;
    if (valid) cutoff_fn(sp, r, f, df);
  }
  const unsigned m = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) s.cnt[w] = __popc(m);
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int k = 0; k < GNW; k++) {
    const int ck = s.cnt[k];
    if (k < w) base += ck;
    total += ck;
  }
  if (valid) {
    const int q = base + __popc(m & ((1u << lane) - 1u));
    s.nbd[3 * q] = dd[0]; s.nbd[3 * q + 1] = dd[1]; s.nbd[3 * q + 2] = dd[2];
    s.nbr[q] = r; s.nbf[q] = f; s.nbdf[q] = df; s.nbs[q] = spc; s.nbj[q] = j; s.nbp[q] = p;
  }
  __syncthreads();
  return total;
}

// radial recursions (descriptors.f95:8218-8258) and harmonics of the batch [b0, b0 + nb) as independent items: first the nb x n_grid radial
// items, then the nb x MP harmonic items, so that all but one warp run a single code path
template <bool GRAD>
__device__ __forceinline__ void g_batch_items(const GSmem& s, const SoapGenDev& g, double alpha, int L, int nlm, int ng, int MP, int b0, int nb) {
  const int L1 = L + 1, nrad = nb * ng, nit = nrad + nb * MP;
  for (int w = threadIdx.x; w < nit; w += GNT) {
    if (w < nrad) {
      const int b = w / ng, gg = w - b * ng, q = b0 + b;
      radial_item<GRAD>(alpha, s.nbr[q], g.r_grid[gg], s.nbf[q], GRAD ? s.nbdf[q] : 0.0, L, s.Phi + (size_t)b * L1 * ng + gg,
                        GRAD ? s.Rr + (size_t)b * L1 * ng + gg : nullptr, ng);
    } else {
      const int w2 = w - nrad, b = w2 / MP, it = w2 - b * MP, q = b0 + b;
      const double rinv = 1.0 / s.nbr[q];
      ylm_item<GRAD>(s.ynorm, L, it, s.nbd[3 * q] * rinv, s.nbd[3 * q + 1] * rinv, s.nbd[3 * q + 2] * rinv, s.Yq + (size_t)b * nlm,
                     GRAD ? s.Gq + (size_t)b * 3 * nlm : nullptr, nlm);
    }
  }
}

// Y1 = X W1, Y2 = X W2 (descriptors.f95:8384-8394)
__device__ __forceinline__ void g_mix(const GSmem& s, const SoapGenDev& g, int nlm, int K1) {
  for (int idx = threadIdx.x; idx < nlm * (g.Ka + g.Kb); idx += GNT) {
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
// one thread per output, its elements (grouped by channel on the host) visited in list order: deterministic, no shared-memory atomics
__device__ __forceinline__ void g_dY(const GSmem& s, const SoapGenDev& g, int nlm, int L1) {
  for (int idx = threadIdx.x; idx < nlm * (g.Ka + g.Kb); idx += GNT) {
    const bool second = idx >= nlm * g.Ka;
    const int t = second ? idx - nlm * g.Ka : idx, Kw = second ? g.Kb : g.Ka;
    const int lm = t / Kw, k = t - lm * Kw, l = s.l_of[lm];
    const double tl = s.tlpo[l];
    const int* off = second ? g.by_jb_off : g.by_ia_off;
    const int* lst = second ? g.by_jb : g.by_ia;
    double acc = 0.0;
    for (int q = off[k]; q < off[k + 1]; q++) {
      const int e = lst[q];
      const double w = s.p[l + L1 * e] * tl * g.pair_fac[e];
      acc += w * (second ? s.Y1[lm * g.Ka + g.pair_ia[e]] : s.Y2[lm * g.Kb + g.pair_jb[e]]);
    }
    (second ? s.dY2 : s.dY1)[t] = acc;
  }
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// X (shared) -> Y1 = X W1, Y2 = X W2 -> the element list of the power spectrum -> normalised descriptor row c (descriptors.f95:8384-8451)
__device__ __forceinline__ void g_power_tail(const SoapDev* __restrict__ sp, const SoapGenDev& g, const GSmem& s, int c, double* __restrict__ x,
                                             double* __restrict__ pnorm) {
  const int L = sp->l_max, nlm = (L + 1) * (L + 1), K1 = sp->n_species * sp->n_max, d = sp->d, d_pad = sp->d_pad, L1 = L + 1, np = g.n_pairs;
  g_mix(s, g, nlm, K1);
  __syncthreads();
  double loc = 0.0;
  for (int idx = threadIdx.x; idx < L1 * np; idx += GNT) {  // element l + (l_max+1) k (:8396-8447)
    const int k = idx / L1, l = idx - k * L1, ia = g.pair_ia[k], jb = g.pair_jb[k];
    double acc = 0.0;
    for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) acc += s.Y1[lm * g.Ka + ia] * s.Y2[lm * g.Kb + jb];
    const double v = acc * s.tlpo[l] * g.pair_fac[k];
    s.p[idx] = v;
    loc += v * v;
  }
  const double nrm = sqrt(gblock_sum(loc, s.red));  // :8450-8451
  const double inv = sp->normalise ? 1.0 / nrm : 1.0;
  double* xr = x + (size_t)c * d_pad;
  for (int q = threadIdx.x; q < d_pad; q += GNT) xr[q] = q < d - 1 ? s.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (threadIdx.x == 0) pnorm[c] = nrm;
}

// u = dE/dp from the K-split partials of gradPredict, pulled back through the normalisation, the element list and the mixing matrices:
// leaves Lambda = dE/dX_lm [lm][K1] in s.X (synchronised)
__device__ __forceinline__ void g_lambda_head(const SoapDev* __restrict__ sp, const SoapGenDev& g, const GSmem& s, int c, const double* __restrict__ x,
                                              const double* __restrict__ xlm, const double* __restrict__ pnorm, const double* __restrict__ gvec, int ldg,
                                              int g_splits, size_t g_split_stride) {
  const int L = sp->l_max, nlm = (L + 1) * (L + 1), K1 = sp->n_species * sp->n_max, d = sp->d, d_pad = sp->d_pad, L1 = L + 1;
  for (int k = threadIdx.x; k < nlm * K1; k += GNT) s.X[k] = xlm[(size_t)c * nlm * K1 + k];
  for (int k = threadIdx.x; k < nlm * g.Ka; k += GNT) s.dY1[k] = 0.0;
  for (int k = threadIdx.x; k < nlm * g.Kb; k += GNT) s.dY2[k] = 0.0;
  // u = dE/dp: gradPredict (the K-split partials of GEMM-2 added in a fixed order) pulled back through x = p / |p| (:8595-8600)
  const double* xr = x + (size_t)c * d_pad;
  const double* gr = gvec + (size_t)c * ldg;
  const double nrm = pnorm[c];
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += GNT) {
    double gv = gr[q];
    for (int k = 1; k < g_splits; k++) gv += gr[(size_t)k * g_split_stride + q];
    s.p[q] = gv;
    loc += xr[q] * gv;
  }
  const double sdot = gblock_sum(loc, s.red);  // (synchronises: X and the zeroed dY are visible afterwards)
  if (sp->normalise)
    for (int q = threadIdx.x; q < d - 1; q += GNT) s.p[q] = (s.p[q] - xr[q] * sdot) / nrm;
  g_mix(s, g, nlm, K1);
  __syncthreads();
  g_dY(s, g, nlm, L1);  // dE/dY1, dE/dY2 (product rule on the element list)
  __syncthreads();
  // Lambda = dE/dX = dE/dY1 W1^T + dE/dY2 W2^T  (into X)
  for (int idx = threadIdx.x; idx < nlm * K1; idx += GNT) {
    const int lm = idx / K1, ic = idx - lm * K1;
    double acc = 0.0;
    for (int k = 0; k < g.Ka; k++) acc += s.dY1[lm * g.Ka + k] * g.W1[(size_t)ic * g.Ka + k];
    for (int k = 0; k < g.Kb; k++) acc += s.dY2[lm * g.Kb + k] * g.W2[(size_t)ic * g.Kb + k];
    s.X[idx] = acc;
  }
  __syncthreads();
}

// Xt (shared, radial functions on the grid) -> X = Xt . P_l + central term -> xlm (kept for the adjoint) -> descriptor row c
__device__ __forceinline__ void g_forward_post(const SoapDev* __restrict__ sp, const SoapGenDev& g, const GSmem& s, int c, int i, const int* __restrict__ Z,
                                               double* __restrict__ x, double* __restrict__ xlm, double* __restrict__ pnorm, int global_mode) {
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng;
  // radial_coefficient = radial_fun . P_l (:8261-8270); the map is linear, so it is applied once per centre
  for (int idx = threadIdx.x; idx < nlm * K1; idx += GNT) {
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
  for (int k = threadIdx.x; k < nlm * K1; k += GNT) xlm[(size_t)c * nlm * K1 + k] = s.X[k];
  if (global_mode) return;  // average=T: the power spectrum is taken of the SUM over the centres (k_soap_global_power)
  g_power_tail(sp, g, s, c, x, pnorm);
}

__global__ void __launch_bounds__(GNT) k_soap_forward_gen(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ centres,
                                                         const int* __restrict__ n_centres_dev, const int* __restrict__ nbr_off,
                                                         const int* __restrict__ nbr_end, const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                         const double* __restrict__ pos, const int* __restrict__ Z, Lattice9 lat, double* __restrict__ x,
                                                         double* __restrict__ xlm, double* __restrict__ pnorm, int global_mode, int NB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng, d = sp->d,
            d_pad = sp->d_pad, L1 = L + 1, MP = L / 2 + 1 + (L & 1), np = g.n_pairs;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, false, NB, &s, smem_raw);
  const int i = centres[c];
  const double alpha = sp->alpha;
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * Kg; k += GNT) s.Xt[k] = 0.0;
  __syncthreads();
  const int pbeg = nbr_off[i], pend = nbr_end[i], lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pb = pbeg; pb < pend; pb += GNT) {
    const int cnt = g_load_chunk(sp, s, i, pb, pend, nbr_j, nbr_s, pos, Z, lat);
    for (int b0 = 0; b0 < cnt; b0 += NB) {
      const int nb = min(NB, cnt - b0);
      g_batch_items<false>(s, g, alpha, L, nlm, ng, MP, b0, nb);
      __syncthreads();
      // Xt_lm(s, g) += sum_b Y_lm(b) f Phi_l(b; g) (:8289-8295 before the radial map) on the FP64 tensor cores: for every l the
      // (2l+1) x n_grid block is a small GEMM over the neighbours of the batch (A = Y, 8 lm rows x 4 neighbours; B = Phi, 4 neighbours x 8
      // grid points; one species at a time, the other species' neighbours masked out of A).  The 8 x 8 tiles are dealt to the warps
      // round-robin; each tile's result is added to its Xt elements by the lanes that hold it.
      {
        const int fr = lane >> 2, fk = lane & 3, ntn = (ng + 7) >> 3;
        int tile = 0;
        for (int l = 0; l <= L; l++) {
          const int rows = 2 * l + 1, rbase = l * l;
          for (int mt = 0; mt * 8 < rows; mt++)
            for (int nt = 0; nt < ntn; nt++)
              for (int sk = 0; sk < ns; sk++, tile++) {
                if ((tile & (GNW - 1)) != warp) continue;
                const bool rv = mt * 8 + fr < rows, gv = nt * 8 + fr < ng;
                const int lm = rbase + mt * 8 + fr;
                const double* ph = s.Phi + l * ng + nt * 8 + fr;
                double c0 = 0.0, c1 = 0.0;
                for (int k0 = 0; k0 < nb; k0 += 4) {
                  const int b = k0 + fk;
                  const bool bv = b < nb;
                  const double a = (rv && bv && (ns == 1 || s.nbs[b0 + b] == sk)) ? s.Yq[b * nlm + lm] : 0.0;
                  const double bb = (gv && bv) ? ph[(size_t)b * L1 * ng] : 0.0;
                  dmma(c0, c1, a, bb);
                }
                if (rv) {
                  const int g0 = nt * 8 + 2 * fk;
                  double* dst = s.Xt + lm * Kg + sk * ng + g0;
                  if (g0 < ng) dst[0] += c0;
                  if (g0 + 1 < ng) dst[1] += c1;
                }
              }
        }
      }
      __syncthreads();
    }
  }
  g_forward_post(sp, g, s, c, i, Z, x, xlm, pnorm, global_mode);
}

template <int NB>
__global__ void __launch_bounds__(GNT) k_soap_adjoint_gen(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ centres,
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
  gcarve(L, n, ns, d_pad, g, true, NB, &s, smem_raw);
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
    for (int k = threadIdx.x; k < nlm * Kg; k += GNT) {  // stored with lm contiguous (see the neighbour phase)
      const int lm = k / Kg, t = k - lm * Kg;
      s.Xt[t * nlm + lm] = Lt_global[k];
    }
    __syncthreads();
  } else {
  g_lambda_head(sp, g, s, c, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride);
  // Lambda~ = Lambda P_l^T on the radial grid, stored as Xt[(s, g)][lm]: lm contiguous
  for (int idx = threadIdx.x; idx < nlm * Kg; idx += GNT) {
    const int lm = idx / Kg, t = idx - lm * Kg, sk = t / ng, gg = t - sk * ng, l = s.l_of[lm];
    double acc = 0.0;
    for (int a = 0; a < n; a++) acc += s.X[lm * K1 + sk * n + a] * g.P[((size_t)l * ng + gg) * n + a];
    s.Xt[t * nlm + lm] = acc;
  }
  __syncthreads();
  }  // !Lt_global
  // neighbour phase: f_gp,k = sum_lm sum_g Lambda~_lm(s, g) d/dr_k [ f Phi_l(g) Y_lm ]   (IPModel_GAP.f95:479 without grad_data)
  double acc12[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // thread b < NB: centre force (3) and virial (9) of the neighbours it finalises
  const int pbeg = nbr_off[i], pend = nbr_end[i], lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pb = pbeg; pb < pend; pb += GNT) {
    const int cnt = g_load_chunk(sp, s, i, pb, pend, nbr_j, nbr_s, pos, Z, lat);
    for (int b0 = 0; b0 < cnt; b0 += NB) {
      const int nb = min(NB, cnt - b0);
      g_batch_items<true>(s, g, alpha, L, nlm, ng, MP, b0, nb);
      __syncthreads();
      // one thread per lm: A = sum_g Lambda~ R_l(g), B = sum_g Lambda~ Phi_l(g) for every neighbour of the batch, then the four lm sums
      //   SA = sum_lm Y_lm A,  G_k = sum_lm (grad Y_lm)_k B
      if constexpr (NB == 8) {
        // on the FP64 tensor cores: per l a GEMM over the radial index, A = Lambda~ (8 lm rows x 4 (s, g)), B = R_l or Phi_l of the batch
        // (4 (s, g) x 8 neighbours, zero outside the neighbour's own species); the lanes multiply their part of the 8 x 8 result by
        // Y_lm / grad Y_lm and keep partial lm sums for their two neighbours; tiles dealt to the warps round-robin
        const int fr = lane >> 2, fk = lane & 3;
        double pS[2] = {0.0, 0.0}, p0[2] = {0.0, 0.0}, p1[2] = {0.0, 0.0}, p2[2] = {0.0, 0.0};
        const bool bv = fr < nb;
        const int sb = bv ? s.nbs[b0 + fr] : -1;
        int tile = 0;
        for (int l = 0; l <= L; l++) {
          const int rows = 2 * l + 1, rbase = l * l;
          for (int mt = 0; mt * 8 < rows; mt++, tile++) {
            if ((tile & (GNW - 1)) != warp) continue;
            const bool rv = mt * 8 + fr < rows;
            const int lm = rbase + mt * 8 + fr;
            const double* rr = s.Rr + (size_t)fr * L1 * ng + l * ng;
            const double* ph = s.Phi + (size_t)fr * L1 * ng + l * ng;
            double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
            for (int k0 = 0; k0 < Kg; k0 += 4) {
              const int k = k0 + fk;
              const double a = (rv && k < Kg) ? s.Xt[(size_t)k * nlm + lm] : 0.0;
              double br = 0.0, bp = 0.0;
              if (k < Kg) {
                const int sk = k / ng, gg = k - sk * ng;
                if (sk == sb) { br = rr[gg]; bp = ph[gg]; }
              }
              dmma(a0, a1, a, br);
              dmma(c0, c1, a, bp);
            }
            if (rv) {
#pragma unroll
              for (int jj = 0; jj < 2; jj++) {
                const int bj = 2 * fk + jj;
                if (bj < nb) {
                  const double* gq = s.Gq + (size_t)bj * 3 * nlm + lm;
                  const double A = jj ? a1 : a0, B = jj ? c1 : c0;
                  pS[jj] += A * s.Yq[bj * nlm + lm];
                  p0[jj] += B * gq[0];
                  p1[jj] += B * gq[nlm];
                  p2[jj] += B * gq[2 * nlm];
                }
              }
            }
          }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1)
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            pS[jj] += __shfl_xor_sync(0xffffffffu, pS[jj], o);
            p0[jj] += __shfl_xor_sync(0xffffffffu, p0[jj], o);
            p1[jj] += __shfl_xor_sync(0xffffffffu, p1[jj], o);
            p2[jj] += __shfl_xor_sync(0xffffffffu, p2[jj], o);
          }
        if (fr == 0)
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            const int bj = 2 * fk + jj;
            if (bj < nb) {
              double* pp = s.part + (bj * GNW + warp) * 4;
              pp[0] = pS[jj]; pp[1] = p0[jj]; pp[2] = p1[jj]; pp[3] = p2[jj];
            }
          }
      } else {
      double SA[NB], G0[NB], G1[NB], G2[NB];
#pragma unroll
      for (int b = 0; b < NB; b++) SA[b] = G0[b] = G1[b] = G2[b] = 0.0;
      for (int lm = threadIdx.x; lm < nlm; lm += GNT) {
        const int lo = s.l_of[lm] * ng;
#pragma unroll
        for (int b = 0; b < NB; b++) {
          if (b < nb) {
            const double* lam = s.Xt + (size_t)(s.nbs[b0 + b] * ng) * nlm + lm;
            const double* ph = s.Phi + (size_t)b * L1 * ng + lo;
            const double* rr = s.Rr + (size_t)b * L1 * ng + lo;
            double A = 0.0, B = 0.0;
            for (int gg = 0; gg < ng; gg++) {
              const double la = lam[(size_t)gg * nlm];
              A += la * rr[gg];
              B += la * ph[gg];
            }
            const double* gq = s.Gq + (size_t)b * 3 * nlm + lm;
            SA[b] += A * s.Yq[b * nlm + lm];
            G0[b] += B * gq[0];
            G1[b] += B * gq[nlm];
            G2[b] += B * gq[2 * nlm];
          }
        }
      }
#pragma unroll
      for (int b = 0; b < NB; b++) {
        if (b < nb) {
          const double v0 = warp_sum(SA[b]), v1 = warp_sum(G0[b]), v2 = warp_sum(G1[b]), v3 = warp_sum(G2[b]);
          if (lane == 0) {
            double* pp = s.part + (b * GNW + warp) * 4;
            pp[0] = v0; pp[1] = v1; pp[2] = v2; pp[3] = v3;
          }
        }
      }
      }  // NB != 8
      __syncthreads();
      if (threadIdx.x < nb) {  // thread b finalises neighbour b: the four warps' partials in a fixed order
        const int b = threadIdx.x, q = b0 + b;
        const double* pp = s.part + b * GNW * 4;
        double sa = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
        for (int wq = 0; wq < GNW; wq++) { sa += pp[4 * wq]; g0 += pp[4 * wq + 1]; g1 += pp[4 * wq + 2]; g2 += pp[4 * wq + 3]; }
        const double r = s.nbr[q], rinv = 1.0 / r, dx = s.nbd[3 * q], dy = s.nbd[3 * q + 1], dz = s.nbd[3 * q + 2];
        const double ux = dx * rinv, uy = dy * rinv, uz = dz * rinv;
        const double ug = ux * g0 + uy * g1 + uz * g2;
        // f_gp,k = sum_lm sum_g Lambda~_lm(s, g) d/dr_k [ f Phi_l(g) Y_lm ]  (IPModel_GAP.f95:479 without grad_data);
        // grad Y = (g - u (u.g)) / r  (GradSphericalYCartesian_all, angular_functions.f95:205-278, from the polynomial extension)
        const double f0 = (sa * ux + (g0 - ux * ug) * rinv) * e_scale;
        const double f1 = (sa * uy + (g1 - uy * ug) * rinv) * e_scale;
        const double f2 = (sa * uz + (g2 - uz * ug) * rinv) * e_scale;
        const int j = s.nbj[q];
        if (force) {  // IPModel_GAP.f95:479-491: F_j -= f_gp ; the centre row is minus the sum ; W_j -= (pos_j - pos_i) (x) f_gp
          if (fpair) {  // deterministic scatter (see gap_device.cuh): slot = position of the entry in the neighbour list
            const size_t slot = (size_t)s.nbp[q];
            fpair[3 * slot + 0] = -f0; fpair[3 * slot + 1] = -f1; fpair[3 * slot + 2] = -f2;
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
      __syncthreads();  // the next batch overwrites the batch tables and the partials
    }
  }
  // centre force and virial: the NB finalising threads' sums, added in thread order
  if (threadIdx.x < NB)
#pragma unroll
    for (int k = 0; k < 12; k++) s.part[threadIdx.x * 12 + k] = acc12[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[12];
    for (int k = 0; k < 12; k++) {
      double t = 0.0;
      for (int b = 0; b < NB; b++) t += s.part[b * 12 + k];
      tot[k] = t;
    }
    if (force && fpair)
      for (int k = 0; k < 3; k++) force[3 * (size_t)i + k] = tot[k];
    else if (force)
      for (int k = 0; k < 3; k++) atomicAdd(&force[3 * (size_t)i + k], tot[k]);
    if (vir_part)
      for (int k = 0; k < 9; k++) vir_part[9 * (size_t)c + k] = tot[3 + k];
  }
}

// ---- compression modes on the EQUISPACED_GAUSS basis: the density expansion X_lm and the neighbour phase are those of the default power
//      spectrum, so they run on soap.cu's DMMA kernels (forward with skip_power, adjoint with lambda_in); only the channel mixing and the
//      element list are different, and these two small kernels supply them: X_lm -> descriptor, and dE/dx -> Lambda = dE/dX_lm.
__global__ void __launch_bounds__(GNT) k_soap_power_gen(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ n_centres_dev,
                                                        const double* __restrict__ xlm, double* __restrict__ x, double* __restrict__ pnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;
  const int L = sp->l_max, nlm = (L + 1) * (L + 1), K1 = sp->n_species * sp->n_max;
  GSmem s;
  gcarve(L, sp->n_max, sp->n_species, sp->d_pad, g, false, 0, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * K1; k += GNT) s.X[k] = xlm[(size_t)c * nlm * K1 + k];
  __syncthreads();
  g_power_tail(sp, g, s, c, x, pnorm);
}

__global__ void __launch_bounds__(GNT) k_soap_lambda_gen(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ n_centres_dev,
                                                         const double* __restrict__ x, const double* __restrict__ xlm, const double* __restrict__ pnorm,
                                                         const double* __restrict__ gvec, int ldg, int g_splits, size_t g_split_stride,
                                                         double* __restrict__ lambda_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;
  const int L = sp->l_max, nlm = (L + 1) * (L + 1), K1 = sp->n_species * sp->n_max;
  GSmem s;
  gcarve(L, sp->n_max, sp->n_species, sp->d_pad, g, true, 0, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  g_lambda_head(sp, g, s, c, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride);
  for (int k = threadIdx.x; k < nlm * K1; k += GNT) lambda_out[(size_t)c * nlm * K1 + k] = s.X[k];
}

// ---- GTO / POLY on the default path's kernels: the radial functions of these bases are the same Phi_l(r; r_g) on a grid of 3 n_max points,
//      followed by a per-l map to n_max functions.  The grid is cut into passes of at most 16 points; each pass is a run of soap.cu's
//      kernels on a clone of the descriptor whose "basis points" are that slice of the grid, with the identity as basis transform and no
//      central term (forward with skip_power: xlm = Xt on the slice; adjoint with lambda_in = Lambda~ on the slice).  These two kernels
//      sit in between: slices of Xt -> X = Xt . P_l + central term -> descriptor, and dE/dx -> Lambda -> Lambda~ = Lambda P_l^T in slices.
//      Pass buffers: [pass][centre][lm][s * gp + a], pass_stride doubles apart.
__global__ void __launch_bounds__(GNT) k_soap_power_grid(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ centres,
                                                         const int* __restrict__ n_centres_dev, const int* __restrict__ Z,
                                                         const double* __restrict__ xt_pass, size_t pass_stride, int gp, double* __restrict__ x,
                                                         double* __restrict__ xlm, double* __restrict__ pnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;
  const int L = sp->l_max, ns = sp->n_species, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng, K1p = ns * gp;
  GSmem s;
  gcarve(L, sp->n_max, ns, sp->d_pad, g, false, -1, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * Kg; k += GNT) {
    const int lm = k / Kg, t = k - lm * Kg, sk = t / ng, gg = t - sk * ng, pass = gg / gp, a = gg - pass * gp;
    s.Xt[k] = xt_pass[(size_t)pass * pass_stride + ((size_t)c * nlm + lm) * K1p + sk * gp + a];
  }
  __syncthreads();
  g_forward_post(sp, g, s, c, centres[c], Z, x, xlm, pnorm, 0);
}

__global__ void __launch_bounds__(GNT) k_soap_lambda_grid(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ n_centres_dev,
                                                          const double* __restrict__ x, const double* __restrict__ xlm, const double* __restrict__ pnorm,
                                                          const double* __restrict__ gvec, int ldg, int g_splits, size_t g_split_stride,
                                                          double* __restrict__ lam_pass, size_t pass_stride, int gp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, nlm = (L + 1) * (L + 1), K1 = ns * n, ng = g.n_grid, Kg = ns * ng, K1p = ns * gp;
  GSmem s;
  gcarve(L, n, ns, sp->d_pad, g, true, 0, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  g_lambda_head(sp, g, s, c, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride);
  for (int idx = threadIdx.x; idx < nlm * Kg; idx += GNT) {  // Lambda~ = Lambda P_l^T on the radial grid, written slice by slice
    const int lm = idx / Kg, t = idx - lm * Kg, sk = t / ng, gg = t - sk * ng, l = s.l_of[lm], pass = gg / gp, a = gg - pass * gp;
    double acc = 0.0;
    for (int b = 0; b < n; b++) acc += s.X[lm * K1 + sk * n + b] * g.P[((size_t)l * ng + gg) * n + b];
    lam_pass[(size_t)pass * pass_stride + ((size_t)c * nlm + lm) * K1p + sk * gp + a] = acc;
  }
}

// ---- average=T (global SOAP, descriptors.f95:8357-8367, 8738-9008): ONE descriptor per configuration from the sum of the density
//      expansions of all centres.  One CTA each: the sum + power spectrum, and the pull-back dE/dx -> Lambda~ shared by all centres.
__global__ void __launch_bounds__(GNT) k_soap_global_power(const SoapDev* __restrict__ sp, SoapGenDev g, const int* __restrict__ n_centres_dev,
                                                          const double* __restrict__ xlm, double* __restrict__ Xg, double* __restrict__ x,
                                                          double* __restrict__ pnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), d = sp->d, d_pad = sp->d_pad, L1 = L + 1,
            np = g.n_pairs, nc = *n_centres_dev;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, false, 1, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * K1; k += GNT) {  // fixed order over the centres: deterministic
    double acc = 0.0;
    for (int c = 0; c < nc; c++) acc += xlm[(size_t)c * nlm * K1 + k];
    s.X[k] = acc;
    Xg[k] = acc;
  }
  __syncthreads();
  g_mix(s, g, nlm, K1);
  __syncthreads();
  double loc = 0.0;
  for (int idx = threadIdx.x; idx < L1 * np; idx += GNT) {
    const int k = idx / L1, l = idx - k * L1, ia = g.pair_ia[k], jb = g.pair_jb[k];
    double acc = 0.0;
    for (int lm = l * l; lm < (l + 1) * (l + 1); lm++) acc += s.Y1[lm * g.Ka + ia] * s.Y2[lm * g.Kb + jb];
    const double v = acc * s.tlpo[l] * g.pair_fac[k];
    s.p[idx] = v;
    loc += v * v;
  }
  double nrm = sqrt(gblock_sum(loc, s.red));
  if (nrm == 0.0) nrm = 2.2250738585072014e-308;  // tiny(1.0_dp), :8842
  const double inv = sp->normalise ? 1.0 / nrm : 1.0;
  for (int q = threadIdx.x; q < d_pad; q += GNT) x[q] = q < d - 1 ? s.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (threadIdx.x == 0) pnorm[0] = nrm;
}

__global__ void __launch_bounds__(GNT) k_soap_global_lambda(const SoapDev* __restrict__ sp, SoapGenDev g, const double* __restrict__ Xg,
                                                           const double* __restrict__ x, const double* __restrict__ pnorm,
                                                           const double* __restrict__ gvec, int g_splits, size_t g_split_stride,
                                                           double* __restrict__ Lt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = sp->l_max, n = sp->n_max, ns = sp->n_species, K1 = ns * n, nlm = (L + 1) * (L + 1), ng = g.n_grid, Kg = ns * ng, d = sp->d,
            d_pad = sp->d_pad, L1 = L + 1, np = g.n_pairs;
  GSmem s;
  gcarve(L, n, ns, d_pad, g, true, 1, &s, smem_raw);
  g_tables(sp, s, L, nlm);
  for (int k = threadIdx.x; k < nlm * K1; k += GNT) s.X[k] = Xg[k];
  for (int k = threadIdx.x; k < nlm * g.Ka; k += GNT) s.dY1[k] = 0.0;
  for (int k = threadIdx.x; k < nlm * g.Kb; k += GNT) s.dY2[k] = 0.0;
  const double nrm = pnorm[0];
  double loc = 0.0;
  for (int q = threadIdx.x; q < d - 1; q += GNT) {
    double gv = gvec[q];
    for (int k = 1; k < g_splits; k++) gv += gvec[(size_t)k * g_split_stride + q];
    s.p[q] = gv;
    loc += x[q] * gv;
  }
  const double sdot = gblock_sum(loc, s.red);
  if (sp->normalise)
    for (int q = threadIdx.x; q < d - 1; q += GNT) s.p[q] = (s.p[q] - x[q] * sdot) / nrm;
  g_mix(s, g, nlm, K1);
  __syncthreads();
  g_dY(s, g, nlm, L1);
  __syncthreads();
  for (int idx = threadIdx.x; idx < nlm * K1; idx += GNT) {
    const int lm = idx / K1, ic = idx - lm * K1;
    double acc = 0.0;
    for (int k = 0; k < g.Ka; k++) acc += s.dY1[lm * g.Ka + k] * g.W1[(size_t)ic * g.Ka + k];
    for (int k = 0; k < g.Kb; k++) acc += s.dY2[lm * g.Kb + k] * g.W2[(size_t)ic * g.Kb + k];
    s.X[idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nlm * Kg; idx += GNT) {
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

// smallest working set (one neighbour per batch): what a shape needs at least
size_t soap_general_smem(const SoapDev& h, const SoapGenDev& g) { return gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, true, 1, nullptr, nullptr); }

namespace {
// neighbours per batch: as many as keep two CTAs on an SM (about 110 KiB each), at least one.  (Measured on the config-A cell: 16 per batch
// forward beats 8 with a third resident CTA -- fewer barriers win over occupancy; 256-thread CTAs are faster for n_grid = 24, slower for 8.)
int pick_batch(const SoapDev& h, const SoapGenDev& g, bool adjoint, int nb_max) {
  int nb = nb_max;
  while (nb > 1 && gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, adjoint, nb, nullptr, nullptr) > 110 * 1024) nb >>= 1;
  return nb;
}
template <int NB>
void launch_adjoint_nb(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                       const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                       const double* x, const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride,
                       const double* epart, int n_tiles_n, double* local_e, double e_scale, double* force, double* vir_part, double* local_virial,
                       const double* Lt, double* fpair, cudaStream_t st) {
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, true, NB, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_adjoint_gen<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_adjoint_gen<NB><<<n_centres_ub, GNT, sm, st>>>(sp, g, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg,
                                                       g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, Lt, fpair);
}
}  // namespace

void launch_soap_forward_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                                 const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                                 double* x, double* xlm, double* pnorm, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const int NB = pick_batch(h, g, false, 16);
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, false, NB, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_forward_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_forward_gen<<<n_centres_ub, GNT, sm, st>>>(sp, g, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm,
                                                   g.global_mode, NB);
  *launches += 1;
  if (g.global_mode) {
    const size_t sm2 = soap_general_smem(h, g);
    cudaFuncSetAttribute(k_soap_global_power, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    k_soap_global_power<<<1, GNT, sm2, st>>>(sp, g, n_centres_dev, xlm, g.Xg, x, pnorm);
    *launches += 1;
  }
}

void launch_soap_adjoint_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                                 const int* nbr_off, const int* nbr_end, const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat,
                                 const double* x, const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride,
                                 const double* epart, int n_tiles_n, double* local_e, double e_scale, double* force, double* vir_part,
                                 double* local_virial, double* fpair, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const double* Lt = nullptr;
  if (g.global_mode) {  // one descriptor: its energy is shared by all centres, its dE/dX by all neighbour phases
    if (epart) {
      k_global_energy<<<(n_centres_ub + 255) / 256, 256, 0, st>>>(epart, n_tiles_n, centres, n_centres_dev, e_scale, local_e);
      *launches += 1;
    }
    if (!gvec) return;  // energy only
    const size_t sm = soap_general_smem(h, g);
    cudaFuncSetAttribute(k_soap_global_lambda, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_soap_global_lambda<<<1, GNT, sm, st>>>(sp, g, g.Xg, x, pnorm, gvec, g_splits, g_split_stride, g.Lt);
    *launches += 1;
    Lt = g.Lt;
    epart = nullptr;
    n_tiles_n = 0;
  }
  const int NB = pick_batch(h, g, true, 8);
#define GAP_ADJ_GEN(B)                                                                                                                                  \
  launch_adjoint_nb<B>(sp, h, g, centres, n_centres_dev, n_centres_ub, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg, g_splits, \
                       g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, Lt, fpair, st)
  if (NB >= 8) GAP_ADJ_GEN(8);
  else if (NB >= 4) GAP_ADJ_GEN(4);
  else if (NB >= 2) GAP_ADJ_GEN(2);
  else GAP_ADJ_GEN(1);
#undef GAP_ADJ_GEN
  *launches += 1;
}

void launch_soap_power_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* n_centres_dev, int n_centres_ub, const double* xlm,
                               double* x, double* pnorm, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, false, 0, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_power_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_power_gen<<<n_centres_ub, GNT, sm, st>>>(sp, g, n_centres_dev, xlm, x, pnorm);
  *launches += 1;
}

void launch_soap_lambda_general(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* n_centres_dev, int n_centres_ub, const double* x,
                                const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride,
                                double* lambda_out, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, true, 0, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_lambda_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_lambda_gen<<<n_centres_ub, GNT, sm, st>>>(sp, g, n_centres_dev, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride, lambda_out);
  *launches += 1;
}

void launch_soap_power_grid(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* centres, const int* n_centres_dev, int n_centres_ub,
                            const int* Z, const double* xt_pass, size_t pass_stride, int gp, double* x, double* xlm, double* pnorm, cudaStream_t st,
                            int* launches) {
  if (n_centres_ub <= 0) return;
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, false, -1, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_power_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_power_grid<<<n_centres_ub, GNT, sm, st>>>(sp, g, centres, n_centres_dev, Z, xt_pass, pass_stride, gp, x, xlm, pnorm);
  *launches += 1;
}

void launch_soap_lambda_grid(const SoapDev* sp, const SoapDev& h, const SoapGenDev& g, const int* n_centres_dev, int n_centres_ub, const double* x,
                             const double* xlm, const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride, double* lam_pass,
                             size_t pass_stride, int gp, cudaStream_t st, int* launches) {
  if (n_centres_ub <= 0) return;
  const size_t sm = gcarve(h.l_max, h.n_max, h.n_species, h.d_pad, g, true, 0, nullptr, nullptr);
  cudaFuncSetAttribute(k_soap_lambda_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_soap_lambda_grid<<<n_centres_ub, GNT, sm, st>>>(sp, g, n_centres_dev, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride, lam_pass, pass_stride, gp);
  *launches += 1;
}

}  // namespace gapb200
