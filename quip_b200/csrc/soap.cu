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
//  * The three dense contractions of the descriptor run on the FP64 TENSOR CORES (mma.sync.m8n8k4.f64, SASS DMMA):
//      density expansion   Xt_lm(c)   = sum_q  Y_lm(q) Phi_l(c; q)          (M = lm, N = channel, K = neighbour)
//      power spectrum      p_l(c,c')  = sum_m  X_lm(c) X_lm(c')             (M = N = channel, K = m)
//      adjoint             A_lm(q)    = sum_c  Lambda~_lm(c) R_l(c; q)      (M = lm, N = neighbour, K = channel)
//    plus the small basis-transform and Lambda products; fragments are read straight from shared memory whose row
//    strides are = 4 (mod 16) doubles, which makes every fragment load bank-conflict free.
//  * The transform_basis product is hoisted out of the neighbour loop (it is linear): X = Xt.T once per centre and
//    Lambda~ = Lambda.T^T once per centre, instead of the reference's per-neighbour matmul (descriptors.f95:8261-8263).
//
// Two kernel families.  Specialised shapes (n_max, l_max, n_species) run WARP-per-centre kernels (k_soap_forward_w,
// k_soap_adjoint_w): a warp owns its centre from the first load to the last store, the radial and harmonic recursions feed
// the DMMA fragments directly from registers, and nothing needs a block barrier (ncu showed a fifth of the block kernels'
// stall samples at barriers).  Every other shape -- and the adjoint of shapes whose per-warp state would leave fewer than
// 12 warps per SM -- runs the generic kernels: one CTA (128 threads = 4 warps) per centre, several CTAs resident per SM; per
// tile of neighbours the threads first evaluate per-neighbour items -- (neighbour, radial basis point) -> Phi_l / R_l by the
// reference's upward recursion, (neighbour, m) -> Y_{l,+-m} (and gradient) by recursion in l -- into shared memory, then
// the warps run the tensor-core contractions over output tiles.  All reductions are fixed-order (deterministic) except the
// final force scatter, which uses FP64 atomics on HBM.
#include "gap_device.cuh"
#include "soap_device.cuh"

namespace gapb200 {

namespace {

// ------------------------------------------------------------------------------------------------------------
// geometry of the shared-memory tables (identical on host and device)
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int ceil4(int v) { return (v + 3) & ~3; }
// Harmonic items per neighbour: the orders m are paired as (m, L+1-m) so that every item runs about L+2 recursion steps
// (order m alone costs L-m+1): item 0 -> m = 0, item k >= 1 -> m = k and, if different and larger, m = L+1-k.
__host__ __device__ inline int m_pairs(int L) { return L / 2 + 1 + (L & 1); }
__host__ __device__ inline int ceil8(int v) { return (v + 7) & ~7; }
__host__ __device__ inline int stride4mod16(int v) {  // smallest s >= v with s = 4 (mod 16)
  int s = (v & ~15) + 4;
  return s >= v ? s : s + 16;
}
struct Geo {
  int n, L, L1, nlm, ns, K1, K18, n2, n8, RFS, YS, XS, XR, TS, NTM, NTN, MP;
};
__host__ __device__ __forceinline__ Geo make_geo(int n_max, int l_max, int n_species) {
  Geo g;
  g.n = n_max; g.L = l_max; g.L1 = l_max + 1; g.nlm = (l_max + 1) * (l_max + 1); g.ns = n_species; g.K1 = n_species * n_max;
  g.K18 = ceil8(g.K1);            // channels padded to DMMA tiles
  g.n2 = (g.n + 1) & ~1;
  g.n8 = ceil8(g.n);
  g.RFS = stride4mod16(g.L1 * g.n2);   // per-neighbour radial table [l][a]
  g.YS = stride4mod16(g.nlm + 7);      // forward: per-neighbour harmonics row = DMMA A operand (tile reads run up to 7 rows past nlm)
  g.MP = m_pairs(g.L);
  g.XS = g.K18 + 4;                    // X / Lambda row stride: = 4 or 12 (mod 16)
  g.XR = g.nlm + 8;                    // rows allocated
  g.TS = stride4mod16(g.n8);           // padded transform_basis row stride
  // lm tiles of 8 rows, never straddling two l: sum_l ceil((2l+1)/8) in closed form: l <= 3 -> 1 tile, 4..7 -> 2, 8..11 -> 3, 12 -> 4
  int ntm = 0;
#pragma unroll
  for (int l = 0; l <= SOAP_LMAX_CAP; l++)
    if (l <= l_max) ntm += (2 * l + 1 + 7) / 8;
  g.NTM = ntm;
  g.NTN = g.K18 / 8;
  return g;
}

struct Smem {
  double* Tp;      // n8 x TS  transform_basis zero-padded: Tp[a*TS + a'] = T(a, a')
  double* rb;      // n
  double* ynorm;   // (L+1)(L+2)/2
  double* invint;  // 16: 1/k (0 for k = 0)
  double* dblf;    // 20: (2k-1)!! (0 beyond the table)
  double* X;       // XR x XS : forward Xt -> X ; adjoint X -> Lambda~   (row lm, column = channel s*n + a)
  double* X2;      // adjoint: XR x XS Lambda
  double* nbd;     // NBCAP*3 displacement
  double* nbr;     // NBCAP distance
  double* nbf;     // NBCAP f_cut
  double* nbdf;    // NBCAP f_cut'
  double* red;     // 64
  double* rf;      // forward: TN x RFS
  double* Y;       // forward: TN x YS
  double* acc;     // adjoint: NW x 8 x 12 centre force / virial partials, one slot per (warp, fragment row)
  double* p;       // d_pad : forward power spectrum (aliases the staging area) ; adjoint u = dE/dp (own region)
  int* nbs;        // NBCAP species
  int* nbj;        // NBCAP neighbour atom
  int* nbp;        // NBCAP slot of the entry in the neighbour list (deterministic scatter: the pair force is stored per slot)
  int* mt_lm0;     // NTM first row of the lm tile
  int* mt_l;       // NTM its l
  int* col_s;      // K18 species of channel (-1: padding)
  int* col_a;      // K18 radial index of channel
  int* wcount;     // NW x SOAP_SPECIES_CAP
  int* seg;        // SOAP_SPECIES_CAP + 1: the compacted neighbours are sorted by species; species k occupies [seg[k], seg[k+1])
};

__host__ __device__ __forceinline__ size_t carve(const Geo& g, int d_pad, bool adjoint, Smem* s, unsigned char* base) {
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o += ((cnt + 1) & ~(size_t)1) * sizeof(double); return r; };
  size_t oT = take((size_t)g.n8 * g.TS), orb = take(g.n), oyn = take((size_t)g.L1 * (g.L1 + 1) / 2), oinv = take(16), odbl = take(20),
         oX = take((size_t)g.XR * g.XS), onbd = take(NBCAP * 3), onbr = take(NBCAP), onbf = take(NBCAP), onbdf = take(NBCAP), ored = take(64);
  size_t ostage = o, orf = 0, oY = 0, oX2 = 0, oacc = 0, op;
  if (adjoint) {
    oX2 = take((size_t)g.XR * g.XS);
    op = take(d_pad);
    oacc = take((size_t)NW * 8 * 12);
  } else {
    orf = take((size_t)TNF * g.RFS);
    oY = take((size_t)TNF * g.YS);
    size_t stage_bytes = o - ostage;
    op = ostage;  // forward: the power spectrum reuses the staging area after the neighbour loop
    if ((size_t)d_pad * sizeof(double) > stage_bytes) o = ostage + (size_t)d_pad * sizeof(double);
  }
  size_t oi = o;
  o += sizeof(int) * (NBCAP * 3 + 2 * g.NTM + 2 * g.K18 + NW * SOAP_SPECIES_CAP + SOAP_SPECIES_CAP + 2);
  o = (o + 15) & ~(size_t)15;
  if (s) {
    s->Tp = (double*)(base + oT); s->rb = (double*)(base + orb); s->ynorm = (double*)(base + oyn); s->invint = (double*)(base + oinv);
    s->dblf = (double*)(base + odbl); s->X = (double*)(base + oX); s->X2 = (double*)(base + oX2);
    s->nbd = (double*)(base + onbd); s->nbr = (double*)(base + onbr); s->nbf = (double*)(base + onbf); s->nbdf = (double*)(base + onbdf);
    s->red = (double*)(base + ored); s->rf = (double*)(base + orf); s->Y = (double*)(base + oY); s->acc = (double*)(base + oacc);
    s->p = (double*)(base + op);
    s->nbs = (int*)(base + oi); s->nbj = s->nbs + NBCAP; s->nbp = s->nbj + NBCAP; s->mt_lm0 = s->nbp + NBCAP; s->mt_l = s->mt_lm0 + g.NTM; s->col_s = s->mt_l + g.NTM;
    s->col_a = s->col_s + g.K18; s->wcount = s->col_a + g.K18; s->seg = s->wcount + NW * SOAP_SPECIES_CAP;
  }
  return o;
}

__device__ __forceinline__ void load_tables(const SoapDev* sp, const Geo& g, const Smem& s) {
  for (int k = threadIdx.x; k < g.n8 * g.TS; k += NT) {
    int a = k / g.TS, b = k - a * g.TS;
    s.Tp[k] = (a < g.n && b < g.n) ? sp->T[a + g.n * b] : 0.0;
  }
  for (int k = threadIdx.x; k < g.n; k += NT) s.rb[k] = sp->r_basis[k];
  for (int k = threadIdx.x; k < g.L1 * (g.L1 + 1) / 2; k += NT) s.ynorm[k] = sp->ynorm[k];
  if (threadIdx.x < 16) s.invint[threadIdx.x] = threadIdx.x <= LC + 1 ? c_invint[threadIdx.x] : 0.0;
  if (threadIdx.x < 20) s.dblf[threadIdx.x] = threadIdx.x <= LC ? c_dblfact[threadIdx.x] : 0.0;
  if (threadIdx.x <= g.L) {  // thread l writes the lm tiles of its l
    const int l = threadIdx.x;
    int t = 0;
    for (int k = 0; k < l; k++) t += (2 * k + 1 + 7) / 8;
    for (int r0 = l * l; r0 < (l + 1) * (l + 1); r0 += 8) { s.mt_lm0[t] = r0; s.mt_l[t] = l; t++; }
  }
  for (int k = threadIdx.x; k < g.K18; k += NT) {
    s.col_s[k] = k < g.K1 ? k / g.n : -1;
    s.col_a[k] = k < g.K1 ? k % g.n : 0;
  }
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Ordered (deterministic) block compaction of up to NBCAP CSR entries of centre i into shared memory, SORTED BY SPECIES
// (stable within a species): species k occupies [seg[k], seg[k+1]), so that a tile of neighbours of one species only
// touches that species' n_max channels.
__device__ __forceinline__ int gather_neighbours(const SoapDev* sp, const Smem& s, int ns, int i, int pbeg, int pend,
                                                 const int* __restrict__ nbr_j, const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                 const int* __restrict__ Z, const Lattice9& lat) {
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
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int myrank = 0;
  for (int k = 0; k < ns; k++) {
    const bool mine = valid && spc == k;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) s.wcount[w * SOAP_SPECIES_CAP + k] = __popc(bal);
    if (mine) myrank = __popc(bal & ((1u << lane) - 1u));
  }
  __syncthreads();
  int total = 0, q = 0;
  for (int k = 0; k < ns; k++) {
    int before = 0, tk = 0;
#pragma unroll
    for (int ww = 0; ww < NW; ww++) {
      const int c = s.wcount[ww * SOAP_SPECIES_CAP + k];
      tk += c;
      if (ww < w) before += c;
    }
    if (spc == k) q = total + before + myrank;
    if (threadIdx.x == 0) s.seg[k] = total;
    total += tk;
  }
  if (threadIdx.x == 0) s.seg[ns] = total;
  if (valid) {
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
    s.nbp[q] = p;
  }
  __syncthreads();
  return total;
}

// ------------------------------------------------------------------------------------------------
// Register-resident density expansion (specialised n_max, l_max): one warp, 4 neighbours q of one species per call.
//     Xt_lm(a) += sum_q Y_lm(q) Phi_l(a; q)                                              (descriptors.f95:8289-8295)
// as DMMA tiles C[m][a] += A[m][q] B[q][a] per level l, with the operands produced where the tensor core wants them:
// lane (fr, fk) runs the harmonic recursion in l for (order |m| = 8g + fr, neighbour fk) -> A elements of the cos tile P_g
// and the sin tile N_g, and the radial recursion for (neighbour fk, basis point a = 8 nt + fr) -> B elements.  The
// accumulators of ALL (l, tile, nt) stay in registers across the warp's neighbours and are flushed once per centre.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fwd_tiles(int L) { return 2 * L + 1 + (L >= 8 ? 2 * (L - 7) : 0); }
__host__ __device__ constexpr int fwd_tile_index(int L, int l, int gi, int sgn) {
  return gi == 0 ? (sgn == 0 ? l : L + l) : 2 * L + 1 + 2 * (l - 8) + sgn;
}

template <int CN, int CL>
__device__ __forceinline__ void forward_group(const Smem& s, const double alpha, const int q, const bool vq, const int fr, const int fk,
                                              double (&acc)[fwd_tiles(CL)][(CN + 7) / 8][2]) {
  // q = this lane's neighbour (buffer index); vq false = K padding: the lane is given a valid neighbour of the group and f = 0
  constexpr int NTN = (CN + 7) / 8, NG = CL >= 8 ? 2 : 1;
  const double r = s.nbr[q], rinv = 1.0 / r;
  const double f = vq ? s.nbf[q] : 0.0;
  const double ux = s.nbd[3 * q] * rinv, uy = s.nbd[3 * q + 1] * rinv, uz = s.nbd[3 * q + 2] * rinv;
  const double tar = 2.0 * alpha * r;
  double bl[NTN], blp[NTN], inv[NTN];
  bool za[NTN];
#pragma unroll
  for (int nt = 0; nt < NTN; nt++) {
    const int a = 8 * nt + fr;
    bl[nt] = blp[nt] = inv[nt] = 0.0;
    za[nt] = true;
    if (a < CN) {
      const double rb = s.rb[a], arg = tar * rb;
      if (arg == 0.0) bl[nt] = exp(-alpha * (rb * rb + r * r));
      else {
        const double exp_p = exp(-alpha * (r + rb) * (r + rb)), exp_m = exp(-alpha * (r - rb) * (r - rb));
        const double iv = 1.0 / arg, blm = 0.5 * (exp_m + exp_p) * iv, b0 = 0.5 * (exp_m - exp_p) * iv;
        inv[nt] = iv;
        bl[nt] = b0;
        blp[nt] = blm - b0 * iv;
        za[nt] = false;
      }
    }
  }
  double Cm[NG], Sm[NG], p1[NG], p2[NG];
  {
    double ca = 1.0, sa = 0.0;
    for (int k = 0; k < fr; k++) {
      const double t = ca;
      ca = ux * t - uy * sa;
      sa = ux * sa + uy * t;
    }
    Cm[0] = ca; Sm[0] = sa;
    if constexpr (NG == 2) {
      const double c2 = ux * ux - uy * uy, s2 = 2.0 * ux * uy, c4 = c2 * c2 - s2 * s2, s4 = 2.0 * c2 * s2, c8 = c4 * c4 - s4 * s4, s8 = 2.0 * c4 * s4;
      Cm[NG - 1] = c8 * ca - s8 * sa; Sm[NG - 1] = c8 * sa + s8 * ca;
    }
#pragma unroll
    for (int gi = 0; gi < NG; gi++) { p1[gi] = s.dblf[8 * gi + fr]; p2[gi] = 0.0; }
  }
#pragma unroll
  for (int l = 0; l <= CL; l++) {
    double phi[NTN];
#pragma unroll
    for (int nt = 0; nt < NTN; nt++) {
      if (l > 0) {
        const double blm = bl[nt], b = blp[nt];
        bl[nt] = za[nt] ? 0.0 : b;
        blp[nt] = za[nt] ? 0.0 : blm - (double)(2 * l + 1) * b * inv[nt];
      }
      phi[nt] = f * bl[nt];
    }
    const double tz = (double)(2 * l - 1) * uz;
#pragma unroll
    for (int gi = 0; gi < NG; gi++) {
      if (8 * gi <= l) {
        const int mv = 8 * gi + fr;
        double yc = 0.0, ys = 0.0;
        if (mv <= l) {
          double pl;
          if (l == mv) pl = p1[gi];
          else {
            pl = (tz * p1[gi] - (double)(l + mv - 1) * p2[gi]) * s.invint[l - mv];
            p2[gi] = p1[gi]; p1[gi] = pl;
          }
          const double qv = pl * s.ynorm[l * (l + 1) / 2 + mv];
          yc = qv * Cm[gi];
          ys = qv * Sm[gi];
        }
#pragma unroll
        for (int nt = 0; nt < NTN; nt++) dmma(acc[fwd_tile_index(CL, l, gi, 0)][nt][0], acc[fwd_tile_index(CL, l, gi, 0)][nt][1], yc, phi[nt]);
        if (l >= 1) {
#pragma unroll
          for (int nt = 0; nt < NTN; nt++) dmma(acc[fwd_tile_index(CL, l, gi, 1)][nt][0], acc[fwd_tile_index(CL, l, gi, 1)][nt][1], ys, phi[nt]);
        }
      }
    }
  }
}

// adds a warp's accumulators into Xt (rows lm, columns c0 + a); invalid rows (|m| > l) and padding channels are skipped
template <int CN, int CL>
__device__ __forceinline__ void forward_flush(double* X, const int XS, const int c0, const int fr, const int fk,
                                              const double (&acc)[fwd_tiles(CL)][(CN + 7) / 8][2]) {
  constexpr int NTN = (CN + 7) / 8, NG = CL >= 8 ? 2 : 1;
#pragma unroll
  for (int l = 0; l <= CL; l++)
#pragma unroll
    for (int gi = 0; gi < NG; gi++)
      if (8 * gi <= l) {
        const int mrow = 8 * gi + fr;
        if (mrow <= l) {
#pragma unroll
          for (int sgn = 0; sgn < 2; sgn++) {
            if (sgn == 1 && (l == 0 || mrow == 0)) continue;
            double* xr = X + (l * l + l + (sgn ? -mrow : mrow)) * XS + c0 + 2 * fk;
#pragma unroll
            for (int nt = 0; nt < NTN; nt++)
#pragma unroll
              for (int j = 0; j < 2; j++)
                if (8 * nt + 2 * fk + j < CN) xr[8 * nt + j] += acc[fwd_tile_index(CL, l, gi, sgn)][nt][j];
          }
        }
      }
}

// ------------------------------------------------------------------------------------------------
// forward: x (normalised power spectrum), X_lm (kept for the adjoint), |p|
// ------------------------------------------------------------------------------------------------
// CN / CL / CNS: compile-time n_max / l_max / n_species (0 = read them from the model: generic instantiation).  With
// constants every table stride, loop bound and index division folds at compile time.
template <int CN, int CL, int CNS>
__global__ void __launch_bounds__(NT, 4) k_soap_forward(const SoapDev* __restrict__ sp, const int* __restrict__ centres,
                                                        const int* __restrict__ n_centres_dev,
                                                        const int* __restrict__ nbr_off, const int* __restrict__ nbr_end, const int* __restrict__ nbr_j,
                                                        const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                        const int* __restrict__ Z, Lattice9 lat, double* __restrict__ x,
                                                        double* __restrict__ xlm, double* __restrict__ pnorm, int skip_power) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) return;  // the grid is sized by an upper bound; the centre count never leaves the device
  const Geo g = make_geo(CN ? CN : sp->n_max, CN ? CL : sp->l_max, CN ? CNS : sp->n_species);
  const int n = g.n, L = g.L, L1 = g.L1, nlm = g.nlm, K1 = g.K1, d = sp->d, d_pad = sp->d_pad, ns = g.ns;
  Smem s;
  carve(g, d_pad, false, &s, smem_raw);
  const int i = centres[c];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;  // DMMA fragment coordinates
  const double alpha = sp->alpha;
  load_tables(sp, g, s);
  for (int k = threadIdx.x; k < g.XR * g.XS; k += NT) s.X[k] = 0.0;
  __syncthreads();

  const int pbeg = nbr_off[i], pend = nbr_end[i];
  for (int pb = pbeg; pb < pend; pb += NBCAP) {
    int nv = gather_neighbours(sp, s, ns, i, pb, min(pb + NBCAP, pend), nbr_j, nbr_s, pos, Z, lat);
    for (int t0 = 0; t0 < nv; t0 += TNF) {
      const int tn = min(TNF, nv - t0), tn4 = ceil4(tn);
      // ---- stage: radial items (q, a) and harmonic items (q, m); rows q in [tn, tn4) are zero (K padding) ----
      const int n_rad = tn * n, n_items = n_rad + tn * g.MP;
      for (int it = threadIdx.x; it < n_items; it += NT) {
        if (it < n_rad) {
          int q = it / n, a = it - q * n;
          radial_item<false>(alpha, s.nbr[t0 + q], s.rb[a], s.nbf[t0 + q], 0.0, L, s.rf + q * g.RFS + a, nullptr, g.n2);
        } else {
          int r2 = it - n_rad;
          int m = r2 / tn, q = r2 - m * tn;
          double rinv = 1.0 / s.nbr[t0 + q];
          ylm_item<false>(s.ynorm, L, m, s.nbd[3 * (t0 + q)] * rinv, s.nbd[3 * (t0 + q) + 1] * rinv, s.nbd[3 * (t0 + q) + 2] * rinv,
                          s.Y + q * g.YS, nullptr, 0);
        }
      }
      for (int it = threadIdx.x; it < (tn4 - tn) * (g.RFS + g.YS); it += NT) {
        int q = tn + it / (g.RFS + g.YS), k = it % (g.RFS + g.YS);
        if (k < g.RFS) s.rf[q * g.RFS + k] = 0.0;
        else s.Y[q * g.YS + (k - g.RFS)] = 0.0;
      }
      __syncthreads();
      // ---- density expansion on the tensor cores: Xt[lm][c] += sum_q Y[q][lm] * Phi[q][l][a(c)] [species(q) == s(c)]
      //      (descriptors.f95:8289-8295, before the basis transform); one warp per (lm tile, channel tile) ----
      for (int t = warp; t < g.NTM * g.NTN; t += NW) {
        const int mt = t / g.NTN, nt = t - mt * g.NTN;
        const int lm0 = s.mt_lm0[mt], l = s.mt_l[mt];
        const int ch = nt * 8 + fr, cs = s.col_s[ch], ca = s.col_a[ch];  // this thread's B column
        const double* ap = s.Y + fk * g.YS + lm0 + fr;
        const double* bp = s.rf + fk * g.RFS + l * g.n2 + ca;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < tn4; k0 += 4) {
          double a = ap[k0 * g.YS];
          // padding rows q >= tn are zero; padding channels (cs < 0) only feed padding columns of X, which nothing reads
          double b = (ns == 1 || s.nbs[t0 + k0 + fk] == cs) ? bp[k0 * g.RFS] : 0.0;
          dmma(c0, c1, a, b);
        }
        const int lm = lm0 + fr;
        if (lm < (l + 1) * (l + 1)) {
          double* xr = s.X + lm * g.XS + nt * 8 + 2 * fk;
          xr[0] += c0;
          xr[1] += c1;
        }
      }
      __syncthreads();
    }
  }
  // ---- basis transform on the tensor cores, per species: X[lm][s,a'] = sum_a Xt[lm][s,a] T(a,a').  A warp owns 8 rows
  //      (all species of them): it reads a species block completely before overwriting it ----
  {
    const int n_mt = (nlm + 7) / 8, n_nt = g.n8 / 8, n_ks = ceil4(n) / 4;
    for (int mt = warp; mt < n_mt; mt += NW) {
      for (int sk = 0; sk < ns; sk++) {
        double acc[2][2] = {{0, 0}, {0, 0}};  // n8 <= 16: at most two output tiles
        for (int ks = 0; ks < n_ks; ks++) {
          const double a = s.X[(mt * 8 + fr) * g.XS + sk * n + ks * 4 + fk];  // columns >= n of the block meet zero rows of Tp
#pragma unroll
          for (int nt = 0; nt < 2; nt++)
            if (nt < n_nt) dmma(acc[nt][0], acc[nt][1], a, s.Tp[(ks * 4 + fk) * g.TS + nt * 8 + fr]);
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
          for (int j = 0; j < 2; j++) {
            int ap = nt * 8 + 2 * fk + j;
            if (nt < n_nt && ap < n) s.X[(mt * 8 + fr) * g.XS + sk * n + ap] = acc[nt][j];
          }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  // central atom term (descriptors.f95:8151-8182): only a = 1 is non-zero because the Cholesky factor is lower triangular
  if (threadIdx.x < ns) {
    int k = threadIdx.x;
    if (sp->cras || sp->species_Z[k] == Z[i] || sp->species_Z[k] == 0) s.X[k * n] += sp->central_weight * sp->chol00 * 0.28209479177387814347;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nlm * K1; k += NT) {
    int lm = k / K1, ic = k - lm * K1;
    xlm[(size_t)c * nlm * K1 + k] = s.X[lm * g.XS + ic];
  }
  if (skip_power) return;  // compression modes: the power spectrum of the mixed channels is taken by soap_general.cu from xlm
  // ---- power spectrum on the tensor cores (descriptors.f95:8370-8418): p_l(ia,jb) = sum_m X_lm(ia) X_lm(jb) / sqrt(2l+1),
  //      element q = l + (l_max+1) * pair(ia, jb<=ia), off-diagonal pairs times sqrt(2); one warp per (l, tile pair) ----
  double loc = 0.0;
  {
    const int n_pairs = g.NTN * (g.NTN + 1) / 2;
    for (int t = warp; t < L1 * n_pairs; t += NW) {
      const int l = t / n_pairs;
      int pr = t - l * n_pairs, ti = 0;
      while ((ti + 1) * (ti + 2) / 2 <= pr) ti++;
      const int tj = pr - ti * (ti + 1) / 2;
      const int lm_beg = l * l, lm_end = (l + 1) * (l + 1);
      double c0 = 0.0, c1 = 0.0;
      for (int k0 = lm_beg; k0 < lm_end; k0 += 4) {
        const int lm = k0 + fk;
        const double a = lm < lm_end ? s.X[lm * g.XS + ti * 8 + fr] : 0.0;
        const double b = lm < lm_end ? s.X[lm * g.XS + tj * 8 + fr] : 0.0;
        dmma(c0, c1, a, b);
      }
      const int ia = ti * 8 + fr;
      const double scale = sp->tlpo[l];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int jb = tj * 8 + 2 * fk + j;
        if (ia < K1 && jb <= ia) {
          double v = (j ? c1 : c0) * scale;
          if (ia != jb) v *= 1.41421356237309504880;
          s.p[l + L1 * (ia * (ia + 1) / 2 + jb)] = v;
          loc += v * v;
        }
      }
    }
  }
  double nrm = sqrt(block_sum(loc, s.red));  // :8450-8451 (block_sum synchronises: p is complete afterwards)
  double inv = sp->normalise ? 1.0 / nrm : 1.0;
  double* xr = x + (size_t)c * d_pad;
  for (int q = threadIdx.x; q < d_pad; q += NT) xr[q] = q < d - 1 ? s.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (threadIdx.x == 0) pnorm[c] = nrm;
}

// ------------------------------------------------------------------------------------------------
// WARP-PER-CENTRE kernels (specialised n_max, l_max, n_species): a warp owns a centre from the neighbour gather to the
// final store, so nothing in the centre's life needs a block barrier; the four warps of a CTA only share the read-only
// tables.  Per-warp shared memory: X (nlm x K1, padded), a compacted neighbour buffer of NBW entries (reused for the
// power spectrum / dE/dp vector once the neighbours are consumed).
// ------------------------------------------------------------------------------------------------
constexpr int NBW = 64;  // compacted neighbours buffered per warp; longer rows are worked off in several passes

struct WSmem {
  double *Tp, *rb, *ynorm, *invint, *dblf;  // block-shared tables
  double *X, *acc, *nbd, *nbr, *nbf, *nbdf, *p;  // per warp
  int *nbj, *nbs, *ord, *nbp;
};
constexpr int NBWA = 32;  // adjoint: every 32-entry chunk of the neighbour row is worked off at once
__host__ __device__ __forceinline__ size_t carve_w(const Geo& g, int d_pad, bool adjoint, int warp, WSmem* w, unsigned char* base) {
  const int NBW = adjoint ? NBWA : gapb200::NBW;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o += ((cnt + 1) & ~(size_t)1) * sizeof(double); return r; };
  const size_t oT = take((size_t)g.n8 * g.TS), orb = take(g.n), oyn = take((size_t)g.L1 * (g.L1 + 1) / 2), oinv = take(16), odbl = take(20);
  const size_t tables = o;
  o = 0;
  // adjoint: 8 x 12 centre force / virial slots, then an 8 x XS scratch tile (Lambda between its two tensor-core products)
  const size_t oX = take((size_t)g.XR * g.XS), oacc = take(adjoint ? 96 + 8 * g.XS : 0);
  const size_t oreg = o;
  const size_t onbd = take(3 * NBW), onbr = take(NBW), onbf = take(NBW), onbdf = take(adjoint ? NBW : 0);
  const size_t oint = o;
  o += sizeof(int) * 4 * NBW;
  if (o - oreg < (size_t)d_pad * sizeof(double)) o = oreg + (size_t)d_pad * sizeof(double);
  o = (o + 15) & ~(size_t)15;
  const size_t per_warp = o;
  if (w) {
    w->Tp = (double*)(base + oT); w->rb = (double*)(base + orb); w->ynorm = (double*)(base + oyn); w->invint = (double*)(base + oinv);
    w->dblf = (double*)(base + odbl);
    unsigned char* wb = base + tables + (size_t)warp * per_warp;
    w->X = (double*)(wb + oX); w->acc = (double*)(wb + oacc); w->nbd = (double*)(wb + onbd); w->nbr = (double*)(wb + onbr);
    w->nbf = (double*)(wb + onbf); w->nbdf = (double*)(wb + onbdf); w->p = (double*)(wb + oreg);
    w->nbj = (int*)(wb + oint); w->nbs = w->nbj + NBW; w->ord = w->nbs + NBW; w->nbp = w->ord + NBW;
  }
  return tables + (size_t)NW * per_warp;
}
__device__ __forceinline__ void load_tables_w(const SoapDev* sp, const Geo& g, const WSmem& w) {
  for (int k = threadIdx.x; k < g.n8 * g.TS; k += NT) {
    int a = k / g.TS, b = k - a * g.TS;
    w.Tp[k] = (a < g.n && b < g.n) ? sp->T[a + g.n * b] : 0.0;
  }
  for (int k = threadIdx.x; k < g.n; k += NT) w.rb[k] = sp->r_basis[k];
  for (int k = threadIdx.x; k < g.L1 * (g.L1 + 1) / 2; k += NT) w.ynorm[k] = sp->ynorm[k];
  if (threadIdx.x < 16) w.invint[threadIdx.x] = threadIdx.x <= LC + 1 ? c_invint[threadIdx.x] : 0.0;
  if (threadIdx.x < 20) w.dblf[threadIdx.x] = threadIdx.x <= LC ? c_dblfact[threadIdx.x] : 0.0;
}

// 32 CSR entries of centre i -> appended (ordered) to the warp's compacted buffer; returns the new fill
template <bool ADJ>
__device__ __forceinline__ int gather_chunk_w(const SoapDev* sp, const WSmem& w, int fill, int i, int p0, int pend, const int* __restrict__ nbr_j,
                                              const int* __restrict__ nbr_s, const double* __restrict__ pos, const int* __restrict__ Z,
                                              const Lattice9& lat, int lane) {
  const int p = p0 + lane;
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
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const int q = fill + __popc(bal & ((1u << lane) - 1u));
    double f, df;
    cutoff_fn(sp, r, f, df);
    w.nbd[3 * q] = dd[0];
    w.nbd[3 * q + 1] = dd[1];
    w.nbd[3 * q + 2] = dd[2];
    w.nbr[q] = r;
    w.nbf[q] = f;
    if (ADJ) { w.nbdf[q] = df; w.nbj[q] = j; w.nbp[q] = p; }
    w.nbs[q] = spc;
  }
  return fill + __popc(bal);
}
// stable partition of the buffer by species: ord[seg[k] .. seg[k+1]) lists the entries of species k
template <int CNS>
__device__ __forceinline__ void sort_species_w(const WSmem& w, int fill, int lane, int (&seg)[CNS + 1]) {
  int base = 0;
#pragma unroll
  for (int sk = 0; sk < CNS; sk++) {
    seg[sk] = base;
    for (int b0 = 0; b0 < fill; b0 += 32) {
      const int idx = b0 + lane;
      const bool mine = idx < fill && (CNS == 1 || w.nbs[idx] == sk);
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      if (mine) w.ord[base + __popc(bal & ((1u << lane) - 1u))] = idx;
      base += __popc(bal);
    }
  }
  seg[CNS] = base;
  __syncwarp();
}

// the Smem view forward_group / adjoint_tile read (neighbour arrays + tables) for a warp-owned buffer
__device__ __forceinline__ Smem view_of(const WSmem& w) {
  Smem s;
  s.Tp = w.Tp; s.rb = w.rb; s.ynorm = w.ynorm; s.invint = w.invint; s.dblf = w.dblf; s.X = w.X; s.nbd = w.nbd; s.nbr = w.nbr; s.nbf = w.nbf;
  s.nbdf = w.nbdf; s.nbj = w.nbj; s.nbp = w.nbp; s.nbs = w.nbs; s.p = w.p; s.acc = w.acc;
  s.X2 = nullptr; s.red = nullptr; s.rf = nullptr; s.Y = nullptr; s.mt_lm0 = nullptr; s.mt_l = nullptr; s.col_s = nullptr; s.col_a = nullptr;
  s.wcount = nullptr; s.seg = nullptr;
  return s;
}

template <int CN, int CL, int CNS>
__global__ void __launch_bounds__(NT, (CN > 8 ? 2 : 4)) k_soap_forward_w(const SoapDev* __restrict__ sp, const int* __restrict__ centres,
                                                                         const int* __restrict__ n_centres_dev,
                                                                         const int* __restrict__ nbr_off, const int* __restrict__ nbr_end, const int* __restrict__ nbr_j,
                                                                         const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                                         const int* __restrict__ Z, Lattice9 lat, double* __restrict__ x,
                                                                         double* __restrict__ xlm, double* __restrict__ pnorm, int skip_power) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Geo g = make_geo(CN, CL, CNS);
  constexpr int n = CN, L1 = CL + 1, nlm = (CL + 1) * (CL + 1), K1 = CN * CNS, ns = CNS;
  constexpr int NTN = (CN + 7) / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int d = sp->d, d_pad = sp->d_pad;
  WSmem w;
  carve_w(g, d_pad, false, warp, &w, smem_raw);
  pdl_launch_dependents();
  load_tables_w(sp, g, w);  // model constants: independent of the preceding kernels
  pdl_wait();
  __syncthreads();  // the only block barrier: the tables
  const int c = blockIdx.x * NW + warp;
  if (c >= *n_centres_dev) return;
  const Smem s = view_of(w);
  const int i = centres[c];
  const double alpha = sp->alpha;
  for (int k = lane; k < g.XR * g.XS; k += 32) w.X[k] = 0.0;
  __syncwarp();

  // ---- density expansion: gather -> groups of 4 neighbours of one species -> register accumulators -> X ----
  double acc[fwd_tiles(CL)][NTN][2];
  auto zero_acc = [&]() {
#pragma unroll
    for (int t = 0; t < fwd_tiles(CL); t++)
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) acc[t][nt][0] = acc[t][nt][1] = 0.0;
  };
  zero_acc();
  const int pbeg = nbr_off[i], pend = nbr_end[i];
  int fill = 0;
  for (int p0 = pbeg; p0 < pend; p0 += 32) {
    fill = gather_chunk_w<false>(sp, w, fill, i, p0, pend, nbr_j, nbr_s, pos, Z, lat, lane);
    if (fill > NBW - 32 || p0 + 32 >= pend) {  // the buffer cannot take another chunk, or the row is finished: work it off
      __syncwarp();
      int seg[CNS + 1];
      if (CNS == 1) { seg[0] = 0; seg[CNS] = fill; }
      else sort_species_w<CNS>(w, fill, lane, seg);
#pragma unroll
      for (int sk = 0; sk < CNS; sk++) {
        const int cnt = seg[sk + 1] - seg[sk];
        for (int t0 = 0; t0 < cnt; t0 += 4) {
          const bool vq = fk < cnt - t0;
          const int k = seg[sk] + t0 + (vq ? fk : 0);
          forward_group<CN, CL>(s, alpha, CNS == 1 ? k : w.ord[k], vq, fr, fk, acc);
        }
        if (CNS > 1 && cnt > 0) {
          forward_flush<CN, CL>(w.X, g.XS, sk * CN, fr, fk, acc);
          zero_acc();
        }
      }
      fill = 0;
      __syncwarp();
    }
  }
  if (CNS == 1) forward_flush<CN, CL>(w.X, g.XS, 0, fr, fk, acc);
  __syncwarp();
  // ---- basis transform on the tensor cores, per species: X[lm][s,a'] = sum_a Xt[lm][s,a] T(a,a') ----
  {
    constexpr int n_mt = (nlm + 7) / 8, n_nt = (CN + 7) / 8, n_ks = (CN + 3) / 4;
    for (int mt = 0; mt < n_mt; mt++) {
#pragma unroll
      for (int sk = 0; sk < ns; sk++) {
        double ta[n_nt][2];
#pragma unroll
        for (int nt = 0; nt < n_nt; nt++) ta[nt][0] = ta[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < n_ks; ks++) {
          const double a = w.X[(mt * 8 + fr) * g.XS + sk * n + ks * 4 + fk];  // columns >= n of the block meet zero rows of Tp
#pragma unroll
          for (int nt = 0; nt < n_nt; nt++) dmma(ta[nt][0], ta[nt][1], a, w.Tp[(ks * 4 + fk) * g.TS + nt * 8 + fr]);
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < n_nt; nt++)
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const int ap = nt * 8 + 2 * fk + j;
            if (ap < n) w.X[(mt * 8 + fr) * g.XS + sk * n + ap] = ta[nt][j];
          }
        __syncwarp();
      }
    }
  }
  // central atom term (descriptors.f95:8151-8182): only a = 1 is non-zero because the Cholesky factor is lower triangular
  if (lane < ns) {
    if (sp->cras || sp->species_Z[lane] == Z[i] || sp->species_Z[lane] == 0) w.X[lane * n] += sp->central_weight * sp->chol00 * 0.28209479177387814347;
  }
  __syncwarp();
  for (int k = lane; k < nlm * K1; k += 32) {
    const int lm = k / K1, ic = k - lm * K1;
    xlm[(size_t)c * nlm * K1 + k] = w.X[lm * g.XS + ic];
  }
  if (skip_power) return;
  // ---- power spectrum on the tensor cores (descriptors.f95:8370-8418) ----
  double loc = 0.0;
  {
    constexpr int NTC = (K1 + 7) / 8;
#pragma unroll 1
    for (int l = 0; l < L1; l++) {
      const int lm_beg = l * l, lm_end = (l + 1) * (l + 1);
      const double scale = sp->tlpo[l];
#pragma unroll
      for (int ti = 0; ti < NTC; ti++)
#pragma unroll
        for (int tj = 0; tj <= ti; tj++) {
          double c0 = 0.0, c1 = 0.0;
          for (int k0 = lm_beg; k0 < lm_end; k0 += 4) {
            const int lm = k0 + fk;
            const double a = lm < lm_end ? w.X[lm * g.XS + ti * 8 + fr] : 0.0;
            const double b = lm < lm_end ? w.X[lm * g.XS + tj * 8 + fr] : 0.0;
            dmma(c0, c1, a, b);
          }
          const int ia = ti * 8 + fr;
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const int jb = tj * 8 + 2 * fk + j;
            if (ia < K1 && jb <= ia) {
              double v = (j ? c1 : c0) * scale;
              if (ia != jb) v *= 1.41421356237309504880;
              w.p[l + L1 * (ia * (ia + 1) / 2 + jb)] = v;
              loc += v * v;
            }
          }
        }
    }
  }
  const double nrm = sqrt(warp_sum(loc));  // :8450-8451
  __syncwarp();
  const double inv = sp->normalise ? 1.0 / nrm : 1.0;
  double* xr = x + (size_t)c * d_pad;
  for (int q = lane; q < d_pad; q += 32) xr[q] = q < d - 1 ? w.p[q] * inv : (q == d - 1 ? sp->sigma0 : 0.0);
  if (lane == 0) pnorm[c] = nrm;
}

// One tile of up to 8 neighbours (all of species sk) of the adjoint, run by ONE warp entirely in registers:
//     A_lm(q) = sum_a Lambda~_lm(sk,a) R_l(a; q),   B_lm(q) = sum_a Lambda~_lm(sk,a) Phi_l(a; q)        (FP64 tensor cores)
//     f_gp(q) = sum_lm [ A_lm Y_lm(q) u_q + B_lm grad Y_lm(q) ]                                          (IPModel_GAP.f95:479)
// as DMMA tiles C[q][m] = sum_a R[q][a] Lambda~[l,m][a], level by level in l.  The fragment layout of mma.m8n8k4 is
// exploited so that nothing is staged through shared memory:
//   * lane (fr, fk) holds A-operand element (row q = fr, k = a = 4 ks + fk): it runs the reference's upward recursion in l
//     (descriptors.f95:8218-8258) for ITS (neighbour, basis point) items and hands R_l / Phi_l to the tensor core as they appear;
//   * it receives C elements (row q = fr, columns m = 2 fk, 2 fk + 1 of the column tile): for its neighbour and those two
//     orders it runs the harmonic recursion in l (Y_{l,+-m}, grad Y_{l,+-m}; GradSphericalYCartesian_all,
//     angular_functions.f95:205-278) in step with the level loop and folds the products into four scalars.
// Column tiles at level l: P_g (m = 8g + column, cos type) and N_g (m = -(8g + column), sin type), g = 0 [, 1 if l >= 8];
// +m and -m of a thread share the Legendre recursion.  Four shuffles finish the sum over m; lane fk == 0 scatters.
template <int CN, int CL>
__device__ __forceinline__ void adjoint_tile(const Smem& s, const Geo& g, const double alpha, const int sk, const int q0, const int tn, const int fr,
                                             const int fk, const double e_scale, double* __restrict__ force, double* __restrict__ fpair,
                                             double* __restrict__ local_virial, double* accs, const int* ord = nullptr) {
  constexpr bool SPEC = CN != 0;
  constexpr int KS = SPEC ? (CN + 3) / 4 : (SOAP_NMAX_CAP + 3) / 4;   // DMMA k steps over the radial channels of one species
  constexpr int NG = SPEC ? (CL >= 8 ? 2 : 1) : 2;                    // groups of 8 orders |m|
  constexpr int NJ1 = SPEC ? (CL >= 9 ? 2 : (CL == 8 ? 1 : 0)) : 2;   // orders of the second group a thread can own
  constexpr int NMJ = 2 + NJ1;
  const int n = g.n, L = SPEC ? CL : g.L, XS = g.XS;
  const bool vq = fr < tn;
  // padding rows reuse the tile's first neighbour with f = f' = 0: their A fragments vanish; ord (may be NULL) lists the
  // buffer entries of the tile's species
  const int q = ord ? ord[q0 + (vq ? fr : 0)] : q0 + (vq ? fr : 0);
  const double r = s.nbr[q], rinv = 1.0 / r;
  const double dx = s.nbd[3 * q], dy = s.nbd[3 * q + 1], dz = s.nbd[3 * q + 2];
  const double f = vq ? s.nbf[q] : 0.0, df = vq ? s.nbdf[q] : 0.0;
  const double ux = dx * rinv, uy = dy * rinv, uz = dz * rinv;
  const double tar = 2.0 * alpha * r;

  // ---- radial items (q, a = 4 ks + fk): state of the recursion i_{l+1} = i_{l-1} - (2l+1) i_l / x ----
  double bl[KS], blp[KS], inv[KS], tarb[KS];
  bool za[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    const int a = 4 * ks + fk;
    bl[ks] = blp[ks] = inv[ks] = tarb[ks] = 0.0;
    za[ks] = true;
    if (a < n) {
      const double rb = s.rb[a], arg = tar * rb;
      tarb[ks] = 2.0 * alpha * rb;
      if (arg == 0.0) bl[ks] = exp(-alpha * (rb * rb + r * r));  // :8225-8240: i_0 = exp(-alpha r^2), i_{l>0} = 0
      else {
        const double exp_p = exp(-alpha * (r + rb) * (r + rb)), exp_m = exp(-alpha * (r - rb) * (r - rb));
        const double iv = 1.0 / arg, blm = 0.5 * (exp_m + exp_p) * iv, b0 = 0.5 * (exp_m - exp_p) * iv;
        inv[ks] = iv;
        bl[ks] = b0;
        blp[ks] = blm - b0 * iv;
        za[ks] = false;
      }
    }
  }
  // ---- harmonic state: powers (ux + i uy)^m and Legendre-derivative recursions of this thread's orders ----
  double Cm[NMJ], Sm[NMJ], cM[NMJ], sM[NMJ], p1[NMJ], p2[NMJ], z1[NMJ], z2[NMJ];
  {
    double ca = 1.0, sa = 0.0, cb = 0.0, sb = 0.0;  // P_k and P_{k-1}, P_k = (ux + i uy)^k
    for (int k = 0; k < 2 * fk; k++) {
      const double t = ca, u = sa;
      ca = ux * t - uy * u;
      sa = ux * u + uy * t;
      cb = t;
      sb = u;
    }
    const double m0 = (double)(2 * fk);
    Cm[0] = ca; Sm[0] = sa; cM[0] = m0 * cb; sM[0] = m0 * sb;
    const double c1 = ux * ca - uy * sa, s1 = ux * sa + uy * ca;
    Cm[1] = c1; Sm[1] = s1; cM[1] = (m0 + 1.0) * ca; sM[1] = (m0 + 1.0) * sa;
    if constexpr (NJ1 >= 1) {
      const double c2 = ux * ux - uy * uy, s2 = 2.0 * ux * uy, c4 = c2 * c2 - s2 * s2, s4 = 2.0 * c2 * s2, c8 = c4 * c4 - s4 * s4, s8 = 2.0 * c4 * s4;
      const double c6 = c4 * c2 - s4 * s2, s6 = c4 * s2 + s4 * c2, c7 = c6 * ux - s6 * uy, s7 = c6 * uy + s6 * ux;
      Cm[2] = c8 * ca - s8 * sa; Sm[2] = c8 * sa + s8 * ca;                                    // P_{8+2fk}
      cM[2] = (m0 + 8.0) * (c7 * ca - s7 * sa); sM[2] = (m0 + 8.0) * (c7 * sa + s7 * ca);      // (8+2fk) P_{7+2fk}
      if constexpr (NJ1 >= 2) {
        Cm[NMJ - 1] = c8 * c1 - s8 * s1; Sm[NMJ - 1] = c8 * s1 + s8 * c1;                      // P_{9+2fk}
        cM[NMJ - 1] = (m0 + 9.0) * Cm[2]; sM[NMJ - 1] = (m0 + 9.0) * Sm[2];
      }
    }
#pragma unroll
    for (int jj = 0; jj < NMJ; jj++) {
      const int mv = (jj < 2 ? 0 : 8) + 2 * fk + (jj & 1);
      p1[jj] = s.dblf[mv];  // Q_m^m = (2m-1)!!
      p2[jj] = z1[jj] = z2[jj] = 0.0;
    }
  }

  double SA = 0.0, G0 = 0.0, G1 = 0.0, G2 = 0.0;
  const double* Xc = s.X + sk * n + fk;  // Lambda~ columns of this species, this thread's k offset
#pragma unroll(SPEC ? 16 : 1)
  for (int l = 0; l <= L; l++) {
    double phi[KS], rr[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      if (l > 0) {
        const double blm = bl[ks], b = blp[ks];
        bl[ks] = za[ks] ? 0.0 : b;
        blp[ks] = za[ks] ? 0.0 : blm - (double)(2 * l + 1) * b * inv[ks];
      }
      phi[ks] = f * bl[ks];
      rr[ks] = f * (-tar * bl[ks] + (double)l * bl[ks] * rinv + blp[ks] * tarb[ks]) + df * bl[ks];  // :8254-8255 and f' Phi (:8261-8263)
    }
    const int lmc = l * l + l;
    const double tz = (double)(2 * l - 1) * uz;
#pragma unroll
    for (int gi = 0; gi < NG; gi++) {
      if (8 * gi <= l) {
        const int mcol = 8 * gi + fr;  // |m| of this thread's B-operand column
        const bool vP = mcol <= l, vN = vP && mcol >= 1;
        double aP[2] = {0.0, 0.0}, bP[2] = {0.0, 0.0}, aN[2] = {0.0, 0.0}, bN[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          const double xp = vP ? Xc[(lmc + mcol) * XS + 4 * ks] : 0.0;
          dmma(aP[0], aP[1], rr[ks], xp);
          dmma(bP[0], bP[1], phi[ks], xp);
        }
        if (l >= 1) {
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            const double xn = vN ? Xc[(lmc - mcol) * XS + 4 * ks] : 0.0;
            dmma(aN[0], aN[1], rr[ks], xn);
            dmma(bN[0], bN[1], phi[ks], xn);
          }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int jj = 2 * gi + j;
          if (jj < NMJ) {
            const int mv = 8 * gi + 2 * fk + j;
            if (mv <= l) {
              double pl, zl;
              if (l == mv) { pl = p1[jj]; zl = 0.0; }
              else {
                pl = (tz * p1[jj] - (double)(l + mv - 1) * p2[jj]) * s.invint[l - mv];
                zl = (l == mv + 1) ? s.dblf[mv + 1] : (tz * z1[jj] - (double)(l + mv) * z2[jj]) * s.invint[l - mv - 1];
                p2[jj] = p1[jj]; p1[jj] = pl;
                z2[jj] = z1[jj]; z1[jj] = zl;
              }
              const double nrm = s.ynorm[l * (l + 1) / 2 + mv];
              const double qv = pl * nrm, qz = zl * nrm;
              SA += qv * (aP[j] * Cm[jj] + aN[j] * Sm[jj]);
              G0 += qv * (bP[j] * cM[jj] + bN[j] * sM[jj]);
              G1 += qv * (bN[j] * cM[jj] - bP[j] * sM[jj]);
              G2 += qz * (bP[j] * Cm[jj] + bN[j] * Sm[jj]);
            }
          }
        }
      }
    }
  }
  // sum over the orders held by the four lanes of a fragment row
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    SA += __shfl_xor_sync(0xffffffffu, SA, o);
    G0 += __shfl_xor_sync(0xffffffffu, G0, o);
    G1 += __shfl_xor_sync(0xffffffffu, G1, o);
    G2 += __shfl_xor_sync(0xffffffffu, G2, o);
  }
  if (fk == 0 && vq) {
    const double ug = ux * G0 + uy * G1 + uz * G2;
    // f_gp,k = sum_lm [ A_lm Y_lm u_k + B_lm grad_k Y_lm ],  grad Y = (g - u (u.g)) / r
    const double f0 = (SA * ux + (G0 - ux * ug) * rinv) * e_scale;
    const double f1 = (SA * uy + (G1 - uy * ug) * rinv) * e_scale;
    const double f2 = (SA * uz + (G2 - uz * ug) * rinv) * e_scale;
    // IPModel_GAP.f95:479-491: F_j -= f_gp ; centre row is minus the sum ; W_j -= (pos_j - pos_i) (x) f_gp
    const int j = s.nbj[q];
    if (force) {
      if (fpair) {  // deterministic scatter: the pair force goes to the slot of its list entry, summed per atom in a fixed order afterwards
        const size_t pp = (size_t)s.nbp[q];
        fpair[3 * pp + 0] = -f0; fpair[3 * pp + 1] = -f1; fpair[3 * pp + 2] = -f2;
      } else {
        atomicAdd(&force[3 * (size_t)j + 0], -f0);
        atomicAdd(&force[3 * (size_t)j + 1], -f1);
        atomicAdd(&force[3 * (size_t)j + 2], -f2);
      }
      accs[0] += f0; accs[1] += f1; accs[2] += f2;
    }
    const double wv[9] = {dx * f0, dy * f0, dz * f0, dx * f1, dy * f1, dz * f1, dx * f2, dy * f2, dz * f2};  // column-major (a + 3b)
#pragma unroll
    for (int k = 0; k < 9; k++) accs[3 + k] -= wv[k];
    if (local_virial)
#pragma unroll
      for (int k = 0; k < 9; k++) atomicAdd(&local_virial[9 * (size_t)j + k], -wv[k]);
  }
}

// ------------------------------------------------------------------------------------------------
// adjoint: gvec = dE_i/dx  ->  forces / virial
// ------------------------------------------------------------------------------------------------
template <int CN, int CL, int CNS>
__global__ void __launch_bounds__(NT, 4) k_soap_adjoint(const SoapDev* __restrict__ sp, const int* __restrict__ centres,
                                                        const int* __restrict__ n_centres_dev,
                                                        const int* __restrict__ nbr_off, const int* __restrict__ nbr_end, const int* __restrict__ nbr_j,
                                                        const int* __restrict__ nbr_s, const double* __restrict__ pos,
                                                        const int* __restrict__ Z, Lattice9 lat, const double* __restrict__ x,
                                                        const double* __restrict__ xlm, const double* __restrict__ pnorm,
                                                        const double* __restrict__ gvec, int ldg, int g_splits, size_t g_split_stride,
                                                        const double* __restrict__ epart, int n_tiles_n, double* __restrict__ local_e,
                                                        double e_scale, double* __restrict__ force, double* __restrict__ vir_part,
                                                        double* __restrict__ local_virial, double* __restrict__ fpair,
                                                        const double* __restrict__ lambda_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  if (c >= *n_centres_dev) {
    if (vir_part && threadIdx.x < 9) vir_part[9 * (size_t)c + threadIdx.x] = 0.0;  // unused slot of the upper-bound grid
    return;
  }
  const int i = centres[c];
  // E_i = sum over the column tiles of GEMM-1 (fixed order) ; local_e(centre) += E_i  (IPModel_GAP.f95:454-459, cc = 1, |ci| = 1)
  if (epart && threadIdx.x < 32) {  // warp 0: the column-tile partials in parallel, shuffle tree (fixed order)
    double t = 0.0;
    for (int k = threadIdx.x; k < n_tiles_n; k += 32) t += epart[(size_t)c * n_tiles_n + k];
    t = warp_sum(t);
    if (threadIdx.x == 0) local_e[i] += e_scale * t;
  }
  const Geo g = make_geo(CN ? CN : sp->n_max, CN ? CL : sp->l_max, CN ? CNS : sp->n_species);
  const int n = g.n, L = g.L, L1 = g.L1, nlm = g.nlm, K1 = g.K1, d = sp->d, ns = g.ns;
  Smem s;
  carve(g, sp->d_pad, true, &s, smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const double alpha = sp->alpha;
  load_tables(sp, g, s);
  // u = dE/dp: pull gradPredict back through x = p/|p| (reference forward form: descriptors.f95:8595-8600)
  const double* xr = x + (size_t)c * sp->d_pad;
  const double* gr = gvec + (size_t)c * ldg;
  // lambda_in != NULL (compression modes): Lambda = dE/dX_lm [centre][lm][K1] was formed by soap_general.cu from the element list and the
  // mixing matrices; this kernel starts at the basis-transform pull-back
  // X_lm -> shared: asynchronous copies issued first so that their latency overlaps the gradPredict loads below
  if (!lambda_in) {
  for (int k = threadIdx.x; k < nlm * K1; k += NT) {
    int lm = k / K1, ic = k - lm * K1;
    cp_async8(s.X + lm * g.XS + ic, xlm + (size_t)c * nlm * K1 + k);
  }
  cp_async_commit();
  }
  // gradPredict arrives as g_splits partial sums (the K splits of GEMM-2), added here in a fixed order; four elements per
  // thread are in flight at a time
  double loc = 0.0;
  const double nrm = lambda_in ? 1.0 : pnorm[c];
  for (int q0 = 0; q0 < (lambda_in ? 0 : d - 1); q0 += 4 * NT) {
    double xv[4], gv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int q = q0 + u * NT + threadIdx.x;
      xv[u] = q < d - 1 ? xr[q] : 0.0;
      gv[u] = q < d - 1 ? gr[q] : 0.0;
    }
    for (int k = 1; k < g_splits; k++)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        int q = q0 + u * NT + threadIdx.x;
        if (q < d - 1) gv[u] += gr[(size_t)k * g_split_stride + q];
      }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int q = q0 + u * NT + threadIdx.x;
      if (q < d - 1) s.p[q] = gv[u];
      loc += xv[u] * gv[u];
    }
  }
  // zero padding of X (rows >= nlm, columns >= K1) and the Lambda buffer
  for (int k = threadIdx.x; k < g.XR * g.XS; k += NT) {
    int lm = k / g.XS, ic = k - lm * g.XS;
    if (lm >= nlm || ic >= K1) s.X[k] = 0.0;
    s.X2[k] = 0.0;
  }
  double sdot = block_sum(loc, s.red);
  if (sp->normalise && !lambda_in)
    for (int q = threadIdx.x; q < d - 1; q += NT) s.p[q] = (s.p[q] - xr[q] * sdot) / nrm;
  cp_async_wait_all();
  __syncthreads();
  if (lambda_in) {
    for (int k = threadIdx.x; k < nlm * K1; k += NT) {
      int lm = k / K1, ic = k - lm * K1;
      s.X2[lm * g.XS + ic] = lambda_in[(size_t)c * nlm * K1 + k];
    }
  }
  // ---- Lambda = dE/dX_lm on the tensor cores: Lambda[lm][ia] = sum_jb X[lm][jb] U~_l(jb, ia) / sqrt(2l+1), with the symmetric
  //      U~_l(ia,jb) = 2 u (ia == jb) or sqrt(2) u (ia != jb), u = dE/dp at (l, pair(ia, jb)) ----
  for (int t = warp; t < (lambda_in ? 0 : g.NTM * g.NTN); t += NW) {
    const int mt = t / g.NTN, nt = t - mt * g.NTN;
    const int lm0 = s.mt_lm0[mt], l = s.mt_l[mt];
    const int ia = nt * 8 + fr;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < g.K18; k0 += 4) {
      const int jb = k0 + fk;
      const double a = s.X[(lm0 + fr) * g.XS + jb];
      double b = 0.0;
      if (ia < K1 && jb < K1) {
        const int hi = ia > jb ? ia : jb, lo = ia > jb ? jb : ia;
        const double u = s.p[l + L1 * (hi * (hi + 1) / 2 + lo)];
        b = ia == jb ? 2.0 * u : 1.41421356237309504880 * u;
      }
      dmma(c0, c1, a, b);
    }
    const int lm = lm0 + fr;
    if (lm < (l + 1) * (l + 1)) {
      const double sc = sp->tlpo[l];
      double* lr = s.X2 + lm * g.XS + nt * 8 + 2 * fk;
      lr[0] = c0 * sc;
      lr[1] = c1 * sc;
    }
  }
  __syncthreads();
  // ---- Lambda~[lm][s,a] = sum_a' Lambda[lm][s,a'] T(a,a') : the basis transform pulled back once per centre (into s.X) ----
  {
    const int n_mt = (nlm + 7) / 8, n_nt = g.n8 / 8, n_ks = ceil4(n) / 4;
    for (int t = warp; t < n_mt * ns; t += NW) {
      const int mt = t / ns, sk = t - mt * ns;
      double acc[2][2] = {{0, 0}, {0, 0}};
      for (int ks = 0; ks < n_ks; ks++) {
        const int ap = ks * 4 + fk;  // k index = a'
        const double a = ap < n ? s.X2[(mt * 8 + fr) * g.XS + sk * n + ap] : 0.0;
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
          if (nt < n_nt) dmma(acc[nt][0], acc[nt][1], a, s.Tp[(nt * 8 + fr) * g.TS + ap]);  // B[k=a'][col=a] = T(a,a')
      }
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          int a = nt * 8 + 2 * fk + j;
          if (nt < n_nt && a < n) s.X[(mt * 8 + fr) * g.XS + sk * n + a] = acc[nt][j];
        }
    }
  }
  __syncthreads();

  // ---- neighbour phase: every warp works through tiles of 8 neighbours on its own (no block barrier, no staging) ----
  double* accs = s.acc + (warp * 8 + fr) * 12;  // centre force (3) and virial (9) partials of this (warp, fragment row), lane fk == 0
  if (fk == 0)
#pragma unroll
    for (int k = 0; k < 12; k++) accs[k] = 0.0;
  const int pbeg = nbr_off[i], pend = nbr_end[i];
  for (int pb = pbeg; pb < pend; pb += NBCAP) {
    gather_neighbours(sp, s, ns, i, pb, min(pb + NBCAP, pend), nbr_j, nbr_s, pos, Z, lat);
    int tbase = 0;
    for (int sk = 0; sk < ns; sk++) {
      const int sbeg = s.seg[sk], send = s.seg[sk + 1];
      const int ntile = (send - sbeg + 7) >> 3;
      for (int t = (warp + NW - (tbase & (NW - 1))) & (NW - 1); t < ntile; t += NW)
        adjoint_tile<CN, CL>(s, g, alpha, sk, sbeg + 8 * t, min(8, send - sbeg - 8 * t), fr, fk, e_scale, force, fpair, local_virial, accs);
      tbase += ntile;
    }
    if (pb + NBCAP < pend) __syncthreads();  // the next pass overwrites the compacted list
  }
  // centre force / virial: the 32 slots are added in a fixed order
  __syncthreads();
  if (threadIdx.x < 12) {
    const int k = threadIdx.x;
    double t = 0.0;
#pragma unroll 8
    for (int sl = 0; sl < NW * 8; sl++) t += s.acc[sl * 12 + k];
    if (k < 3) {
      if (force && fpair) force[3 * (size_t)i + k] = t;  // deterministic scatter: `force` is the per-atom buffer of the centres' own sums
      else if (force) atomicAdd(&force[3 * (size_t)i + k], t);
    } else if (vir_part) vir_part[9 * (size_t)c + (k - 3)] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// adjoint, one WARP per centre (specialised shapes): as k_soap_forward_w, a warp owns its centre from the first load to the
// last atomic, so nothing needs a block barrier (the four warps of a CTA only share the read-only tables) and the load
// latencies of sixteen independent centres per SM overlap.  Per-warp shared memory: X (-> Lambda~ in place, one l-aligned
// 8-row tile at a time through a scratch tile), the 8 x 12 centre force / virial slots, and ONE region that first holds
// u = dE/dp and then, when u is dead, the compacted neighbour chunk.
// ------------------------------------------------------------------------------------------------
template <int CN, int CL, int CNS>
__global__ void __launch_bounds__(NT, 4) k_soap_adjoint_w(const SoapDev* __restrict__ sp, const int* __restrict__ centres,
                                                          const int* __restrict__ n_centres_dev, int n_centres_ub,
                                                          const int* __restrict__ nbr_off, const int* __restrict__ nbr_end,
                                                          const int* __restrict__ nbr_j, const int* __restrict__ nbr_s,
                                                          const double* __restrict__ pos, const int* __restrict__ Z, Lattice9 lat,
                                                          const double* __restrict__ x, const double* __restrict__ xlm,
                                                          const double* __restrict__ pnorm, const double* __restrict__ gvec, int ldg, int g_splits,
                                                          size_t g_split_stride, const double* __restrict__ epart, int n_tiles_n,
                                                          double* __restrict__ local_e, double e_scale, double* __restrict__ force,
                                                          double* __restrict__ vir_part, double* __restrict__ local_virial, double* __restrict__ fpair,
                                                          const double* __restrict__ lambda_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Geo g = make_geo(CN, CL, CNS);
  constexpr int n = CN, L1 = CL + 1, nlm = (CL + 1) * (CL + 1), K1 = CN * CNS, ns = CNS;
  constexpr int K18 = (K1 + 7) & ~7, NTN = K18 / 8, n_nt = (CN + 7) / 8, n_ks = (CN + 3) / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int d = sp->d, d_pad = sp->d_pad;
  WSmem w;
  carve_w(g, d_pad, true, warp, &w, smem_raw);
  pdl_launch_dependents();
  load_tables_w(sp, g, w);  // model constants: independent of the preceding kernels
  pdl_wait();
  __syncthreads();  // the only block barrier: the tables
  const int c = blockIdx.x * NW + warp;
  if (c >= n_centres_ub) return;
  if (c >= *n_centres_dev) {
    if (vir_part && lane < 9) vir_part[9 * (size_t)c + lane] = 0.0;  // unused slot of the upper-bound grid
    return;
  }
  const int i = centres[c];
  const double alpha = sp->alpha;
  const int XS = g.XS;
  // E_i = sum over the column tiles of GEMM-1 (shuffle tree, fixed order) ; local_e(centre) += E_i  (IPModel_GAP.f95:454-459)
  if (epart) {
    double t = 0.0;
    for (int k = lane; k < n_tiles_n; k += 32) t += epart[(size_t)c * n_tiles_n + k];
    t = warp_sum(t);
    if (lane == 0) local_e[i] += e_scale * t;
  }
  // X_lm -> shared (asynchronous), zero padding; u = dE/dp from the K-split partials of gradPredict, pulled back through x = p/|p|
  if (!lambda_in) {
    for (int k = lane; k < nlm * K1; k += 32) {
      const int lm = k / K1, ic = k - lm * K1;
      cp_async8(w.X + lm * XS + ic, xlm + (size_t)c * nlm * K1 + k);
    }
    cp_async_commit();
  }
  for (int k = lane; k < g.XR * XS; k += 32) {
    const int lm = k / XS, ic = k - lm * XS;
    if (lm >= nlm || ic >= K1) w.X[k] = 0.0;
  }
  const double* xr = x + (size_t)c * d_pad;
  const double* gr = gvec + (size_t)c * ldg;
  const double nrm = lambda_in ? 1.0 : pnorm[c];
  double loc = 0.0;
  for (int q0 = 0; q0 < (lambda_in ? 0 : d - 1); q0 += 4 * 32) {
    double xv[4], gv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = q0 + u * 32 + lane;
      xv[u] = q < d - 1 ? xr[q] : 0.0;
      gv[u] = q < d - 1 ? gr[q] : 0.0;
    }
    for (int k = 1; k < g_splits; k++)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int q = q0 + u * 32 + lane;
        if (q < d - 1) gv[u] += gr[(size_t)k * g_split_stride + q];
      }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = q0 + u * 32 + lane;
      if (q < d - 1) w.p[q] = gv[u];
      loc += xv[u] * gv[u];
    }
  }
  const double sdot = warp_sum(loc);
  __syncwarp();
  if (sp->normalise && !lambda_in)
    for (int q = lane; q < d - 1; q += 32) w.p[q] = (w.p[q] - xr[q] * sdot) / nrm;
  cp_async_wait_all();
  __syncwarp();

  // ---- Lambda = dE/dX_lm and Lambda~ = Lambda . T^T on the tensor cores, one l-aligned tile of 8 lm rows at a time, in place:
  //      Lambda[lm][ia] = sum_jb X[lm][jb] U~_l(jb, ia) / sqrt(2l+1),  U~_l(ia,jb) = 2 u (ia == jb) or sqrt(2) u (ia != jb) ----
  double* const scr = w.acc + 96;  // 8 x XS
#pragma unroll 1
  for (int l = 0; l <= CL; l++) {
    const double sc = sp->tlpo[l];
    const int lm_end = (l + 1) * (l + 1);
#pragma unroll 1
    for (int lm0 = l * l; lm0 < lm_end; lm0 += 8) {
      if (lambda_in) {  // Lambda rows of this tile straight from the general path's pre-kernel (already scaled)
        for (int k = lane; k < 8 * K18; k += 32) {
          const int r = k / K18, col = k - r * K18;
          scr[r * XS + col] = (lm0 + r < lm_end && col < K1) ? lambda_in[((size_t)c * nlm + lm0 + r) * K1 + col] : 0.0;
        }
      } else {
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) {
        const int ia = nt * 8 + fr;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < K18; k0 += 4) {
          const int jb = k0 + fk;
          const double a = w.X[(lm0 + fr) * XS + jb];
          double b = 0.0;
          if (ia < K1 && jb < K1) {
            const int hi = ia > jb ? ia : jb, lo = ia > jb ? jb : ia;
            const double u = w.p[l + L1 * (hi * (hi + 1) / 2 + lo)];
            b = ia == jb ? 2.0 * u : 1.41421356237309504880 * u;
          }
          dmma(c0, c1, a, b);
        }
        scr[fr * XS + nt * 8 + 2 * fk] = c0 * sc;
        scr[fr * XS + nt * 8 + 2 * fk + 1] = c1 * sc;
      }
      }
      __syncwarp();
      const bool row_ok = lm0 + fr < lm_end;  // rows beyond this l belong to the next tile's l: left alone
#pragma unroll
      for (int sk = 0; sk < ns; sk++) {
        double ta[n_nt][2];
#pragma unroll
        for (int nt = 0; nt < n_nt; nt++) ta[nt][0] = ta[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < n_ks; ks++) {
          const int ap = ks * 4 + fk;  // k index = a'
          const double a = ap < n ? scr[fr * XS + sk * n + ap] : 0.0;
#pragma unroll
          for (int nt = 0; nt < n_nt; nt++) dmma(ta[nt][0], ta[nt][1], a, w.Tp[(nt * 8 + fr) * g.TS + ap]);  // B[k=a'][col=a] = T(a,a')
        }
#pragma unroll
        for (int nt = 0; nt < n_nt; nt++)
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const int a = nt * 8 + 2 * fk + j;
            if (a < n && row_ok) w.X[(lm0 + fr) * XS + sk * n + a] = ta[nt][j];
          }
      }
      __syncwarp();
    }
  }

  // ---- neighbour phase: the u region now takes the compacted neighbours, 32 CSR entries at a time ----
  double* accs = w.acc + fr * 12;  // centre force (3) and virial (9) partials of fragment row fr, lane fk == 0
  if (fk == 0)
#pragma unroll
    for (int k = 0; k < 12; k++) accs[k] = 0.0;
  __syncwarp();
  const Smem s = view_of(w);
  const int pbeg = nbr_off[i], pend = nbr_end[i];
  for (int p0 = pbeg; p0 < pend; p0 += 32) {
    const int fill = gather_chunk_w<true>(sp, w, 0, i, p0, pend, nbr_j, nbr_s, pos, Z, lat, lane);
    __syncwarp();
    int seg[CNS + 1];
    if (CNS == 1) { seg[0] = 0; seg[CNS] = fill; }
    else sort_species_w<CNS>(w, fill, lane, seg);
#pragma unroll
    for (int sk = 0; sk < CNS; sk++) {
      const int cnt = seg[sk + 1] - seg[sk];
      for (int t0 = 0; t0 < cnt; t0 += 8)
        adjoint_tile<CN, CL>(s, g, alpha, sk, seg[sk] + t0, min(8, cnt - t0), fr, fk, e_scale, force, fpair, local_virial, accs, CNS == 1 ? nullptr : w.ord);
    }
    __syncwarp();
  }
  // centre force / virial: the 8 slots are added in a fixed order
  __syncwarp();
  if (lane < 12) {
    double t = 0.0;
#pragma unroll
    for (int sl = 0; sl < 8; sl++) t += w.acc[sl * 12 + lane];
    if (lane < 3) {
      if (force && fpair) force[3 * (size_t)i + lane] = t;
      else if (force) atomicAdd(&force[3 * (size_t)i + lane], t);
    } else if (vir_part) vir_part[9 * (size_t)c + (lane - 3)] = t;
  }
}

__global__ void k_select_centres(const int* __restrict__ Z, int first, int last, const SoapDev* __restrict__ sp, int* __restrict__ flags) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > last - first) return;
  int f = 0;
  if (t < last - first) {
    int Zi = Z[first + t];
    for (int k = 0; k < sp->n_Z; k++)
      if (Zi >= 0 && (sp->centre_Z[k] == Zi || sp->centre_Z[k] == 0)) f = 1;  // descriptors.f95:7962 ; Zi < 0: outside the atom mask
  }
  flags[t] = f;
}
__global__ void k_compact(const int* __restrict__ scan, const int* __restrict__ flags, int first, int n, int* __restrict__ centres) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && flags[t]) centres[scan[t]] = first + t;
}

}  // namespace

size_t soap_forward_smem(const SoapDev& h) { return carve(make_geo(h.n_max, h.l_max, h.n_species), h.d_pad, false, nullptr, nullptr); }
size_t soap_adjoint_smem(const SoapDev& h) { return carve(make_geo(h.n_max, h.l_max, h.n_species), h.d_pad, true, nullptr, nullptr); }

// (n_max, l_max, n_species) combinations with a fully specialised instantiation; everything else runs the generic one
#define SOAP_SPECIALISATIONS(X) X(8, 8, 1) X(12, 8, 1) X(10, 6, 2) X(12, 6, 1) X(10, 12, 1)
#define SOAP_ADJOINT_W(X) X(8, 8, 1) X(12, 6, 1)  // (12,8,1) and (10,6,2): per-warp state leaves < 12 warps per SM, measured slower than the block kernel

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

void launch_soap_forward(const SoapDev* sp, const SoapDev& h, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off, const int* nbr_end,
                         const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, double* x, double* xlm, double* pnorm,
                         cudaStream_t st, int* launches, int skip_power) {
  const int n_centres = n_centres_ub;
  if (n_centres <= 0) return;
  size_t sm = soap_forward_smem(h);
  *launches += 1;
#define GO(N, L, S)                                                                                                       \
  if (h.n_max == N && h.l_max == L && h.n_species == S) {                                                                  \
    const size_t smw = carve_w(make_geo(N, L, S), h.d_pad, false, 0, nullptr, nullptr);                                    \
    cudaFuncSetAttribute(k_soap_forward_w<N, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw);                \
    launch_pdl(k_soap_forward_w<N, L, S>, dim3((n_centres + NW - 1) / NW), dim3(NT), smw, st, sp, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, \
               pos, Z, lat, x, xlm, pnorm, skip_power);                                                                    \
    return;                                                                                                                \
  }
  SOAP_SPECIALISATIONS(GO)
#undef GO
  cudaFuncSetAttribute(k_soap_forward<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  launch_pdl(k_soap_forward<0, 0, 0>, dim3(n_centres), dim3(NT), sm, st, sp, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm,
             skip_power);
}

void launch_soap_adjoint(const SoapDev* sp, const SoapDev& h, const int* centres, const int* n_centres_dev, int n_centres_ub, const int* nbr_off, const int* nbr_end,
                         const int* nbr_j, const int* nbr_s, const double* pos, const int* Z, Lattice9 lat, const double* x, const double* xlm,
                         const double* pnorm, const double* gvec, int ldg, int g_splits, size_t g_split_stride, const double* epart, int n_tiles_n,
                         double* local_e, double e_scale, double* force, double* vir_part, double* local_virial, double* fpair, cudaStream_t st,
                         int* launches, const double* lambda_in) {
  const int n_centres = n_centres_ub;
  if (n_centres <= 0) return;
  size_t sm = soap_adjoint_smem(h);
  *launches += 1;
  // shapes whose per-warp state leaves room for 16 warps per SM run the warp-per-centre kernel
#define GOW(N, L, S)                                                                                                                         \
  if (h.n_max == N && h.l_max == L && h.n_species == S) {                                                                                     \
    const size_t smw = carve_w(make_geo(N, L, S), h.d_pad, true, 0, nullptr, nullptr);                                                       \
    cudaFuncSetAttribute(k_soap_adjoint_w<N, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw);                                   \
    launch_pdl(k_soap_adjoint_w<N, L, S>, dim3((n_centres + NW - 1) / NW), dim3(NT), smw, st, sp, centres, n_centres_dev, n_centres, nbr_off, nbr_end,  \
               nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm, gvec, ldg, g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part,  \
               local_virial, fpair, lambda_in);                                                                                              \
    return;                                                                                                                                   \
  }
  SOAP_ADJOINT_W(GOW)
#undef GOW
#define GO(N, L, S)                                                                                                                          \
  if (h.n_max == N && h.l_max == L && h.n_species == S) {                                                                                     \
    cudaFuncSetAttribute(k_soap_adjoint<N, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                                      \
    launch_pdl(k_soap_adjoint<N, L, S>, dim3(n_centres), dim3(NT), sm, st, sp, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, \
               pnorm, gvec, ldg, g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, fpair, lambda_in); \
    return;                                                                                                                                   \
  }
  SOAP_SPECIALISATIONS(GO)
#undef GO
  cudaFuncSetAttribute(k_soap_adjoint<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  launch_pdl(k_soap_adjoint<0, 0, 0>, dim3(n_centres), dim3(NT), sm, st, sp, centres, n_centres_dev, nbr_off, nbr_end, nbr_j, nbr_s, pos, Z, lat, x, xlm, pnorm,
             gvec, ldg, g_splits, g_split_stride, epart, n_tiles_n, local_e, e_scale, force, vir_part, local_virial, fpair, lambda_in);
}

}  // namespace gapb200
