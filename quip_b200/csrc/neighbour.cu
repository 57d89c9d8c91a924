// neighbour.cu -- device neighbour list (replaces the serial connection_calc_connect,
// src/libAtoms/Connection.f95:1035-1310, and the accessor connection_neighbour :2011-2138).
//
// Design (B200-first, not the reference's linked lists): atoms are binned into a cell grid with
// cell width >= cutoff, sorted by cell with a stable radix sort (order inside a cell = ascending atom
// index, so everything downstream is deterministic; the cell ranges are read off the sorted keys, no
// counting pass), and one WARP per atom -- lanes sweep the candidates of the (2R+1)^3 surrounding
// cells 32 at a time -- writes the FULL list (both directions; the reference stores a half list plus
// back references) in one of two row layouts:
//   exact       count, exclusive scan, fill -> packed CSR rows (first call for a geometry, stage-level entry points)
//   speculative ONE pass into fixed-capacity rows (row capacity = largest row of the previous call + 25 %), no
//               count / scan; the largest row comes back asynchronously and the host verifies it afterwards.
// Consumers read a row as [nbr_off[i], nbr_end[i]) in both layouts.
//
// Bit-exactness: the accept test d < cutoff uses the reference's operation order without FMA
// (image_diff/norm_nofma in gap_device.cuh), on the ORIGINAL positions and the total integer shift
// (cell image - map_shift(i) + map_shift(j), Connection.f95:1275), so the set of (i, j, shift, d)
// equals the reference's whatever the binning details are.
#include <cub/cub.cuh>

#include "gap_device.cuh"

namespace gapb200 {

namespace {

__device__ __forceinline__ void frac_coords(const double* g, const double* p, double* t) {
#pragma unroll
  for (int r = 0; r < 3; r++) t[r] = g[r] * p[0] + g[r + 3] * p[1] + g[r + 6] * p[2];
}

__global__ void k_frac_minmax(const double* __restrict__ pos, int N, CellGrid grid, double* __restrict__ part /* [gridDim][6] */) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    double t[3];
    frac_coords(grid.g, pos + 3 * (size_t)i, t);
#pragma unroll
    for (int r = 0; r < 3; r++) {
      mn[r] = fmin(mn[r], t[r]);
      mx[r] = fmax(mx[r], t[r]);
    }
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int r = 0; r < 3; r++) {
    double a = BR(tmp).Reduce(mn[r], cub::Min());
    __syncthreads();
    double b = BR(tmp).Reduce(mx[r], cub::Max());
    __syncthreads();
    if (threadIdx.x == 0) {
      part[6 * blockIdx.x + r] = a;
      part[6 * blockIdx.x + 3 + r] = b;
    }
  }
}
__global__ void k_minmax_final(const double* __restrict__ part, int nb, double* __restrict__ out6) {
  int r = threadIdx.x;
  if (r >= 6) return;
  double v = part[r];
  for (int b = 1; b < nb; b++) v = r < 3 ? fmin(v, part[6 * b + r]) : fmax(v, part[6 * b + r]);
  out6[r] = v;
}

// err_flag: set when an atom lies more than MAX_MAP_SHIFT periodic images away from the cell.  Shifts are carried as int8: with
// |map_shift| <= 30 and at most 60 cell images (build_connect) every intermediate s - map_shift(i) (<= 90) and every total shift
// s - map_shift(i) + map_shift(j) (<= 120) fits.
constexpr int MAX_MAP_SHIFT = 30;
// cell_count != NULL: also count the atoms of every cell (counting-sort path)
__global__ void k_bin(const double* __restrict__ pos, int N, CellGrid grid, int* __restrict__ cell_of, int* __restrict__ mshift,
                      int* __restrict__ iota, int* __restrict__ err_flag, int* __restrict__ cell_count) {
  pdl_launch_dependents();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double t[3];
  frac_coords(grid.g, pos + 3 * (size_t)i, t);
  int c[3], ms[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    double u;
    if (grid.pbc[r]) {
      double fl = -floor(t[r] + 0.5);  // map_shift, Connection.f95:1574
      if (!(fabs(fl) <= (double)MAX_MAP_SHIFT)) { atomicOr(err_flag, 1); fl = 0.0; }
      ms[r] = (int)fl;
      u = t[r] + (double)ms[r] + 0.5;
    } else {
      ms[r] = 0;
      u = (t[r] - grid.toff[r]) * grid.tscale[r];
    }
    int ci = (int)floor((double)grid.n[r] * u);
    c[r] = ci < 0 ? 0 : (ci >= grid.n[r] ? grid.n[r] - 1 : ci);
  }
  int cell = (c[2] * grid.n[1] + c[1]) * grid.n[0] + c[0];
  cell_of[i] = cell;
  mshift[i] = pack_shift(ms[0], ms[1], ms[2]);
  iota[i] = i;
  if (cell_count) atomicAdd(&cell_count[cell], 1);
}

// counting sort, step 2: every atom takes a slot of its cell's range (in arbitrary order; the counts are consumed)
__global__ void k_place(const int* __restrict__ cell_of, int N, const int* __restrict__ cell_start, int* __restrict__ cell_count,
                        int* __restrict__ slots) {
  pdl_launch_dependents();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int c = cell_of[i];
  slots[cell_start[c] + atomicSub(&cell_count[c], 1) - 1] = i;
}
// step 3: one thread per slot: the atom's place inside its cell is its rank by atom index (what the stable radix sort delivers, so
// both paths give the same arrays); positions are gathered on the way
__global__ void k_cell_sort_gather(const double* __restrict__ pos, const int* __restrict__ mshift, const int* __restrict__ cell_of,
                                   const int* __restrict__ slots, int N, const int* __restrict__ cell_start, int* __restrict__ sort_idx,
                                   int* __restrict__ sort_keys, double* __restrict__ spos, int* __restrict__ smshift, int* __restrict__ slot_of) {
  pdl_launch_dependents();
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const int i = slots[t], c = cell_of[i];
  const int b = cell_start[c], e = cell_start[c + 1];
  int rank = 0;
  for (int q = b; q < e; q++) rank += slots[q] < i;  // a cell holds a handful of atoms
  const int p = b + rank;
  sort_idx[p] = i;
  slot_of[i] = p;
  sort_keys[p] = c;
  spos[3 * (size_t)p + 0] = pos[3 * (size_t)i + 0];
  spos[3 * (size_t)p + 1] = pos[3 * (size_t)i + 1];
  spos[3 * (size_t)p + 2] = pos[3 * (size_t)i + 2];
  smshift[p] = mshift[i];
}

// positions in cell-sorted order, and the cell ranges from the sorted keys: cell_start[c] = first sorted slot whose
// key is >= c (every thread fills the cells between its predecessor's key and its own; cell_start[ncell] = N)
__global__ void k_gather_sorted(const double* __restrict__ pos, const int* __restrict__ mshift, const int* __restrict__ sort_idx,
                                const int* __restrict__ sort_keys, int N, int ncell, double* __restrict__ spos, int* __restrict__ smshift,
                                int* __restrict__ cell_start, int* __restrict__ slot_of) {
  pdl_launch_dependents();
  pdl_wait();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int key = sort_keys[p], prev = p > 0 ? sort_keys[p - 1] : -1;
  for (int c = prev + 1; c <= key; c++) cell_start[c] = p;
  if (p == N - 1)
    for (int c = key + 1; c <= ncell; c++) cell_start[c] = N;
  int i = sort_idx[p];
  slot_of[i] = p;
  spos[3 * (size_t)p + 0] = pos[3 * (size_t)i + 0];
  spos[3 * (size_t)p + 1] = pos[3 * (size_t)i + 1];
  spos[3 * (size_t)p + 2] = pos[3 * (size_t)i + 2];
  smshift[p] = mshift[i];
}

__device__ __forceinline__ int floor_div(int a, int n) { return (a >= 0) ? a / n : -((-a + n - 1) / n); }

// One WARP per atom (in cell-sorted order).  The (2R+1)^3 surrounding cells are taken 32 at a time: lane k looks up
// cell k's range and image shift, a warp scan turns the counts into a flattened candidate index space, and the lanes
// sweep that space 32 candidates per step (a shuffle-based binary search maps a candidate to its cell).  Accepted
// pairs are written in candidate order through ballot/popc, so the list order is deterministic: cells in (o2,o1,o0)
// order, ascending atom index inside a cell.
constexpr int NEIGH_WARPS = 4;
// MODE 0: count (nn[i], largest row -> *max_row) ; 1: fill packed CSR rows at nbr_off[i] ; 2: one pass into rows of
// fixed capacity row_cap at (i - first) * row_cap, writing nbr_off[i] / nbr_end[i] and the largest row
enum { NEIGH_COUNT = 0, NEIGH_FILL = 1, NEIGH_ONEPASS = 2 };

template <int MODE>
__global__ void __launch_bounds__(NEIGH_WARPS * 32) k_neigh(int N, int first, int last, CellGrid grid, const int* __restrict__ sort_idx,
                                                            const int* __restrict__ slot_of, const int* __restrict__ sort_keys, const double* __restrict__ spos,
                                                            const int* __restrict__ smshift, const int* __restrict__ cell_start,
                                                            int* __restrict__ nn, int* __restrict__ nbr_off, int* __restrict__ nbr_end,
                                                            int* __restrict__ nbr_j, int* __restrict__ nbr_s, double* __restrict__ nbr_d, int cap,
                                                            int row_cap, int* __restrict__ max_row) {
  constexpr bool FILL = MODE != NEIGH_COUNT;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // one warp per CENTRE of this partition (descriptor_atomic_MPI_setup mask: atoms [first, last)); the rows of the other atoms are
  // never read (the descriptor kernels only visit centres of the partition), except by the count / scan of the exact layout
  const int i = first + blockIdx.x * NEIGH_WARPS + (threadIdx.x >> 5);
  if (i >= last) return;
  const int p = slot_of[i];
  const int cell = sort_keys[p];
  const int c0 = cell % grid.n[0], c1 = (cell / grid.n[0]) % grid.n[1], c2 = cell / (grid.n[0] * grid.n[1]);
  const double pi[3] = {spos[3 * (size_t)p], spos[3 * (size_t)p + 1], spos[3 * (size_t)p + 2]};
  int mi0, mi1, mi2;
  unpack_shift(smshift[p], mi0, mi1, mi2);
  const int w0 = 2 * grid.R[0] + 1, w1 = 2 * grid.R[1] + 1, w2 = 2 * grid.R[2] + 1;
  const int n_nb_cells = w0 * w1 * w2;
  const double cutoff = grid.cutoff;
  int count = 0;
  const int row_beg = MODE == NEIGH_FILL ? nbr_off[i] : (MODE == NEIGH_ONEPASS ? (i - first) * row_cap : 0);
  if (MODE == NEIGH_ONEPASS) cap = row_beg + row_cap;  // a fixed-capacity row ends where the next one begins
  int wpos = row_beg;
  for (int cbase = 0; cbase < n_nb_cells; cbase += 32) {
    // lane -> one neighbouring cell
    int ck = cbase + lane, qb = 0, cnt = 0, sh = 0;
    if (ck < n_nb_cells) {
      int o0 = ck % w0 - grid.R[0], o1 = (ck / w0) % w1 - grid.R[1], o2 = ck / (w0 * w1) - grid.R[2];
      int x = c0 + o0, y = c1 + o1, z = c2 + o2, s0 = 0, s1 = 0, s2 = 0;
      bool ok = true;
      if (grid.pbc[0]) { s0 = floor_div(x, grid.n[0]); x -= s0 * grid.n[0]; } else if (x < 0 || x >= grid.n[0]) ok = false;
      if (grid.pbc[1]) { s1 = floor_div(y, grid.n[1]); y -= s1 * grid.n[1]; } else if (y < 0 || y >= grid.n[1]) ok = false;
      if (grid.pbc[2]) { s2 = floor_div(z, grid.n[2]); z -= s2 * grid.n[2]; } else if (z < 0 || z >= grid.n[2]) ok = false;
      if (ok) {
        int nc = (z * grid.n[1] + y) * grid.n[0] + x;
        qb = cell_start[nc];
        cnt = cell_start[nc + 1] - qb;
        sh = pack_shift(s0 - mi0, s1 - mi1, s2 - mi2);
      }
    }
    int pre = cnt;  // inclusive scan of the counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    // candidate t of the flattened index space -> (sorted slot q, image shift) by a shuffle-based binary search; the loads of a
    // candidate (its shift word, position and atom id) are issued one batch AHEAD of their use, so two batches are in flight
    struct Cand { int q, sh, ms, j; double x, y, z; bool in; };
    auto fetch = [&](int base) {
      Cand c;
      const int t = base + lane;
      int k = 0;  // smallest k with pre_k > t
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int v = __shfl_sync(0xffffffffu, pre, k + o - 1);
        if (v <= t) k += o;
      }
      const int pk = __shfl_sync(0xffffffffu, pre, k), ck_cnt = __shfl_sync(0xffffffffu, cnt, k), kqb = __shfl_sync(0xffffffffu, qb, k);
      c.sh = __shfl_sync(0xffffffffu, sh, k);
      c.in = t < total;
      c.q = c.in ? kqb + (t - (pk - ck_cnt)) : 0;
      c.ms = 0; c.j = 0; c.x = c.y = c.z = 0.0;
      if (c.in) {
        c.ms = smshift[c.q];
        c.x = spos[3 * (size_t)c.q]; c.y = spos[3 * (size_t)c.q + 1]; c.z = spos[3 * (size_t)c.q + 2];
        if (FILL) c.j = sort_idx[c.q];
      }
      return c;
    };
    Cand nxt = fetch(0);
    for (int base = 0; base < total; base += 32) {
      const Cand cur = nxt;
      if (base + 32 < total) nxt = fetch(base + 32);
      bool acc = false;
      int t0 = 0, t1 = 0, t2 = 0;
      double d = 0.0;
      if (cur.in) {
        int mj0, mj1, mj2, b0, b1, b2;
        unpack_shift(cur.ms, mj0, mj1, mj2);
        unpack_shift(cur.sh, b0, b1, b2);
        t0 = b0 + mj0; t1 = b1 + mj1; t2 = b2 + mj2;
        if (!(cur.q == p && t0 == 0 && t1 == 0 && t2 == 0)) {  // self, zero shift (:1266-1272)
          const double pj[3] = {cur.x, cur.y, cur.z};
          double dd[3];
          image_diff(pi, pj, grid.lat, t0, t1, t2, dd);
          d = norm_nofma(dd);
          acc = d < cutoff;  // strict, Connection.f95:517
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, acc);
      if (FILL && acc) {
        int w = wpos + __popc(bal & ((1u << lane) - 1u));
        if (w < cap) {  // only a speculatively sized row can overflow; the host then repeats the call with the exact layout
          nbr_j[w] = cur.j;
          nbr_s[w] = pack_shift(t0, t1, t2);
          if (nbr_d) nbr_d[w] = d;
        }
      }
      wpos += __popc(bal);
      count += __popc(bal);
    }
  }
  if (lane == 0) {
    if (MODE == NEIGH_COUNT) nn[i] = count;
    if (MODE == NEIGH_ONEPASS) {
      nbr_off[i] = row_beg;
      nbr_end[i] = row_beg + (count < row_cap ? count : row_cap);
    }
    // largest row: the plain read keeps all but a handful of warps away from the atomic
    if (MODE != NEIGH_FILL && count > *(volatile int*)max_row) atomicMax(max_row, count);
  }
}

}  // namespace

size_t neighbour_cub_bytes(int N, int ncell) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, N);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, ncell + 1);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (int*)nullptr, (int*)nullptr, N + 1);
  size_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

void launch_frac_minmax(const double* pos, int N, const double*, const CellGrid& grid, double* minmax6, cudaStream_t st, int* launches) {
  // scratch for block partials is carved from the tail of minmax6 (allocated as 6 + 6*64 doubles)
  int nb = (N + 255) / 256;
  if (nb > 64) nb = 64;
  if (nb < 1) nb = 1;
  k_frac_minmax<<<nb, 256, 0, st>>>(pos, N, grid, minmax6 + 6);
  k_minmax_final<<<1, 32, 0, st>>>(minmax6 + 6, nb, minmax6);
  *launches += 2;
}

// up to this many atoms cub sorts in ONE single-tile kernel; beyond it its multi-kernel radix sort costs more than a counting sort
constexpr int SINGLE_TILE_SORT_MAX = 4864;

void launch_bin_atoms(const double* pos, int N, const CellGrid& grid, int ncell, NeighbourWork& w, cudaStream_t st, int* launches) {
  int nb = (N + 255) / 256;
  if (N > SINGLE_TILE_SORT_MAX) {
    // counting sort by cell: count (atomics), exclusive scan, place, then every atom finds its rank by index inside its cell
    // k_place hands every count back (atomicSub down to zero): the array is zeroed once, when it is allocated (potential.cu)
    launch_pdl(k_bin, dim3(nb), dim3(256), 0, st, pos, N, grid, w.cell_of, w.mshift, w.iota, w.err_flag, w.cell_count);
    size_t bytes = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, w.cell_count, w.cell_start, ncell + 1, st);
    launch_pdl(k_place, dim3(nb), dim3(256), 0, st, (const int*)w.cell_of, N, (const int*)w.cell_start, w.cell_count, w.iota);
    launch_pdl(k_cell_sort_gather, dim3(nb), dim3(256), 0, st, pos, (const int*)w.mshift, (const int*)w.cell_of, (const int*)w.iota, N, (const int*)w.cell_start,
               w.sort_idx, w.sort_keys, w.spos, w.smshift, w.slot_of);
    *launches += 4;
    return;
  }
  launch_pdl(k_bin, dim3(nb), dim3(256), 0, st, pos, N, grid, w.cell_of, w.mshift, w.iota, w.err_flag, (int*)nullptr);
  int bits = 1;
  while ((1 << bits) < ncell && bits < 31) bits++;
  size_t bytes = w.cub_bytes;
  cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, w.cell_of, w.sort_keys, w.iota, w.sort_idx, N, 0, bits, st);
  launch_pdl(k_gather_sorted, dim3(nb), dim3(256), 0, st, pos, (const int*)w.mshift, (const int*)w.sort_idx, (const int*)w.sort_keys, N, ncell, w.spos,
             w.smshift, w.cell_start, w.slot_of);
  *launches += 3;
}

void launch_neigh_count(const double* pos, int N, int first, int last, const CellGrid& grid, NeighbourWork& w, int* nbr_off, int* max_row,
                        cudaStream_t st, int* launches) {
  (void)pos;
  int nb = (last - first + NEIGH_WARPS - 1) / NEIGH_WARPS;
  cudaMemsetAsync(w.nn, 0, sizeof(int) * (N + 1), st);  // rows of atoms outside the partition are empty
  if (nb > 0)
    k_neigh<NEIGH_COUNT><<<nb, NEIGH_WARPS * 32, 0, st>>>(N, first, last, grid, w.sort_idx, w.slot_of, w.sort_keys, w.spos, w.smshift, w.cell_start, w.nn,
                                                          nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, max_row);
  size_t bytes = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, w.nn, nbr_off, N + 1, st);
  *launches += 2;
}

void launch_neigh_fill(const double* pos, int N, int first, int last, const CellGrid& grid, NeighbourWork& w, int* nbr_off, int* nbr_j,
                       int* nbr_s, double* nbr_d, int cap, cudaStream_t st, int* launches) {
  (void)pos;
  int nb = (last - first + NEIGH_WARPS - 1) / NEIGH_WARPS;
  if (nb > 0)
    k_neigh<NEIGH_FILL><<<nb, NEIGH_WARPS * 32, 0, st>>>(N, first, last, grid, w.sort_idx, w.slot_of, w.sort_keys, w.spos, w.smshift, w.cell_start, w.nn,
                                                         nbr_off, nullptr, nbr_j, nbr_s, nbr_d, cap, 0, nullptr);
  *launches += 1;
}

void launch_neigh_onepass(const double* pos, int N, int first, int last, const CellGrid& grid, NeighbourWork& w, int* nbr_off, int* nbr_end,
                          int* nbr_j, int* nbr_s, int row_cap, int* max_row, cudaStream_t st, int* launches) {
  (void)pos;
  int nb = (last - first + NEIGH_WARPS - 1) / NEIGH_WARPS;
  if (nb > 0)
    launch_pdl(k_neigh<NEIGH_ONEPASS>, dim3(nb), dim3(NEIGH_WARPS * 32), 0, st, N, first, last, grid, (const int*)w.sort_idx, (const int*)w.slot_of,
               (const int*)w.sort_keys, (const double*)w.spos, (const int*)w.smshift, (const int*)w.cell_start, w.nn, nbr_off, nbr_end, nbr_j, nbr_s,
               (double*)nullptr, 0, row_cap, max_row);
  *launches += 1;
}

}  // namespace gapb200
