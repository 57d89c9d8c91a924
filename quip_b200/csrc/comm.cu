// comm.cu -- the reduction of the per-rank partials [E | virial | F] over the GPUs of one node (see gap_comm.h).
//
// Reference semantics: sum_in_place(mpi, f) etc. = MPI_Allreduce(MPI_IN_PLACE, ..., MPI_SUM) (src/libAtoms/MPI_context.f95:668-694),
// issued five times at the end of IPModel_GAP_Calc (src/Potentials/IPModel_GAP.f95:538-556).  Here: ONE reduction of one packed
// buffer, enqueued on the evaluation's stream, by NCCL or -- for latency-bound sizes -- by one kernel over NVLink peer mappings: a one-shot
// pull of all ranks' partials (2-3 ranks) or a reduce-scatter + all-gather by push in self-validating 16-byte cells (4 ranks and more).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gap_comm.h"
#include "gap_device.cuh"
#include "gap_model.h"

namespace gapb200 {

namespace {

#define CUDA_OK(expr)                                                                                               \
  do {                                                                                                              \
    cudaError_t _e = (expr);                                                                                        \
    if (_e != cudaSuccess)                                                                                          \
      throw GapError(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// ---- libnccl, loaded on first use.  A process that already carries an NCCL (PyTorch bundles one) gets THAT copy: glibc
//      resolves the soname against the objects already loaded.  GAP_B200_NCCL_LIB overrides the name.
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl& nccl() {
  static Nccl n = [] {
    Nccl t;
    const char* env = getenv("GAP_B200_NCCL_LIB");
    const char* names[] = {env && *env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      t.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (t.lib) break;
    }
    if (!t.lib) return t;
    auto sym = [&](const char* s) { return dlsym(t.lib, s); };
    t.GetVersion = (decltype(t.GetVersion))sym("ncclGetVersion");
    t.GetUniqueId = (decltype(t.GetUniqueId))sym("ncclGetUniqueId");
    t.CommInitRank = (decltype(t.CommInitRank))sym("ncclCommInitRank");
    t.CommDestroy = (decltype(t.CommDestroy))sym("ncclCommDestroy");
    t.AllReduce = (decltype(t.AllReduce))sym("ncclAllReduce");
    t.AllGather = (decltype(t.AllGather))sym("ncclAllGather");
    t.GetErrorString = (decltype(t.GetErrorString))sym("ncclGetErrorString");
    return t;
  }();
  if (!n.lib || !n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.AllGather || !n.CommDestroy)
    throw GapError("gap_comm: libnccl.so.2 could not be loaded (set GAP_B200_NCCL_LIB); a multi-GPU run needs NCCL");
  return n;
}
void nccl_ok(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return;
  const char* s = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
  throw GapError(std::string("NCCL error in ") + what + ": " + s);
}

constexpr int P2P_MAX_RANKS = 16;
constexpr size_t P2P_FLAG_BYTES = 256;                // [n_ranks] step counters, one 256-byte line
// a rank waits this long for the other ranks' partials before it reports an error (GAP_B200_P2P_TIMEOUT_S, default 60 s: ranks of an MPI host
// may reach the evaluation seconds apart; an NCCL all-reduce would wait for ever)
unsigned long long p2p_timeout_ns() {
  static unsigned long long v = 0;
  if (!v) {
    const char* e = getenv("GAP_B200_P2P_TIMEOUT_S");
    double sec = e && *e ? atof(e) : 60.0;
    if (!(sec > 0.0)) sec = 60.0;
    v = (unsigned long long)(sec * 1e9);
  }
  return v;
}

struct PeerPtrs {
  char* base[P2P_MAX_RANKS];  // base[r]: rank r's block (flags | buffer 0 | buffer 1) as mapped into THIS process
};

// One-shot all-reduce out of peer memory.  Every rank runs this kernel on its own GPU:
//   1. block 0 tells every rank (itself included) that this rank's partial of step `step` is complete: the partial was written
//      by earlier kernels of the same stream, so it is globally visible when this kernel starts;
//   2. every block waits until all G ranks have said so (flags live in the waiter's own memory: local polling);
//   3. the blocks sum the G partials element-wise in RANK order (bit-identical on every rank) reading peers over NVLink with
//      L1-bypassing 16-byte loads, all G loads of an element in flight at once, and store the totals to the local result.
// The partial buffers alternate between steps, so a fast rank can start writing step s+1 while a slow one still reads step s;
// it cannot reach step s+2 (same buffer again) before every rank has signalled s+1, i.e. has left the kernel of step s.
__global__ void __launch_bounds__(256) k_peer_allreduce(PeerPtrs peers, int rank, int n, unsigned step, size_t buf_off, size_t count,
                                                        double* __restrict__ result, int* __restrict__ err, unsigned long long* __restrict__ stamps,
                                                        unsigned long long timeout_ns) {
  __shared__ int timed_out;
  pdl_launch_dependents();
  if (threadIdx.x == 0) timed_out = 0;
  pdl_wait();  // this rank's partial is complete and visible
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(stamps[0]));  // partial ready (ns)
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < n) {
    unsigned* f = reinterpret_cast<unsigned*>(peers.base[threadIdx.x]) + rank;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(f), "r"(step) : "memory");
  }
  if (threadIdx.x < n) {
    const unsigned* f = reinterpret_cast<const unsigned*>(peers.base[rank]) + threadIdx.x;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - step) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
      if (t1 - t0 > timeout_ns) { timed_out = 1; break; }
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(stamps[1]));  // every rank's partial ready
  if (timed_out) {
    if (threadIdx.x == 0) {  // err lives in mapped host memory: the host reads it after its synchronisation
      *(volatile int*)err = 1;
      __threadfence_system();
    }
    return;
  }
  const size_t n2 = count >> 1, stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t idx = t; idx < n2; idx += stride) {
    double sx = 0.0, sy = 0.0;
    for (int r0 = 0; r0 < n; r0 += 8) {
      double vx[8], vy[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        vx[k] = vy[k] = 0.0;
        if (r0 + k < n) {
          const char* p = peers.base[r0 + k] + buf_off + 16 * idx;
          asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];\n" : "=d"(vx[k]), "=d"(vy[k]) : "l"(p));
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (r0 + k < n) { sx += vx[k]; sy += vy[k]; }
    }
    *reinterpret_cast<double2*>(result + 2 * idx) = make_double2(sx, sy);
  }
  if ((count & 1) && t == 0) {  // odd tail element
    double s = 0.0;
    for (int r = 0; r < n; r++) s += __ldcg(reinterpret_cast<const double*>(peers.base[r] + buf_off) + (count - 1));
    result[count - 1] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(stamps[2]));  // block 0 done (a sample of the sum phase)
}


// LOW-LATENCY variant (reduce-scatter + all-gather by PUSH, no flag rounds, no fences, no grid barrier) for latency-bound payloads on 4 or
// more ranks.  Every double travels in a 16-byte cell {low word, step, high word, step}: the two 8-byte halves validate themselves, so the
// receiver simply polls its LOCAL memory until both step words match (the protocol NCCL calls LL).  One launch:
//   1. every rank pushes element i of its partial into the cell (src = me, i) of the rank that owns i (contiguous slices),
//   2. the owner polls its G - 1 incoming cells per element, adds the G values in rank order, stores the total to its own result and
//      pushes it into cell i of every peer,
//   3. every rank polls the cells of the elements it does not own and copies the totals to its result.
// Remote traffic is 2 (G - 1) / G buffer volumes per rank in 16-byte posted WRITES (the one-shot kernel: G - 1 volumes of remote READS); the
// critical path is two one-way NVLink latencies.  Cells alternate between two sets by step parity: a rank can be at most one step ahead of
// a peer that still polls (it needs that peer's step-(s+1) data to finish step s+1), so a set is overwritten only after it has been read.
// All ranks end with the same bits.  Blocks spin on data produced by other GPUs' blocks: the launcher keeps the grid co-resident.
struct LLGeom {
  size_t rs_off, ag_off;   // byte offsets of the two cell regions in a rank's block
  unsigned slice_cap, cap; // cells per (parity, source) in the reduce-scatter region; cells per parity in the all-gather region
};
__device__ __forceinline__ void ll_store(char* p, double v, unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ bool ll_poll(const char* p, unsigned flag, double& v, unsigned long long timeout_ns) {
  unsigned a, b, c, d;
  unsigned long long t0 = 0, t1;
  for (unsigned spin = 0;; spin++) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (b == flag && d == flag) break;
    if ((spin & 1023u) == 1023u) {
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > timeout_ns) return false;
    }
  }
  v = __hiloint2double((int)c, (int)a);
  return true;
}
__global__ void __launch_bounds__(256) k_peer_allreduce_ll(PeerPtrs peers, int rank, int n, unsigned step, size_t buf_off, LLGeom g, unsigned count,
                                                           double* __restrict__ result, int* __restrict__ err, unsigned long long* __restrict__ stamps,
                                                           unsigned long long timeout_ns) {
  pdl_launch_dependents();
  pdl_wait();  // this rank's partial is complete
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(stamps[0]));
  const unsigned stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x, par = step & 1u;
  const double* partial = reinterpret_cast<const double*>(peers.base[rank] + buf_off);
  auto slice_lo = [&](unsigned sr) { return (unsigned)(((unsigned long long)count * sr) / (unsigned)n); };
  const unsigned lo = slice_lo((unsigned)rank), hi = slice_lo((unsigned)rank + 1u);
  // 1. push my contribution to the owners
  for (unsigned idx = t; idx < count; idx += stride) {
    if (idx >= lo && idx < hi) continue;
    unsigned sr = (unsigned)(((unsigned long long)idx * (unsigned)n) / count);
    while (idx >= slice_lo(sr + 1u)) sr++;
    char* dst = peers.base[sr] + g.rs_off + ((size_t)(par * (unsigned)n + (unsigned)rank) * g.slice_cap + (idx - slice_lo(sr))) * 16;
    ll_store(dst, partial[idx], step);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(stamps[1]));
  // 2. my slice: sum in rank order, publish the totals
  bool ok = true;
  for (unsigned idx = lo + t; idx < hi && ok; idx += stride) {
    double sum = 0.0;
    for (int r = 0; r < n && ok; r++) {
      double v;
      if (r == rank) v = partial[idx];
      else ok = ll_poll(peers.base[rank] + g.rs_off + ((size_t)(par * (unsigned)n + (unsigned)r) * g.slice_cap + (idx - lo)) * 16, step, v, timeout_ns);
      sum += v;
    }
    if (!ok) break;
    result[idx] = sum;
    for (int r = 0; r < n; r++)
      if (r != rank) ll_store(peers.base[r] + g.ag_off + ((size_t)par * g.cap + idx) * 16, sum, step);
  }
  // 3. the other slices' totals
  for (unsigned idx = t; idx < count && ok; idx += stride) {
    if (idx >= lo && idx < hi) continue;
    double v;
    ok = ll_poll(peers.base[rank] + g.ag_off + ((size_t)par * g.cap + idx) * 16, step, v, timeout_ns);
    if (ok) result[idx] = v;
  }
  if (!ok) {
    *(volatile int*)err = 1;
    __threadfence_system();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(stamps[2]));
}

}  // namespace

struct GapComm {
  ncclComm_t comm = nullptr;
  int rank = 0, n = 1, device = 0, n_sm = 148;
  // peer-memory transport
  bool p2p_enabled = true;          // GAP_B200_P2P=0 turns it off; a failed handle exchange turns it off for good
  size_t p2p_limit_bytes = 8u << 20;  // largest partial buffer reduced out of peer memory (beyond it NCCL's bandwidth algorithms win)
  char* block = nullptr;            // local: [flags | buffer 0 | buffer 1]
  size_t cap = 0;                   // doubles per buffer
  void* peer_base[P2P_MAX_RANKS] = {nullptr};
  unsigned step = 0;
  bool partial_is_peer = false;     // the partial of the step in flight lives in block (else: in the result buffer)
  int* d_err = nullptr;             // device address of h_err
  int* h_err = nullptr;             // pinned + mapped: set by the peer kernel if it timed out
  unsigned long long* d_stamps = nullptr;  // [3] globaltimer of the last peer reduction: own partial ready, all partials ready, done
  char* d_handles = nullptr;        // [n + 1] x 64 bytes (slot n: this rank's handle, the send buffer)
  int* d_okflag = nullptr;
  LLGeom ll{0, 0, 0, 0};            // cell regions of the low-latency kernel (cap = 0: not allocated for this block)
  size_t ll_limit = 1u << 20;       // largest payload (doubles) the low-latency kernel takes (GAP_B200_P2P_LL_MAX_DOUBLES; 0 = never)
  int ll_min_ranks = 4;
  const char* last = "none";
  long launches = 0;
};

void comm_get_unique_id(char* id128) {
  static_assert(sizeof(ncclUniqueId) == COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  nccl_ok(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
}

GapComm* comm_create(const char* id128, int rank, int n_ranks, int device) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw GapError("gap_potential_set_comm: need 0 <= rank < n_ranks");
  CUDA_OK(cudaSetDevice(device));
  GapComm* c = new GapComm();
  c->rank = rank; c->n = n_ranks; c->device = device;
  try {
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    nccl_ok(nccl().CommInitRank(&c->comm, n_ranks, id, rank), "ncclCommInitRank");
    const char* e = getenv("GAP_B200_P2P");
    if ((e && *e == '0') || n_ranks > P2P_MAX_RANKS || n_ranks < 2) c->p2p_enabled = false;
    const char* lim = getenv("GAP_B200_P2P_MAX_BYTES");
    if (lim && *lim) c->p2p_limit_bytes = (size_t)strtoull(lim, nullptr, 10);
    CUDA_OK(cudaHostAlloc((void**)&c->h_err, sizeof(int), cudaHostAllocMapped));
    *c->h_err = 0;
    CUDA_OK(cudaHostGetDevicePointer((void**)&c->d_err, c->h_err, 0));
    CUDA_OK(cudaMalloc(&c->d_handles, (size_t)(n_ranks + 1) * sizeof(cudaIpcMemHandle_t)));
    CUDA_OK(cudaMalloc(&c->d_okflag, sizeof(int)));
    const char* llm = getenv("GAP_B200_P2P_LL_MAX_DOUBLES");
    if (llm && *llm) c->ll_limit = (size_t)strtoull(llm, nullptr, 10);
    const char* llr = getenv("GAP_B200_P2P_LL_MIN_RANKS");
    if (llr && *llr) c->ll_min_ranks = atoi(llr);
    CUDA_OK(cudaMalloc(&c->d_stamps, 8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemset(c->d_stamps, 0, 8 * sizeof(unsigned long long)));
  } catch (...) {
    comm_destroy(c);
    throw;
  }
  return c;
}

static void p2p_release(GapComm* c) {
  for (int r = 0; r < c->n && r < P2P_MAX_RANKS; r++) {
    if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    c->peer_base[r] = nullptr;
  }
  if (c->block) cudaFree(c->block);
  c->block = nullptr;
  c->cap = 0;
  c->ll = LLGeom{0, 0, 0, 0};
}

void comm_destroy(GapComm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  p2p_release(c);
  if (c->comm) nccl().CommDestroy(c->comm);
  cudaFree(c->d_handles);
  cudaFree(c->d_okflag);
  cudaFree(c->d_stamps);
  if (c->h_err) cudaFreeHost(c->h_err);
  delete c;
}

int comm_rank(const GapComm* c) { return c ? c->rank : 0; }
int comm_size(const GapComm* c) { return c ? c->n : 1; }
const char* comm_last_transport(const GapComm* c) { return c ? c->last : "none"; }
void comm_last_stamps(const GapComm* c, double* wait_us, double* sum_us) {
  *wait_us = *sum_us = 0.0;
  if (!c || !c->d_stamps) return;
  unsigned long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpy(t, c->d_stamps, sizeof(t), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  if (t[1] >= t[0]) *wait_us = (double)(t[1] - t[0]) * 1e-3;
  if (t[2] >= t[1]) *sum_us = (double)(t[2] - t[1]) * 1e-3;
}
long comm_launch_count(const GapComm* c) { return c ? c->launches : 0; }

// (Re)allocate the peer-visible block for `count` doubles per buffer and map every other rank's block.  Collective.
static void p2p_setup(GapComm* c, size_t count, cudaStream_t st) {
  const size_t want = ((count + count / 4 + 1024) + 1) & ~(size_t)1;
  CUDA_OK(cudaStreamSynchronize(st));
  char* nb = nullptr;
  size_t bytes = P2P_FLAG_BYTES + 2 * want * sizeof(double);  // flags | partial 0 | partial 1 | low-latency cells
  LLGeom ll{0, 0, 0, 0};
  if (c->n >= c->ll_min_ranks && count <= c->ll_limit && want < (1u << 30)) {
    ll.cap = (unsigned)want;
    ll.slice_cap = (unsigned)(want / (size_t)c->n + 2);
    ll.rs_off = (bytes + 255) & ~(size_t)255;
    ll.ag_off = ll.rs_off + (size_t)2 * c->n * ll.slice_cap * 16;
    bytes = ll.ag_off + (size_t)2 * ll.cap * 16;
  }
  bool ok = cudaMalloc(&nb, bytes) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = cudaMemset(nb, 0, P2P_FLAG_BYTES) == cudaSuccess && cudaIpcGetMemHandle(&mine, nb) == cudaSuccess;
  if (ok && ll.cap) ok = cudaMemset(nb + ll.rs_off, 0, bytes - ll.rs_off) == cudaSuccess;  // step words: 0 matches no step
  cudaGetLastError();
  // every rank takes part in the exchange even if its own allocation failed (the verdict below is collective)
  CUDA_OK(cudaMemcpyAsync(c->d_handles + (size_t)c->n * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  nccl_ok(nccl().AllGather(c->d_handles + (size_t)c->n * sizeof(mine), c->d_handles, sizeof(mine), ncclChar, c->comm, st), "ncclAllGather (peer handles)");
  std::vector<cudaIpcMemHandle_t> all(c->n);
  CUDA_OK(cudaMemcpyAsync(all.data(), c->d_handles, (size_t)c->n * sizeof(mine), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));  // also: every rank has left the kernels that used the old block
  p2p_release(c);
  void* opened[P2P_MAX_RANKS] = {nullptr};
  for (int r = 0; r < c->n && ok; r++) {
    if (r == c->rank) { opened[r] = nb; continue; }
    if (cudaIpcOpenMemHandle(&opened[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened[r] = nullptr; ok = false; }
  }
  cudaGetLastError();
  int bad = ok ? 0 : 1;
  CUDA_OK(cudaMemcpyAsync(c->d_okflag, &bad, sizeof(int), cudaMemcpyHostToDevice, st));
  nccl_ok(nccl().AllReduce(c->d_okflag, c->d_okflag, 1, ncclInt32, ncclSum, c->comm, st), "ncclAllReduce (peer mapping verdict)");
  CUDA_OK(cudaMemcpyAsync(&bad, c->d_okflag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  if (bad) {  // some rank could not map its peers: everybody stays on NCCL
    for (int r = 0; r < c->n; r++)
      if (r != c->rank && opened[r]) cudaIpcCloseMemHandle(opened[r]);
    if (nb) cudaFree(nb);
    cudaGetLastError();
    c->p2p_enabled = false;
    return;
  }
  c->block = nb;
  c->cap = want;
  c->ll = ll;
  for (int r = 0; r < c->n; r++) c->peer_base[r] = opened[r];
}

double* comm_partial_buffer(GapComm* c, size_t count, double* result, cudaStream_t st) {
  if (!c || c->n < 2) return result;
  c->partial_is_peer = false;
  if (c->p2p_enabled && count * sizeof(double) <= c->p2p_limit_bytes) {
    if (count > c->cap) p2p_setup(c, count, st);
    if (c->p2p_enabled && count <= c->cap) {
      // the step counter advances when the reduction is enqueued (comm_allreduce_packed): an evaluation that throws in between
      // leaves this rank's counter where its peers expect it
      c->partial_is_peer = true;
      return reinterpret_cast<double*>(c->block + P2P_FLAG_BYTES) + (size_t)((c->step + 1u) & 1u) * c->cap;
    }
  }
  return result;
}

void comm_allreduce_packed(GapComm* c, size_t count, double* result, cudaStream_t st) {
  if (!c || c->n < 2 || count == 0) return;
  if (c->partial_is_peer) {
    c->step++;
    PeerPtrs pp;
    for (int r = 0; r < P2P_MAX_RANKS; r++) pp.base[r] = r < c->n ? (char*)c->peer_base[r] : nullptr;
    const size_t buf_off = P2P_FLAG_BYTES + (size_t)(c->step & 1u) * c->cap * sizeof(double);
    int blocks = (int)((count / 2 + 255) / 256);
    if (blocks > 2 * c->n_sm) blocks = 2 * c->n_sm;
    if (blocks < 1) blocks = 1;
    if (c->ll.cap && c->n >= c->ll_min_ranks && count <= c->ll_limit && count <= c->ll.cap && count >= (size_t)c->n) {
      int lb = (int)((count + 255) / 256);
      if (lb > 2 * c->n_sm) lb = 2 * c->n_sm;
      launch_pdl(k_peer_allreduce_ll, dim3(lb), dim3(256), 0, st, pp, c->rank, c->n, c->step, buf_off, c->ll, (unsigned)count, result, c->d_err, c->d_stamps,
                 p2p_timeout_ns());
      c->last = "p2p-ll";
    } else {
      launch_pdl(k_peer_allreduce, dim3(blocks), dim3(256), 0, st, pp, c->rank, c->n, c->step, buf_off, count, result, c->d_err, c->d_stamps, p2p_timeout_ns());
      c->last = "p2p";
    }
  } else {
    nccl_ok(nccl().AllReduce(result, result, count, ncclFloat64, ncclSum, c->comm, st), "ncclAllReduce");
    c->last = "nccl";
  }
  c->launches++;
}

void comm_allreduce_inplace(GapComm* c, double* buf, size_t count, cudaStream_t st) {
  if (!c || c->n < 2 || count == 0) return;
  nccl_ok(nccl().AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->comm, st), "ncclAllReduce");
  c->launches++;
}

void comm_check(GapComm* c) {
  if (!c || !c->h_err) return;
  if (*c->h_err) {
    *c->h_err = 0;
    throw GapError("gap_comm: the peer-memory reduction timed out waiting for another rank (did a rank fail before its evaluation?)");
  }
}

}  // namespace gapb200
