// gap_comm.h -- the path's one collective, behind the C ABI: the sum over ranks of the packed [E | virial(9) | F(3,N)]
// partials (and of local_e / local_virial when the caller asked for them).
//
// Replaces the five MPI_Allreduce calls at the end of IPModel_GAP_Calc (sum_in_place / sum, src/Potentials/IPModel_GAP.f95:538-556 ->
// src/libAtoms/MPI_context.f95:668-694).  One process per GPU; every rank holds a communicator created from a shared
// 128-byte id (gap_comm_get_unique_id on one rank, broadcast by the host -- MPI_Bcast in a Fortran host, torch.distributed in
// quip_b200.ShardedPotential).  Two transports, both on the stream the evaluation is enqueued on:
//   * ncclAllReduce(SUM, f64) over NVLink / NVSwitch (libnccl is loaded at run time: a single-GPU host does not need it);
//   * for latency-bound payloads (config A: 98 KB per 4,096 atoms) a ONE-SHOT peer-memory reduction: every rank's partial lives
//     in a buffer that all other ranks have mapped (CUDA IPC over NVLink P2P); one kernel signals "my partial is complete" to
//     every peer, waits for theirs, and sums the G partials in rank order straight out of peer memory into the caller's result
//     buffer -- one launch, no ring steps, and every rank adds in the same order, so the replicas of a sharded MD run stay
//     bit-identical.  The handles of the peer buffers are exchanged through the NCCL communicator itself (ncclAllGather).
//   * on 4 or more ranks the peer-memory reduction is a LOW-LATENCY reduce-scatter + all-gather by push instead (16-byte self-validating
//     cells, no flag rounds or fences: see k_peer_allreduce_ll in comm.cu); 17 us faster per step than the one-shot kernel on 8 GPUs.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace gapb200 {

struct GapComm;  // opaque

constexpr int COMM_ID_BYTES = 128;  // sizeof(ncclUniqueId) = GAP_COMM_ID_BYTES of the C ABI

void comm_get_unique_id(char* id128);
GapComm* comm_create(const char* id128, int rank, int n_ranks, int device);
void comm_destroy(GapComm* c);
int comm_rank(const GapComm* c);
int comm_size(const GapComm* c);

// Where the evaluation should write this rank's PARTIAL packed buffer of `count` doubles.  With the peer-memory transport it
// is one of the two peer-visible buffers (alternating per step, which removes the trailing barrier of a one-shot
// reduction); otherwise it is `result` itself and the NCCL reduction runs in place.  May (re)allocate and exchange handles:
// collective, every rank calls it with the same count.
double* comm_partial_buffer(GapComm* c, size_t count, double* result, cudaStream_t st);
// sum over ranks of the partial handed out by the last comm_partial_buffer -> result (all ranks)
void comm_allreduce_packed(GapComm* c, size_t count, double* result, cudaStream_t st);
// plain in-place NCCL all-reduce (local_e, local_virial, optional outputs)
void comm_allreduce_inplace(GapComm* c, double* buf, size_t count, cudaStream_t st);
// after the stream has been synchronised: throws GapError if the peer reduction timed out waiting for another rank
void comm_check(GapComm* c);
// "nccl" / "p2p" transport of the last packed reduction, and the number of peer reductions / NCCL calls so far
const char* comm_last_transport(const GapComm* c);
long comm_launch_count(const GapComm* c);
// last peer-memory reduction on this rank (block 0's view): one-shot kernel -- microseconds spent waiting for the other ranks' partials (= the skew
// between the ranks) and microseconds of the sum phase; low-latency kernel -- microseconds of the push phase and of everything after it
void comm_last_stamps(const GapComm* c, double* wait_us, double* sum_us);

}  // namespace gapb200
