/* gap_b200.h -- C ABI of libgapb200.so: the B200-native drop-in for QUIP's "IP GAP" evaluation path.
 *
 * Plain C types only (no torch, no C++): this is what a Fortran host binds through ISO_C_BINDING (see
 * quip_b200/fortran/gap_b200_iface.f90 and INTEGRATION.md) and what quip_b200/potential.py binds with ctypes.
 *
 * Conventions (identical to the Fortran side, so arrays can be passed without copies):
 *   pos      real(dp) pos(3,N)        -> const double[3*N], xyz of atom 0, xyz of atom 1, ...
 *   Z        integer Z(N)             -> const int[N]
 *   lattice  real(dp) lattice(3,3)    -> const double[9], column-major: columns are the cell vectors a, b, c
 *   pbc      logical is_periodic(3)   -> const int[3] (0/1)
 *   force    real(dp) f(3,N)          -> double[3*N]
 *   virial   real(dp) virial(3,3)     -> double[9] column-major
 *   local_e  real(dp) local_e(N)      -> double[N]
 *   local_virial real(dp) (9,N)       -> double[9*N]
 * Output pointers may be NULL = Fortran "optional argument absent": the quantity is not computed
 * (force, virial and local_virial all absent => no gradients at all, IPModel_GAP.f95:416-424).
 * Every function returns 0 (ERROR_NONE) on success; otherwise non-zero, and gap_last_error() describes the
 * failure (the reference's RAISE_ERROR/system_abort conditions are reported this way; the library never exits).
 * There is no CPU fallback: without a usable CUDA device initialise fails.
 */
#ifndef GAP_B200_H
#define GAP_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gap_potential gap_potential;

/* Potential_Filename_Initialise (src/Potentials/Potential.f95:438): args_str e.g. "IP GAP" or
 * "IP GAP label=..." or "" (take init_args from the XML); sparseX side files are resolved relative to the
 * XML's directory (:455-466).  device = CUDA ordinal. */
int gap_potential_filename_initialise(gap_potential** pot, const char* args_str, const char* param_filename, int device);

/* potential_initialise with param_str (Potential.f95:499) -> IPModel_GAP_Initialise_str
 * (src/Potentials/IPModel_GAP.f95:149): param_str is the entire XML text; base_dir (may be NULL = ".")
 * is where sparseX_filename side files are looked up. */
int gap_potential_initialise(gap_potential** pot, const char* args_str, const char* param_str, const char* base_dir, int device);

/* finalise (IPModel_GAP_Finalise, IPModel_GAP.f95:194) */
void gap_potential_finalise(gap_potential* pot);

/* cutoff(pot) (Potential.f95:1044 -> IP_cutoff, IP.f95:704-705) */
double gap_potential_cutoff(const gap_potential* pot);

/* Print (IPModel_GAP_Print, IPModel_GAP.f95:952): writes a description into buf (NUL terminated, truncated to n) */
int gap_potential_print(const gap_potential* pot, char* buf, size_t n);

/* Data-parallel partition = the reference's MPI atom mask (descriptor_atomic_MPI_setup, descriptors.f95:1036-1051),
 * as contiguous blocks of central atoms: this handle evaluates only centres [rank*N/n_ranks, (rank+1)*N/n_ranks).
 * Outputs are then PARTIAL sums; the host reduces them (the reference's sum_in_place calls, IPModel_GAP.f95:538-556). */
int gap_potential_set_partition(gap_potential* pot, int rank, int n_ranks);

/* ---- the reduction over ranks, inside the library (replaces the reference's MPI_context + the five sum_in_place calls,
 * src/libAtoms/MPI_context.f95:668-694, src/Potentials/IPModel_GAP.f95:538-556; the mpi argument of IPModel_GAP_Calc).
 * One process per GPU.  Rank 0 calls gap_comm_get_unique_id and the HOST distributes the 128 bytes to the other ranks
 * (MPI_Bcast in a Fortran / LAMMPS host, torch.distributed in quip_b200.ShardedPotential); every rank then calls
 * gap_potential_set_comm (collective: NCCL communicator over NVLink / NVSwitch).  It also sets the partition (rank, n_ranks).
 * From then on gap_potential_calc, gap_potential_calc_device[_enqueue], gap_md_run and gap_md_run_device return TOTALS on every
 * rank: the packed [E | virial | F(3,N)] partials (and local_e / local_virial when requested) are summed on the evaluation's
 * stream by ncclAllReduce or, for latency-bound sizes, by a peer-memory kernel over NVLink P2P (one-shot pull on 2-3 ranks, low-latency
 * reduce-scatter + all-gather by push from 4 ranks on; csrc/comm.cu;
 * GAP_B200_P2P=0 forces NCCL; GAP_B200_P2P_MAX_BYTES, default 8 MiB, is the largest buffer the peer kernel takes; a rank that waits longer
 * than GAP_B200_P2P_TIMEOUT_S, default 60, for the others' partials reports an error instead of hanging; GAP_B200_P2P_LL_MIN_RANKS, default 4,
 * and GAP_B200_P2P_LL_MAX_DOUBLES, default 2^20, bound the use of the low-latency variant).  The peer kernels' blocks spin on data written by the
 * other GPUs, so they assume that the ranks run the same sequence of evaluations and that each rank's GPU is not saturated by unrelated work.
 * n_ranks = 1 removes the communicator. */
#define GAP_COMM_ID_BYTES 128
int gap_comm_get_unique_id(char* id /* GAP_COMM_ID_BYTES */);
int gap_potential_set_comm(gap_potential* pot, const char* id /* GAP_COMM_ID_BYTES */, int rank, int n_ranks);
/* rank / n_ranks of the handle and the transport of its last reduction ("nccl", "p2p" or "none") */
int gap_potential_comm_info(const gap_potential* pot, int* rank, int* n_ranks, char* transport, size_t n);
/* the last peer-memory reduction on this rank, from %globaltimer inside the kernel.  One-shot kernel: microseconds it waited for the other
 * ranks' partials (the skew between the ranks' evaluations) and microseconds of the sum phase itself.  Low-latency kernel: microseconds of
 * its push phase, and of everything after it (polling for the other ranks' data, summing, publishing and collecting the totals) */
int gap_potential_comm_timing(const gap_potential* pot, double* wait_us, double* sum_us);

/* calc(pot, at, energy, force, virial, local_energy, local_virial, args_str) (Potential.f95:803 ->
 * IPModel_GAP_Calc, IPModel_GAP.f95:233), including the neighbour-list build the reference does in
 * potential_calc (:844-859 -> calc_connect, src/libAtoms/Connection.f95:1035).  All pointers are HOST pointers; arrays the
 * caller has page-locked (cudaHostAlloc / cudaHostRegister) are transferred in place, pageable ones through a staging buffer.
 * With a communicator every rank passes the whole configuration and receives the totals; a rank that does not need an
 * array (e.g. the forces on ranks other than 0) passes NULL and skips that device-to-host copy. */
int gap_potential_calc(gap_potential* pot, int N, const double* pos, const int* Z, const double* lattice, const int* pbc,
                       const char* args_str, double* energy, double* local_e, double* force, double* virial,
                       double* local_virial);

/* at%cutoff_skin (calc_connect's skin, src/libAtoms/Connection.f95:1085-1128; the quip CLI's cutoff_skin=0.5 default, src/Programs/quip.f95:217,
 * 343-345): with skin > 0 the neighbour list of gap_potential_calc / _calc_device / gap_md_run* is built out to cutoff + skin and REUSED
 * while no atom has moved more than skin / 2 since the build (same N, partition, lattice, pbc); pairs beyond a descriptor's own cutoff
 * are dropped by the descriptor kernels, which recompute every distance from the current positions (the reference's calc_dists).
 * Results are identical to a rebuild at every call.  skin = 0 (default): rebuild every call.  _connect_stats counts both outcomes. */
int gap_potential_set_cutoff_skin(gap_potential* pot, double cutoff_skin);
int gap_potential_connect_stats(const gap_potential* pot, long* n_rebuilds, long* n_reuses);

/* Run-to-run reproducible forces.  By default the SOAP force scatter (IPModel_GAP.f95:482, f(:, ii(n)) -= f_gp) adds with FP64 atomics: the
 * summation order, hence the last bits of the forces, varies between runs (as it does between the threads of the reference's OpenMP
 * reduction).  on = 1: every pair force is stored at the slot of its neighbour-list entry and the slots are summed per receiving atom in
 * slot order with a fixed reduction tree (one stable radix sort of the slots per list build): bitwise identical forces for identical
 * inputs.  Energies and virials are reduced in a fixed order in both modes; local_virial and the atom-mask scatter keep atomics. */
int gap_potential_set_deterministic(gap_potential* pot, int on);

/* ---- optional inputs / outputs of IPModel_GAP_Calc that the reference passes through the Atoms object and the calc
 * args string (src/Potentials/IPModel_GAP.f95:324-337, 344-346, 462-488, 558-573).  They are requested with the SAME
 * keys in args_str and fetched after the calc:
 *   atom_mask_name=NAME            only atoms with mask != 0 are centres of descriptors and receive e0; the mask itself (the
 *                                  logical Atoms property NAME in the reference) is supplied with gap_potential_set_atom_mask.
 *                                  As in the reference it cannot be combined with an active partition (:375-378).
 *   energy_per_coordinate=NAME     -> gap_potential_get_energy_per_coordinate: sum of e_i * covariance_cutoff per GP
 *                                  coordinate (:462), before e0 and E_scale
 *   local_gap_variance=NAME [gap_variance_regularisation=0.001]
 *                                  -> gap_potential_get_local_gap_variance: per-atom predictive variance (:464-469; the sparse
 *                                  covariance k_mm is factorised on first use, gp_predict.f95:3970-4085, cuSOLVER) and, when
 *                                  forces or virials were requested, its gradient gap_variance_gradient(3,N) (:484-487).
 *                                  A negative variance is an error, as in gp_predict.f95:3877.
 * With a partition active the fetched arrays are this rank's partial sums (the reference's sum_in_place, :545-549). */
int gap_potential_set_atom_mask(gap_potential* pot, int N, const int* mask /* N logicals; NULL = no mask */);
/* residue ids of the atoms: the integer Atoms property a distance_2b descriptor names with resid_name= for only_intra / only_inter
 * (src/GAP/descriptors.f95:1771-1790, 4660-4668, 4735-4738) */
int gap_potential_set_resid(gap_potential* pot, int N, const int* resid /* N ints; NULL = none */);
int gap_potential_get_energy_per_coordinate(gap_potential* pot, double* energy_per_coordinate /* n_coordinate */);
int gap_potential_get_local_gap_variance(gap_potential* pot, int N, double* local_gap_variance /* N */,
                                         double* gap_variance_gradient /* 3*N or NULL */);

/* Same, GPU-resident: d_pos/d_Z are DEVICE pointers, results stay on the device.
 *   d_packed  : device double[10 + 3*N] = [ E | virial(9) | F(3,N) ]  (the buffer the host all-reduces when the
 *               partition is active: one collective instead of the reference's five); never NULL
 *   d_local_e : device double[N] or NULL ; d_local_virial : device double[9*N] or NULL
 *   want_grad : 0 = energy only
 *   stream    : cudaStream_t (as void*) the work is enqueued on; NULL = the handle's own stream.
 * The neighbour list is sized speculatively from the previous call with the same N and partition and nothing is read
 * back in mid-stream; the entry count is verified when the stream has drained, so from the second call on the
 * function returns after the work has COMPLETED (and transparently repeats the evaluation if the list overflowed).
 * The first call for a given N synchronises once after the neighbour count and returns after enqueueing the rest. */
int gap_potential_calc_device(gap_potential* pot, int N, const double* d_pos, const int* d_Z, const double* lattice,
                              const int* pbc, const char* args_str, int want_grad, double* d_packed, double* d_local_e,
                              double* d_local_virial, void* stream);

/* DynamicalSystem_run (src/Potentials/Potential.f95:2304-2369) for plain NVE dynamics: n_steps of velocity Verlet
 * (advance_verlet1 / advance_verlet2, src/libAtoms/DynamicalSystem.f95:1814, 2159; no thermostat, barostat or constraints),
 * forces from this potential, the neighbour list rebuilt on the device every step.  Positions, velocities and
 * accelerations stay resident on the GPU for the whole run; only the step energies come back.
 *   pos, velo : host double[3*N], in = initial state, out = final state (Angstrom, Angstrom/fs)
 *   mass      : host double[N] in QUIP units (amu * MASSCONVERT, src/libAtoms/Units.f95:68)
 *   epot,ekin : host double[n_steps+1] or NULL: potential / kinetic energy after the initial evaluation and each step */
int gap_md_run(gap_potential* pot, int N, double* pos, double* velo, const int* Z, const double* mass, const double* lattice, const int* pbc,
               double dt, int n_steps, const char* args_str, double* epot, double* ekin);

/* Same dynamics on DEVICE-resident state, for runs partitioned over several GPUs (BASELINE config C: MD with per-step
 * neighbour-list rebuild at 1/2/4/8 GPUs).  Every rank holds the whole state (d_pos, d_velo, d_mass, d_Z: device pointers,
 * updated in place) and evaluates its block of centres (gap_potential_set_partition); after each evaluation has been
 * enqueued the library calls reduce(reduce_ctx, stream), which must enqueue ON THAT STREAM the sum of d_packed[10 + 3*N]
 * over the ranks (the reference's sum_in_place calls, IPModel_GAP.f95:538-556; quip_b200.ShardedPotential passes an NCCL
 * all-reduce).  All ranks then integrate all atoms with identical forces, so the replicas stay bit-identical and no
 * position exchange is needed.  reduce may be NULL on a single rank. */
typedef void (*gap_reduce_fn)(void* reduce_ctx, void* stream);
int gap_md_run_device(gap_potential* pot, int N, double* d_pos, double* d_velo, const int* d_Z, const double* d_mass, const double* lattice,
                      const int* pbc, double dt, int n_steps, const char* args_str, double* d_packed, gap_reduce_fn reduce, void* reduce_ctx,
                      double* epot, double* ekin, void* stream);

/* ---- LAMMPS `pair_style quip` ABI: the three bind(c) symbols of src/Potentials/quip_lammps_wrapper.f95 (:24, :30-56,
 * :158-168) with identical names and argument lists (all by reference), so that LAMMPS' pair_quip.cpp links against
 * libgapb200.so instead of libquip.  The neighbour list is the caller's (full list of the local atoms, 1-based neighbour
 * indices, periodic images as explicit ghost atoms); only the local atoms are centres; forces on ghosts are returned. */
int quip_lammps_api_version(void);
void quip_lammps_potential_initialise(int* quip_potential, int* n_quip_potential, double* quip_cutoff, char* quip_file, int* n_quip_file,
                                      char* quip_string, int* n_quip_string);
void quip_lammps_wrapper(int* nlocal, int* nghost, int* atomic_numbers, int* lmptag, int* inum, int* sum_num_neigh, int* ilist, int* quip_num_neigh,
                         int* quip_neigh, double* lattice, int* quip_potential, int* n_quip_potential, double* quip_x, double* quip_e,
                         double* quip_local_e, double* quip_virial, double* quip_local_virial, double* quip_force);

/* Two-phase form of gap_potential_calc_device for callers that enqueue more work behind the evaluation (e.g. the NCCL
 * all-reduce of d_packed) before they synchronise: _enqueue only enqueues; after the caller has synchronised the stream,
 * _verify reports in *repeat whether the speculatively sized neighbour list overflowed (1 = call _enqueue again; the
 * second attempt sizes the list exactly). */
int gap_potential_calc_device_enqueue(gap_potential* pot, int N, const double* d_pos, const int* d_Z, const double* lattice, const int* pbc,
                                      const char* args_str, int want_grad, double* d_packed, double* d_local_e, double* d_local_virial,
                                      void* stream);
int gap_potential_calc_device_verify(gap_potential* pot, int* repeat);

/* F77-style one-shot entry point, same argument list as quip_wrapper_simple_
 * (src/Potentials/quip_unified_wrapper.f95:311-332) plus the XML file name; pbc = T T T. */
int gap_b200_wrapper_simple(const char* param_filename, const int* N, const double* lattice, const int* Z, const double* pos,
                            double* energy, double* force, double* virial);

/* ---- stage-level entry points (descriptors_wrapper.f95 analogue; used by the parity tests) ---- */

/* calc_connect (Connection.f95:1035) with an explicit cutoff; returns the number of list entries in *n_entries.
 * The list is the FULL list (both i->j and j->i), atom indices 0-based. */
int gap_calc_connect(gap_potential* pot, int N, const double* pos, const double* lattice, const int* pbc, double cutoff,
                     int* n_entries);
/* copy out the last list: offsets[N+1], neighbour j[n], shift[3*n], distance[n] (host pointers) */
int gap_get_connect(gap_potential* pot, int* offsets, int* j, int* shift, double* distance);

/* descriptor_calc for SOAP coordinate i_coord (0-based): two-call protocol.  With x == NULL only *n_desc and *d are
 * returned; otherwise x[n_desc*d] (row per centre) and ci[n_desc] (0-based centre atom) are filled. */
int gap_descriptor_calc(gap_potential* pot, int i_coord, int N, const double* pos, const int* Z, const double* lattice,
                        const int* pbc, int* n_desc, int* d, double* x, int* ci);

/* gp_predict for coordinate i_coord on n descriptor vectors x[n*d] (gpCoordinates_Predict, gp_predict.f95:3627):
 * e[n], grad[n*d] (grad may be NULL). dot_product coordinates only. */
int gap_gp_predict(gap_potential* pot, int i_coord, int n, const double* x, double* e, double* grad);

int gap_potential_n_coordinate(const gap_potential* pot);
/* number of CUDA kernels this handle has launched since initialise */
long gap_potential_launch_count(const gap_potential* pot);
/* device milliseconds of the last calc, by stage: [0] connect [1] soap forward [2] covariance GEMM-1 [3] covariance
 * GEMM-2 [4] soap adjoint + scatter [5] distance_2b [6] everything else (centre selection, memsets, energy rows,
 * totals) [7] whole calc.  CUDA events on the stream the calc was enqueued on; waits for the last calc. */
int gap_potential_last_timings(gap_potential* pot, double* ms8);
/* The per-stage events are instrumentation (system_timer in the reference, e.g. IPModel_GAP.f95:428): recorded only after
 * gap_potential_set_timing(pot, 1); off by default, gap_potential_last_timings then returns zeros.  on = 2 records only the
 * three events that bracket the two covariance GEMMs (slots [2] and [3]; the other slots stay zero). */
int gap_potential_set_timing(gap_potential* pot, int on);

/* Host-only: parse a model exactly as initialise would (XML + descriptor strings + SOAP radial-basis set-up) and
 * write a text description with all derived numbers (%.17g) into buf.  Needs no GPU; used to check the loader. */
int gap_model_describe(const char* args_str, const char* param_str, const char* base_dir, char* buf, size_t n);

const char* gap_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
