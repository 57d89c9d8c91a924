#!/usr/bin/env python
"""bench.py -- GAP-SOAP energy + force + virial throughput (atoms/s) of the B200 path, with its roofline, a parity check against
the CPU oracle on the very step that is timed, and the CPU baseline timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--named-configs auto|none|C|D|all]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Headline workload (BASELINE.json configs[1], SURVEY.md 8(d) "A"): Si diamond, 8x8x8 cubic cells = 4,096 atoms per GPU, rattled
0.05 A, SOAP n_max=8 l_max=8 cutoff 5 A, zeta=4, 2,000 random-init sparse points; one step = one full E+F+V evaluation
INCLUDING the neighbour-list build.  At N GPUs the cell is 8 x 8 x 8N (4,096 N atoms, weak scaling): positions are
replicated, every rank evaluates its block of centres, and the one exchange is the sum of [E | virial | F] over the ranks,
done inside libgapb200.so (gap_potential_set_comm: NCCL, or the one-shot NVLink peer-memory kernel for this size).

value   : inputs resident in HBM, timed with CUDA events per step (L2 flushed between steps), max over ranks.
e2e     : through the host-pointer C ABI (gap_potential_calc; page-locked host arrays, H2D and D2H inside the timed region), wall clock.
roofline: the dominant kernel of the step (by CUDA-event time on the launching stream).
parity  : the GPU result of the benchmarked configuration against the oracle's whole evaluation (rank 0), the neighbour list
          against the oracle's, and at N > 1 the reduced result against an unpartitioned evaluation on rank 0's GPU.  The run FAILS
          if the north-star tolerances are exceeded (1e-8 eV/atom, 1e-6 eV/A, 1e-6 eV).
named_configs: BASELINE configs[3] (C: 262,144-atom amorphous carbon, MD with the list rebuilt every step, strong scaling) at every
          N, configs[4] (D: 1,048,576-atom Si slab, E/F/V sharded) at N = 1 and 8, configs[2] (B: SiC, distance_2b + SOAP) at N = 1,
          measured in the same process.
cpu_baseline / --impl reference: the CPU restatement of QUIP's algorithm (oracle/, OpenMP over atoms) on the host cores.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GAP-SOAP energy+force+virial atoms/sec (incl. neighbour-list build)"
UNIT = "atoms/s"
CELLS = 8
# (n_max, l_max, n_species, sparse points per SOAP coordinate, neighbours per centre) of the named shapes
SHAPES = {"A": (8, 8, 1, 2000, 28.0), "B": (10, 6, 2, 4000, 50.0), "C": (8, 8, 1, 9000, 104.0), "D": (12, 8, 1, 8000, 27.0)}
TOL_E_PER_ATOM, TOL_F, TOL_V = 1e-8, 1e-6, 1e-6  # BASELINE.json north_star


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML polled every millisecond from a thread (the
    default timed region lasts ~15 ms, far below nvidia-smi's sampling period); falls back to `nvidia-smi -lms` without pynvml."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mask, self.power, self.max_mhz = [], 0, [], None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            reasons = sorted(k for k, b in self.BITS.items() if self.mask & b)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(self.sm), "power_w": float(np.median(self.power)) if self.power else None, "source": "nvml, 1 ms polling"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


CONFIG = "A"  # --config: A is the bench line (BASELINE configs[1]); B, C, D select another named shape for the headline legs (exploration)


def build_workload(tmp, n_gpus, descriptor_fn, config=None):
    """The named configuration (A: x n_gpus along z for weak scaling) + its random-init model, written as a reference-format GAP XML."""
    from quip_b200 import synthetic as syn
    from quip_b200.gap_xml import write_gap_xml

    config = config or CONFIG
    if config == "B":
        return syn.build_config_B(tmp, descriptor_fn)
    if config == "C":
        return syn.build_config_C(tmp, descriptor_fn, n_src=12288)
    if config == "D":
        return syn.build_config_D(tmp, descriptor_fn)
    n_max, l_max, n_spec, M, _ = SHAPES["A"]
    atoms = syn.si_diamond(CELLS, CELLS, CELLS * n_gpus, seed=1)
    src = syn.si_diamond(CELLS, CELLS, CELLS, rattle=0.08, seed=101, strain=0.01)
    X = descriptor_fn(syn.SOAP_A, src)
    coord = syn.random_soap_coordinate(syn.SOAP_A, X, M, delta=1.0, zeta=4.0, seed=101)
    xml = write_gap_xml(os.path.join(tmp, "gap_config_A.xml"), [coord], e0={14: -158.54496821}, label="GAP_b200_config_A")
    return atoms, xml


def workload_name(n_gpus, config=None):
    config = config or CONFIG
    if config != "A":
        return {"B": "SiC 32768 atoms, 3x distance_2b + 2x SOAP n_max=10 l_max=6, 4000 sparse points per species (strong scaling over %d GPUs)",
                "C": "amorphous carbon 262144 atoms, SOAP cutoff 5.5 n_max=8 l_max=8, 9000 sparse points (strong scaling over %d GPUs)",
                "D": "Si slab 1048576 atoms, SOAP n_max=12 l_max=8, 8000 sparse points (strong scaling over %d GPUs)"}[config] % n_gpus
    return ("Si diamond %d atoms (8x8x%d cells, rattled 0.05 A), SOAP n_max=8 l_max=8 cutoff=5.0 zeta=4, 2000 sparse points, "
            "single-step E/F/V incl. neighbour list" % (4096 * n_gpus, 8 * n_gpus))


def config_dict(n_gpus, n_atoms):
    """The `config` object, IDENTICAL on both arms (the driver compares them): what is computed, and how each arm is run and timed."""
    from quip_b200 import synthetic as syn

    n_max, l_max, n_spec, M, _ = SHAPES[CONFIG]
    return {"workload": workload_name(n_gpus), "atoms": int(n_atoms), "atoms_per_gpu": int(n_atoms // n_gpus), "sparse_points": M,
            "descriptor_dim": syn.soap_dimension(n_max, l_max, n_spec),
            "parallelism": "B200 arm: centre-block x%d, positions replicated, one in-library reduction of [E|virial|F] (NCCL / NVLink peer memory); "
                           "reference arm: OpenMP over atoms on the host cores" % n_gpus,
            "l2": "B200 arm: flushed between timed steps (512 MiB memset); reference arm: n/a (CPU)",
            "timing": "B200 arm: CUDA events per step on the launching stream around the enqueued step (kernels + reduction + energy read-back), one "
                      "host synchronise + neighbour-list verification per step follows the closing event, max over ranks (at N > 1 the ranks are aligned by a "
                      "one-element all-reduce on the stream between the L2 flush and the opening event); reference arm: wall "
                      "clock of calc_connect + soap_calc + gp_predict/scatter (omp_get_wtime around the phases the reference times)"}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle = restatement of the reference's algorithm; the reference itself is Fortran and cannot be built here)
# ------------------------------------------------------------------------------------------------------------------
def host_threads():
    """All host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(om, atoms, n_centres):
    """One oracle pass over centres [0, n_centres): returns (wall seconds of the pass, wall seconds of its serial calc_connect).  (The
    oracle's descriptor / gp_predict timers are summed over the OpenMP threads; the per-centre share is taken from the wall clock.)"""
    t = time.perf_counter()
    o = om.calc(atoms, force=True, virial=True, first=0, last=n_centres, nthreads=host_threads())
    return time.perf_counter() - t, float(o["timings"][0])


def whole_step_seconds(N, n_centres, wall, t_connect):
    """Seconds of ONE whole evaluation of N atoms implied by a pass over n_centres of its centres: the serial neighbour list of all
    atoms is paid once (calc_connect, Connection.f95:1060), the per-centre phases (soap_calc, descriptors.f95:7757; gp_predict +
    scatter, IPModel_GAP.f95:428) scale with the number of centres."""
    return float(t_connect + (wall - t_connect) * (N / float(n_centres)))


def cpu_baseline_leg(xml, atoms, budget_s=12.0):
    """About budget_s seconds of CPU work on the same workload: a probe sizes the sample; small configurations are repeated whole."""
    from oracle import oracle as orc

    om = orc.Model(xml)
    cores = host_threads()
    N = len(atoms)
    flavours, tot_all = {}, 0.0
    for name in ("loops", "openblas_dgemv"):  # the two per-atom products of gp_predict as vectorised loops / as dgemv from scipy's OpenBLAS
        if name == "openblas_dgemv" and not orc.use_openblas(True):
            continue
        n = min(N, 256)
        rate = n / cpu_sample(om, atoms, n)[0]
        per_pass = int(min(N, max(n, rate * budget_s / 2)))
        passes = int(max(1, min(64, round(rate * budget_s / 2 / per_pass))))
        tot, secs = 0.0, 0.0
        for _ in range(passes):
            t, tc = cpu_sample(om, atoms, per_pass)
            tot += t
            secs += whole_step_seconds(N, per_pass, t, tc)
        flavours[name] = {"value": N * passes / secs, "passes": passes, "centres_per_pass": per_pass}
        tot_all += tot
    orc.use_openblas(False)
    best = max(flavours, key=lambda k: flavours[k]["value"])
    return {"value": flavours[best]["value"], "unit": UNIT, "cores": cores, "kind": "port", "flavour": best,
            "flavours": {k: v["value"] for k, v in flavours.items()},
            "sample": "%d pass(es) over %d of %d centres of the same configuration per flavour, %.1f s in total; a whole step = serial neighbour list of "
                      "all atoms (as calc_connect) + per-centre phases scaled to all centres; OpenMP over atoms with %d threads; oracle/gap_oracle.c "
                      "with the gp_predict products as vectorised loops and as OpenBLAS dgemv (the faster one is `value`); not the gfortran binary"
                      % (flavours[best]["passes"], flavours[best]["centres_per_pass"], N, tot_all, cores)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: the Fortran cannot be compiled in this image) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    orc.build()
    n_gpus = args.gpus
    with tempfile.TemporaryDirectory() as tmp:
        atoms, xml = build_workload(tmp, n_gpus, lambda desc, at: orc.soap_descriptor(desc, at)["data"])
        om = orc.Model(xml)
        cores = host_threads()
        N = len(atoms)
        n = min(N, 512)
        cpu_sample(om, atoms, n)  # (first touch: thread pool, page faults)
        rate = max(n / cpu_sample(om, atoms, n)[0] for _ in range(2))
        flavour = "loops"
        if orc.use_openblas(True):  # gp_predict's two products as dgemv from scipy's OpenBLAS: keep whichever flavour is faster here
            cpu_sample(om, atoms, n)
            rate_blas = max(n / cpu_sample(om, atoms, n)[0] for _ in range(2))
            if rate_blas > rate:
                rate, flavour = rate_blas, "openblas_dgemv"
            else:
                orc.use_openblas(False)
        # each step = a bounded sample of centres sized so that warmup + steps take about two minutes at most; a step that covers
        # every centre is a whole evaluation, a smaller one is scaled to the whole evaluation phase by phase (the list is paid once)
        per_step = int(max(64, min(N, rate * 120.0 / max(1, args.steps + args.warmup))))
        for _ in range(args.warmup):
            cpu_sample(om, atoms, per_step)
        secs = [whole_step_seconds(N, per_step, *cpu_sample(om, atoms, per_step)) for _ in range(args.steps)]
    tot = float(np.sum(secs))
    value = N * args.steps / tot
    sample = ("%d of %d centres per step; step time = calc_connect of all atoms (serial) + (soap_calc + gp_predict/scatter of the sample) x %d/%d; "
              "%d OpenMP threads; gp_predict products: %s" % (per_step, N, N, per_step, cores, flavour))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak" if CONFIG == "A" else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(n_gpus, N),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "QUIP's Fortran GAP path cannot be compiled in this image (no Fortran compiler); this is oracle/gap_oracle.c, a C/OpenMP "
                    "restatement with the reference's loop structure (serial linked-cell list, forward-mode grad_data, the two per-atom BLAS-2 "
                    "products of gp_predict as %s -- the faster of the two flavours in a probe of this run)"
                    % ("dgemv calls into scipy's OpenBLAS" if flavour == "openblas_dgemv" else "vectorised loops")}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 4096^3, best of 5: the FP64 (DMMA) yardstick MEASURED_PEAKS.json does not carry."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def measure_cublas_same_shape(torch, dev, nc, d, M):
    """cuBLAS DGEMM on the covariance stage's own shapes (C = X S^T: nc x M x d, G = A S: nc x d x M), best of 5 each: what the library
    reaches on THIS problem size (without the fused kernel epilogue), next to the 4096^3 figure."""
    X = torch.randn(nc, d, dtype=torch.float64, device=dev)
    S = torch.randn(M, d, dtype=torch.float64, device=dev)
    A = torch.randn(nc, M, dtype=torch.float64, device=dev)
    best = [1e30, 1e30]
    for k, fn in enumerate((lambda: torch.matmul(X, S.t()), lambda: torch.matmul(A, S))):
        fn()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best[k] = min(best[k], e0.elapsed_time(e1))
    return {"gemm1_ms": best[0], "gemm2_ms": best[1], "pair_tflops": 4.0 * nc * d * M / ((best[0] + best[1]) * 1e-3) / 1e12}


DMMA_PROBE_TFLOPS = 37.1  # register-only DMMA loop on a B200 (tools/fp64_pipes.cu, profiles/r01e_fp64_pipes.txt): the pipe's own ceiling


class Ctx:
    """Process-group plumbing shared by the legs."""

    def __init__(self, torch, dist, world, rank, local, dev, real_world):
        self.torch, self.dist, self.world, self.rank, self.local, self.dev, self.real_world = torch, dist, world, rank, local, dev, real_world
        self.share = None

    def barrier(self):
        if self.real_world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.real_world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def shared_dir(self):
        """One scratch directory for the whole job: rank 0 creates it, every rank learns its name (the models are built once, by rank 0)."""
        if self.share is None:
            box = [tempfile.mkdtemp(prefix="gapb200_bench_") if self.rank == 0 else None]
            if self.real_world > 1:
                self.dist.broadcast_object_list(box, src=0)
            self.share = box[0]
        return self.share

    def build_shared(self, tag, builder):
        """(atoms, xml) built ONCE by rank 0 (structure, sparse points through the CUDA descriptor kernels, reference-format XML) and
        read by the other ranks from the shared directory."""
        from quip_b200 import Atoms

        d = os.path.join(self.shared_dir(), tag)
        meta = os.path.join(d, "atoms.npz")
        if self.rank == 0:
            os.makedirs(d, exist_ok=True)
            atoms, xml = builder(d)
            np.savez(meta, numbers=atoms.numbers, positions=atoms.positions, cell=atoms.cell, pbc=np.asarray(atoms.pbc), xml=np.array(xml))
        self.barrier()
        z = np.load(meta)
        return Atoms(z["numbers"], z["positions"], z["cell"], z["pbc"]), str(z["xml"])


def descriptor_builder(device, config):
    """descriptor_fn for synthetic.build_config_*: SOAP vectors of the sparse-point source structure from the CUDA kernels."""
    from quip_b200 import Potential
    from quip_b200 import synthetic as syn

    descs = {"A": [(syn.SOAP_A, syn.soap_dimension(8, 8))], "B": [(syn.SOAP_B % 6, syn.soap_dimension(10, 6, 2)), (syn.SOAP_B % 14, syn.soap_dimension(10, 6, 2))],
             "C": [(syn.SOAP_C, syn.soap_dimension(8, 8))], "D": [(syn.SOAP_D, syn.soap_dimension(12, 8))]}[config]
    which = {dsc: k for k, (dsc, _) in enumerate(descs)}

    def build(tmp, n_gpus):
        boot = syn.bootstrap_xml(os.path.join(tmp, "boot.xml"), descs)
        bp = Potential("", param_filename=boot, device=device)
        try:
            return build_workload(tmp, n_gpus, lambda desc, at: bp.descriptor_calc(at, which[desc])[0], config)
        finally:
            bp.finalise()
    return build


def pinned_atoms(atoms):
    """The same configuration with positions and atomic numbers in page-locked host memory (what the e2e leg hands to the C ABI)."""
    from quip_b200 import Atoms
    from quip_b200.potential import pinned_copy

    return Atoms(pinned_copy(np.asarray(atoms.numbers, dtype=np.int32)), pinned_copy(np.asarray(atoms.positions, dtype=np.float64)), atoms.cell, atoms.pbc)


def parity_block(atoms, xml, reduced, device, world, check_list=True, sample=None, soap_only=False):
    """Rank 0: the benchmarked configuration against the oracle (whole evaluation, or `sample` centres' local energies when the whole
    one is out of reach), the neighbour list against the oracle's, and the reduced N-rank result against an unpartitioned evaluation."""
    from oracle import oracle as orc
    from quip_b200 import Potential

    N = len(atoms)
    out = {}
    t0 = time.perf_counter()
    if soap_only:  # distance_2b splits each pair's energy between both ends, so its partial local energies are not comparable centre by
        # centre when the oracle only visits a sample of centres: compare the SOAP coordinates' share (the GPU side sums only_descriptor runs)
        spec = orc.load_gap_xml(xml)
        spec["coordinates"] = [c for c in spec["coordinates"] if c["descriptor"].split()[0] == "soap"]
        om = orc.Model(model=spec)
    else:
        om = orc.Model(xml)
    if sample is None:
        o = om.calc(atoms, force=True, virial=True, nthreads=host_threads())
        out.update(dE_per_atom=abs(reduced["energy"] - o["energy"]) / N, max_dF=float(np.abs(reduced["force"] - o["force"]).max()),
                   max_dvirial=float(np.abs(reduced["virial"] - o["virial"]).max()), oracle="whole evaluation of all %d atoms" % N)
        ok = out["dE_per_atom"] <= TOL_E_PER_ATOM and out["max_dF"] <= TOL_F and out["max_dvirial"] <= TOL_V
    else:
        first, last = sample
        o = om.calc(atoms, first=first, last=last, local_energy=True, force=False, virial=False, nthreads=host_threads())
        out.update(max_dlocal_e=float(np.abs(reduced["local_energy"][first:last] - o["local_energy"][first:last]).max()),
                   oracle="local energies of centres [%d, %d) (the oracle's neighbour list covers all %d atoms)" % (first, last, N))
        ok = out["max_dlocal_e"] <= TOL_E_PER_ATOM
    if check_list or world > 1:
        p1 = Potential("", param_filename=xml, device=device)  # unpartitioned handle on this GPU
        if check_list:
            def canon(off, j, s, d):
                i = np.repeat(np.arange(len(off) - 1), np.diff(off))
                key = np.lexsort((s[:, 2], s[:, 1], s[:, 0], j, i))
                return np.stack([i[key], j[key], s[key, 0], s[key, 1], s[key, 2]], axis=1), d[key]
            g, oc = canon(*p1.calc_connect(atoms, p1.cutoff())), canon(*orc.Connect(atoms, p1.cutoff()).arrays())
            out["nlist_identical"] = bool(g[0].shape == oc[0].shape and np.array_equal(g[0], oc[0]) and np.array_equal(g[1], oc[1]))
            out["nlist_entries"] = int(len(g[1]))
            ok = ok and out["nlist_identical"]
        if world > 1 and "force" in reduced:
            u = p1.calc(atoms, force=True, virial=True)
            sc_e, sc_f, sc_v = max(1.0, abs(u["energy"])), max(1.0, np.abs(u["force"]).max()), max(1.0, np.abs(u["virial"]).max())
            out["vs_unpartitioned"] = {"rel_dE": abs(reduced["energy"] - u["energy"]) / sc_e, "rel_max_dF": float(np.abs(reduced["force"] - u["force"]).max() / sc_f),
                                       "rel_max_dvirial": float(np.abs(reduced["virial"] - u["virial"]).max() / sc_v)}
            ok = ok and max(out["vs_unpartitioned"].values()) <= 1e-9
        p1.finalise()
    out["tolerances"] = {"dE_per_atom": TOL_E_PER_ATOM, "max_dF": TOL_F, "max_dvirial": TOL_V, "vs_unpartitioned_rel": 1e-9}
    out["ok"] = bool(ok)
    out["seconds"] = round(time.perf_counter() - t0, 2)
    return out


def _e0_of(atoms, xml):
    """per-atom e0 of the model (every only_descriptor run adds it once)"""
    from oracle import oracle as orc

    e0 = orc.load_gap_xml(xml)["e0"]
    return np.array([e0[int(z)] for z in atoms.numbers])


def roofline_block(st, nc, N, world, shape, fp64_peak, traffic=None):
    """Roofline of the dominant kernel from per-stage CUDA-event milliseconds (per step)."""
    from quip_b200 import synthetic as syn

    pk = peaks()
    n_max, l_max, n_spec, M, nn = shape
    d = syn.soap_dimension(n_max, l_max, n_spec)
    nlmK1 = (l_max + 1) ** 2 * n_max * n_spec
    n_soap = n_spec  # one SOAP coordinate per centre species; every centre meets exactly one
    kernels = {
        # HBM-side algorithmic bytes per centre: CSR read (8 B / entry) + pos/Z of the shell + x, X_lm, |p| written (forward);
        # CSR + x + g + X_lm read, forces written (adjoint)
        "soap_forward": {"bound": "hbm", "work": nc * (8.0 * nn + 28.0 * nn + 8.0 * (d + nlmK1 + 1)), "unit": "GB/s"},
        "soap_adjoint": {"bound": "hbm", "work": nc * (8.0 * nn + 28.0 * nn + 8.0 * (2 * d + nlmK1 + 1) + 24.0 * (nn + 1)), "unit": "GB/s"},
        "connect": {"bound": "hbm", "work": N * (24.0 + 4.0) * 2 + nc * nn * 8.0, "unit": "GB/s"},
        # FP64 tensor-core GEMMs: GEMM-1 and GEMM-2 are two launches of ONE kernel (k_dgemm_nt, two epilogues), ranked and reported
        # together: 4 d M flops per centre (F_cov)
        "k_dgemm_nt": {"bound": "tensor", "work": 4.0 * d * M * nc, "unit": "TFLOP/s", "launches": 2 * n_soap},
    }
    st = dict(st)
    st["k_dgemm_nt"] = st["cov_gemm1"] + st["cov_gemm2"]
    dom = max(kernels, key=lambda k: st.get(k, 0.0))
    kd = kernels[dom]
    secs = st[dom] * 1e-3
    if kd["bound"] == "tensor":
        achieved, peak, peak_src = kd["work"] / secs / 1e12, fp64_peak, "cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)"
    else:
        achieved, peak, peak_src = kd["work"] / secs / 1e9, pk["hbm_gbs"], pk["source"]
    pair = 4.0 * d * M * nc / (st["k_dgemm_nt"] * 1e-3) / 1e12 if st["k_dgemm_nt"] > 0 else 0.0
    return {"kernel": dom + (" (GEMM-1 + GEMM-2 launches of the covariance stage)" if dom == "k_dgemm_nt" else ""), "bound": kd["bound"],
            "achieved": achieved, "peak": peak, "unit": kd["unit"], "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "ms_per_launch": st[dom] / kd.get("launches", 1), "stage_ms": {k: round(v, 4) for k, v in st.items()},
            "fp64_dgemm_tflops_measured": fp64_peak, "cov_pair_tflops": pair,
            "frac_of_dmma_probe": pair / DMMA_PROBE_TFLOPS, "dmma_probe_tflops": DMMA_PROBE_TFLOPS}


def leg_static(ctx, args, config, atoms, xml, steps, warmup, with_cpu, full_parity):
    """One named configuration as a static E/F/V benchmark: value (resident), stage times, e2e (host pointers), parity."""
    torch = ctx.torch
    from quip_b200 import ShardedPotential

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    N = len(atoms)
    sp = ShardedPotential("", param_filename=xml, device=ctx.local, rank=rank, world_size=world)
    pot = sp.pot
    lat, pbc = atoms.lattice_fortran, atoms.pbc
    d_pos = torch.tensor(atoms.positions, dtype=torch.float64, device=dev)
    d_Z = torch.tensor(atoms.numbers, dtype=torch.int32, device=dev)
    d_packed = torch.empty(10 + 3 * N, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.int8, device=dev)  # > 126 MB L2

    # ---- value: HBM-resident inputs, CUDA events per step, L2 flushed between steps ----
    # everything below is enqueued on sp.stream (a real stream; events are recorded on the stream the kernels run on)
    torch.cuda.set_stream(sp.stream)
    pot.set_timing(2)  # timed region: only the events that bracket the covariance GEMMs (the roofline kernel)
    for _ in range(max(warmup, 3)):
        sp.calc_resident(N, d_pos, d_Z, lat, pbc, d_packed, want_grad=True)
    ctx.barrier()
    clocks = ClockSampler(ctx.local)
    if rank == 0:
        clocks.start()
    launches0 = pot.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    stage_sum = {}
    # N > 1: the ranks are aligned on the device after the L2 flush and before the step's opening event (a one-element all-reduce on the same
    # stream), as consecutive MD steps are by the previous step's reduction; without it a step would also time the other ranks' flushes
    align = torch.zeros(1, device=dev) if ctx.real_world > 1 else None
    ctx.barrier()
    t_wall = time.perf_counter()
    for k in range(steps):
        flush.zero_()
        if align is not None:
            ctx.dist.all_reduce(align)
        ev[k][0].record()
        sp.calc_resident_enqueue(N, d_pos, d_Z, lat, pbc, d_packed, want_grad=True)  # evaluation + reduction + energy read-back
        ev[k][1].record()
        assert sp.calc_resident_finish(), "speculative neighbour list overflowed on a static geometry"  # one sync + verify per step
        for name, ms in pot.last_timings().items():  # waits for this step's last kernel; per-stage CUDA events on the same stream
            stage_sum[name] = stage_sum.get(name, 0.0) + ms
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall
    launches = pot.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / steps
    value = N / (ms_per_step * 1e-3)
    resident = d_packed.cpu().numpy()

    # ---- the other stages' times, for the record: a separate instrumented pass (every stage event adds ~1 us to the step) ----
    pot.set_timing(1)
    stage_diag = {}
    red_wait, red_sum = [], []
    for k in range(steps):
        flush.zero_()
        if align is not None:
            ctx.dist.all_reduce(align)
        sp.calc_resident_enqueue(N, d_pos, d_Z, lat, pbc, d_packed, want_grad=True)
        assert sp.calc_resident_finish()
        for name, ms in pot.last_timings().items():
            stage_diag[name] = stage_diag.get(name, 0.0) + ms
        if world > 1 and not ctx.real_world == 1:
            ci = pot.comm_info()  # %globaltimer stamps of this step's peer reduction: wait for the other ranks' partials / sum phase
            red_wait.append(ci["last_wait_us"])
            red_sum.append(ci["last_sum_us"])
    for name in stage_diag:
        if name not in ("cov_gemm1", "cov_gemm2"):
            stage_sum[name] = stage_diag[name]
    ctx.barrier()
    del flush

    # ---- e2e: host-pointer C ABI on every rank, page-locked host arrays, H2D + D2H inside the timed region, wall clock.  Every rank
    #      uploads the whole configuration (positions are replicated) and receives E and the virial; rank 0 also reads all forces. ----
    pot.set_timing(False)  # the stage events are instrumentation of the legs above
    hat = pinned_atoms(atoms)
    from quip_b200.potential import pinned_copy
    h_force = pinned_copy(np.zeros((N, 3))) if rank == 0 else None
    for _ in range(3):
        r = sp.calc(hat, force=(rank == 0), virial=True, out_force=h_force)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = sp.calc(hat, force=(rank == 0), virial=True, out_force=h_force)
    torch.cuda.synchronize()
    e2e_value = N * steps / ctx.max_over_ranks(time.perf_counter() - t0)
    assert abs(r["energy"] - resident[0]) <= 1e-9 * abs(resident[0]), (r["energy"], resident[0])
    transport = pot.comm_info()["transport"]
    reduction = {"transport": transport}
    if red_wait:
        reduction.update(wait_us_mean_over_steps_max_over_ranks=ctx.max_over_ranks(float(np.mean(red_wait))),
                         sum_phase_us_mean_over_steps_max_over_ranks=ctx.max_over_ranks(float(np.mean(red_sum))),
                         note="peer-memory kernel, %globaltimer: wait = time between this rank's partial being ready and the last rank's "
                              "(the skew of the ranks' evaluations, not transfer time); sum = reading the partials of all ranks over NVLink")

    # ---- parity of the benchmarked configuration (rank 0; the other ranks wait) ----
    parity = None
    if full_parity is not None:
        if full_parity:
            if rank == 0:
                parity = parity_block(atoms, xml, r, ctx.local, world)
        else:  # sample of centres: local energies, reduced over the ranks inside the library (collective call on every rank)
            soap_only = config == "B"
            if soap_only:  # SOAP coordinates only (see parity_block): sum of the only_descriptor runs; every run adds e0 once
                ks = [k + 1 for k in range(pot.n_coordinate) if k >= 3]
                runs = [pot.calc(hat, local_energy=True, args_str="only_descriptor=%d" % k)["local_energy"] for k in ks]
                rl = {"local_energy": sum(runs) - (len(runs) - 1) * _e0_of(atoms, xml)}
            else:
                rl = pot.calc(hat, local_energy=True)
            if rank == 0:
                first = min(N - 8, 70000 if config == "D" else 5000)
                parity = parity_block(atoms, xml, rl, ctx.local, world, check_list=False, sample=(first, first + 8), soap_only=soap_only)
        ctx.barrier()
    sp.pot.finalise()
    del d_pos, d_Z, d_packed
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    st = {k: v / steps for k, v in stage_sum.items()}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram__bytes_read+write per launch from the committed ncu --set full capture
    if os.path.exists(tpath) and config == "A" and world == 1:
        traffic = json.load(open(tpath)).get("k_dgemm_nt")
    roofline = roofline_block(st, N // world, N, world, SHAPES[config], ctx.fp64_peak, traffic)
    n_max, l_max, n_spec, M_sp, _ = SHAPES[config]
    from quip_b200 import synthetic as syn
    if N // world <= 65536:
        same = measure_cublas_same_shape(torch, dev, N // world // n_spec, syn.soap_dimension(n_max, l_max, n_spec), M_sp)
        roofline["cublas_same_shape"] = same
        roofline["frac_of_cublas_same_shape"] = roofline["cov_pair_tflops"] / same["pair_tflops"]
    roofline["stage_ms_note"] = ("cov_gemm1 / cov_gemm2 / k_dgemm_nt: CUDA events inside the timed region; the other stages: a separate fully "
                                 "instrumented pass of the same steps")
    cpu = cpu_baseline_leg(xml, atoms) if with_cpu else None
    out = {"value": value, "ms_per_step": ms_per_step, "clocks": clk,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 28 * N, "d2h_bytes_per_step": 8 * (10 + 3 * N),
                   "note": "per rank: H2D of all positions + Z from page-locked memory; D2H of E + virial on every rank, of all forces on rank 0"},
           "gpu_launches": int(launches), "roofline": roofline, "wall_s_timed_region": t_wall, "energy_eV": float(resident[0]),
           "reduction_transport": transport, "reduction": reduction}
    if parity is not None:
        out["parity"] = parity
    if cpu is not None:
        out["cpu_baseline"] = cpu
    return out


def leg_md(ctx, atoms, xml, n_steps):
    """BASELINE configs[3]: NVE MD with the neighbour list rebuilt on the device every step (strong scaling over the ranks)."""
    torch = ctx.torch
    from quip_b200 import Atoms, ShardedPotential
    from quip_b200.potential import element_masses

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    N = len(atoms)
    m = element_masses(atoms.numbers)
    rng = np.random.default_rng(7)
    v0 = rng.normal(size=atoms.positions.shape) * np.sqrt(8.617385e-5 * 300.0 / m)[:, None]  # Maxwell, 300 K
    v0 -= (m[:, None] * v0).sum(axis=0) / m.sum()
    sp = ShardedPotential("", param_filename=xml, device=ctx.local, rank=rank, world_size=world)
    copy = lambda: Atoms(atoms.numbers, atoms.positions.copy(), atoms.cell, atoms.pbc)
    sp.run(copy(), v0, dt=1.0, n_steps=1)  # warm-up: buffers, communicator, peer mappings
    # a run of 0 steps = the state's trip to the GPU and back + the initial force evaluation: subtracted, so that the figure is the
    # cost of an MD step itself (integrator + list rebuild + evaluation + reduction + one host synchronisation)
    a0 = copy()
    ctx.barrier()
    t0 = time.perf_counter()
    sp.run(a0, v0, dt=1.0, n_steps=0)
    torch.cuda.synchronize()
    wall0 = ctx.max_over_ranks(time.perf_counter() - t0)
    a1 = copy()
    ctx.barrier()
    t0 = time.perf_counter()
    v1, ep, ek = sp.run(a1, v0, dt=1.0, n_steps=n_steps)
    torch.cuda.synchronize()
    wall = ctx.max_over_ranks(time.perf_counter() - t0)
    # replicas must stay bit-identical: spread of the final positions over the ranks
    spread = 0.0
    if ctx.real_world > 1:
        t = torch.tensor(a1.positions, device=dev)
        lo, hi = t.clone(), t.clone()
        ctx.dist.all_reduce(lo, op=ctx.dist.ReduceOp.MIN)
        ctx.dist.all_reduce(hi, op=ctx.dist.ReduceOp.MAX)
        spread = float((hi - lo).abs().max())
    # stage times and GEMM rate of one static evaluation of the initial configuration; parity sample of its local energies
    pot = sp.pot
    d_pos = torch.tensor(atoms.positions, dtype=torch.float64, device=dev)
    d_Z = torch.tensor(atoms.numbers, dtype=torch.int32, device=dev)
    d_packed = torch.empty(10 + 3 * N, dtype=torch.float64, device=dev)
    torch.cuda.set_stream(sp.stream)
    pot.set_timing(1)
    st = {}
    for _ in range(2):
        sp.calc_resident(N, d_pos, d_Z, atoms.lattice_fortran, atoms.pbc, d_packed, want_grad=True)
        st = pot.last_timings()
    pot.set_timing(False)
    rl = pot.calc(atoms, local_energy=True)  # collective: local energies reduced inside the library
    parity = None
    if rank == 0:
        parity = parity_block(atoms, xml, rl, ctx.local, world, check_list=False, sample=(1000, 1016))
    ctx.barrier()
    transport = pot.comm_info()["transport"]
    pot.finalise()
    del d_pos, d_Z, d_packed
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    etot = ep + ek
    t_steps = max(wall - wall0, 1e-9)
    return {"workload": workload_name(world, "C") + ", NVE velocity Verlet dt = 1 fs from 300 K, neighbour list rebuilt every step", "n_gpus": world,
            "atoms": N, "md_steps": n_steps, "ms_per_md_step": 1e3 * t_steps / n_steps, "atom_steps_per_s": N * n_steps / t_steps,
            "run_wall_s": wall, "run_of_0_steps_wall_s": wall0,
            "timing": "wall clock (max over ranks) of a %d-step run minus that of a 0-step run through the same call (ShardedPotential.run -> "
                      "gap_md_run_device): %d x (integrator + list rebuild + evaluation + reduction, one host synchronisation per step); the 0-step "
                      "run is the state's transfer to the GPU and back plus the initial force evaluation" % (n_steps, n_steps),
            "energy_drift_eV": float(np.abs(etot - etot[0]).max()), "ekin0_eV": float(ek[0]), "replica_spread": spread,
            "roofline": roofline_block(st, N // world, N, world, SHAPES["C"], ctx.fp64_peak), "reduction_transport": transport, "parity": parity}


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    real_world = world
    emulate = args.emulate_world if world == 1 else 0  # profiling aid: rank 0's share of a W-rank step on one GPU, no reduction
    if emulate:
        world = emulate
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if real_world > 1:
        # the bench prints ONE line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION prints it to stdout) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    if args.gpus != world and rank == 0:
        print("bench.py: --gpus %d but WORLD_SIZE=%d; using %d" % (args.gpus, world, world), file=sys.stderr)
    t_start = time.perf_counter()
    leg_s = {}
    ctx = Ctx(torch, dist, world, rank, local, dev, real_world)
    ctx.fp64_peak = measure_fp64_peak(torch, dev) if rank == 0 else 0.0

    # ---- headline: config A (or --config) ----
    build = descriptor_builder(local, CONFIG)
    atoms, xml = ctx.build_shared("headline_" + CONFIG, lambda d: build(d, n_gpus))
    N = len(atoms)
    steps, warmup = args.steps, max(args.warmup, 3)
    head = leg_static(ctx, args, CONFIG, atoms, xml, steps, warmup, with_cpu=(real_world == 1 and not args.no_cpu_baseline and not emulate),
                      full_parity=(None if (args.no_parity or emulate) else (N <= 65536)))

    leg_s["headline"] = round(time.perf_counter() - t_start, 1)
    # ---- the other named configurations of BASELINE.json, in the same process ----
    named = {}
    want = args.named_configs
    if emulate or CONFIG != "A":
        want = "none"
    do_C = want in ("C", "all") or want == "auto"
    do_D = want in ("D", "all") or (want == "auto" and n_gpus in (1, 8))
    do_B = want in ("B", "all") or (want == "auto" and n_gpus == 1)
    if do_C:
        buildC = descriptor_builder(local, "C")
        atomsC, xmlC = ctx.build_shared("named_C", lambda d: buildC(d, n_gpus))
        t_leg = time.perf_counter()
        named["C_md"] = leg_md(ctx, atomsC, xmlC, args.md_steps)
        leg_s["C_md"] = round(time.perf_counter() - t_leg, 1)
    if do_B:  # BASELINE configs[2]: SiC, 2 species, 3 x distance_2b + one SOAP (n_max=10 l_max=6, 4,000 sparse points) per centre species
        t_leg = time.perf_counter()
        buildB = descriptor_builder(local, "B")
        atomsB, xmlB = ctx.build_shared("named_B", lambda d: buildB(d, n_gpus))
        b_leg = leg_static(ctx, args, "B", atomsB, xmlB, 5, 3, with_cpu=False, full_parity=False)
        if rank == 0:
            named["B"] = {"workload": workload_name(n_gpus, "B"), "n_gpus": n_gpus, "atoms": len(atomsB), "steps": 5,
                          "ms_per_step": b_leg["ms_per_step"], "atoms_per_s": b_leg["value"], "e2e_atoms_per_s": b_leg["e2e"]["value"],
                          "roofline": b_leg["roofline"], "reduction_transport": b_leg["reduction_transport"], "parity": b_leg.get("parity"),
                          "gpu_launches": b_leg["gpu_launches"]}
        leg_s["B"] = round(time.perf_counter() - t_leg, 1)
    if do_D:
        t_leg = time.perf_counter()
        buildD = descriptor_builder(local, "D")
        atomsD, xmlD = ctx.build_shared("named_D", lambda d: buildD(d, n_gpus))
        saved = CONFIG
        d_leg = leg_static(ctx, args, "D", atomsD, xmlD, 3, 3, with_cpu=False, full_parity=False)
        if rank == 0:
            named["D"] = {"workload": workload_name(n_gpus, "D"), "n_gpus": n_gpus, "atoms": len(atomsD), "steps": 3,
                          "ms_per_step": d_leg["ms_per_step"], "atoms_per_s": d_leg["value"], "e2e_atoms_per_s": d_leg["e2e"]["value"],
                          "roofline": d_leg["roofline"], "reduction_transport": d_leg["reduction_transport"], "parity": d_leg.get("parity"),
                          "gpu_launches": d_leg["gpu_launches"]}
        assert saved == CONFIG
        leg_s["D"] = round(time.perf_counter() - t_leg, 1)

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak" if CONFIG == "A" else "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(n_gpus, N), "clocks": head["clocks"], "e2e": head["e2e"],
                "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "wall_s_timed_region": head["wall_s_timed_region"],
                "energy_eV": head["energy_eV"], "reduction_transport": head["reduction_transport"], "reduction": head["reduction"]}
        for k in ("parity", "cpu_baseline"):
            if k in head:
                line[k] = head[k]
        if named:
            line["named_configs"] = named
        # wall clock of the whole run by leg (model building, warm-up, oracle comparisons and the CPU leg included), seconds
        line["wall_s_by_leg"] = dict(leg_s, total=round(time.perf_counter() - t_start, 1))
        emit(json.dumps(line))
        bad = [k for k, p in [("headline", head.get("parity"))] + [(k, v.get("parity")) for k, v in named.items() if v] if p is not None and not p["ok"]]
        if ctx.share:
            shutil.rmtree(ctx.share, ignore_errors=True)
        if bad:
            print("bench.py: PARITY FAILURE in %s" % bad, file=sys.stderr)
            if real_world > 1:
                dist.destroy_process_group()
            raise SystemExit(3)
    if real_world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(text):
    """The ONE line of the contract goes to the process's original stdout; everything libraries print meanwhile (NCCL's
    version banner, torch warnings) has been routed to stderr by main()."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the benchmarked configuration (profiling runs)")
    ap.add_argument("--named-configs", default="auto", choices=["auto", "none", "B", "C", "D", "all"],
                    help="auto: config C MD at every N, config D at N = 1 and 8, config B at N = 1; none: headline only")
    ap.add_argument("--md-steps", type=int, default=12, help="MD steps of the config C leg")
    ap.add_argument("--emulate-world", type=int, default=0, help="profiling aid (1 GPU): run rank 0's share of a W-rank weak-scaling step, no reduction")
    ap.add_argument("--config", default="A", choices=["A", "B", "C", "D"], help="A = the bench line; B/C/D = the other BASELINE shapes (exploration)")
    args = ap.parse_args()
    global CONFIG
    CONFIG = args.config
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
