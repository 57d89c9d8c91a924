#!/usr/bin/env python
"""Sharded MD check (BASELINE config C shape, reduced): run under torchrun on G GPUs,
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/md_sharded_check.py [n_atoms_target] [steps]
Every rank runs ShardedPotential.run (gap_md_run_device + NCCL all-reduce hook); rank 0 also runs the single-GPU driver
gap_md_run on the same initial state and prints the deviation, the energy drift and the MD throughput."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from quip_b200 import Potential, ShardedPotential
    from quip_b200 import synthetic as syn
    from quip_b200.potential import element_masses

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    # first argument: cells per edge of a config-A-shaped Si system, or "C" = BASELINE config C (262,144-atom amorphous carbon,
    # 9,000 sparse points; the single-GPU comparison run is skipped there)
    config_c = len(sys.argv) > 1 and sys.argv[1] == "C"
    n_cells = int(sys.argv[1]) if len(sys.argv) > 1 and not config_c else 6
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix="md_sharded_r%d_" % rank)
    boot = syn.bootstrap_xml(os.path.join(tmp, "boot.xml"), [(syn.SOAP_C if config_c else syn.SOAP_A, syn.soap_dimension(8, 8))])
    bp = Potential("", param_filename=boot, device=local)
    if config_c:
        atoms, xml = syn.build_config_C(tmp, lambda desc, at: bp.descriptor_calc(at, 0)[0])
    else:
        atoms, xml = syn.build_config_A(tmp, lambda desc, at: bp.descriptor_calc(at, 0)[0], n_cells=n_cells, M=500, seed=1)
    bp.finalise()
    m = element_masses(atoms.numbers)
    rng = np.random.default_rng(7)
    v0 = rng.normal(size=atoms.positions.shape) * np.sqrt(8.617385e-5 * 300.0 / m)[:, None]
    v0 -= (m[:, None] * v0).sum(axis=0) / m.sum()
    sp = ShardedPotential("", param_filename=xml, device=local, rank=rank, world_size=world)
    a1 = type(atoms)(atoms.numbers, atoms.positions.copy(), atoms.cell, True)
    sp.run(a1, v0, dt=0.5, n_steps=1 if config_c else 2)  # warm-up (buffers, NCCL communicator)
    a1 = type(atoms)(atoms.numbers, atoms.positions.copy(), atoms.cell, True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v1, ep1, ek1 = sp.run(a1, v0, dt=0.5, n_steps=steps)
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    out = None
    if rank == 0 and config_c:
        etot = ep1 + ek1
        out = {"config": "C: amorphous carbon, SOAP cutoff 5.5 n_max=8 l_max=8, 9000 sparse points, neighbour list rebuilt every step",
               "n_gpus": world, "atoms": len(atoms), "steps": steps, "energy_drift_eV": float(np.abs(etot - etot[0]).max()),
               "ekin0_eV": float(ek1[0]), "md_atom_steps_per_s": len(atoms) * (steps + 1) / dt_wall, "ms_per_md_step": 1e3 * dt_wall / (steps + 1)}
    elif rank == 0:
        single = Potential("", param_filename=xml, device=local)
        a0 = type(atoms)(atoms.numbers, atoms.positions.copy(), atoms.cell, True)
        v0s, ep0, ek0 = single.run(a0, v0, dt=0.5, n_steps=steps)
        etot = ep1 + ek1
        out = {"n_gpus": world, "atoms": len(atoms), "steps": steps, "max_dpos_vs_single_gpu": float(np.abs(a1.positions - a0.positions).max()),
               "max_dvel_vs_single_gpu": float(np.abs(v1 - v0s).max()), "max_depot_vs_single_gpu": float(np.abs(ep1 - ep0).max()),
               "energy_drift_eV": float(np.abs(etot - etot[0]).max()), "ekin0_eV": float(ek1[0]),
               "md_atom_steps_per_s": len(atoms) * (steps + 1) / dt_wall, "ms_per_md_step": 1e3 * dt_wall / (steps + 1)}
    if world > 1:
        # replicas must stay bit-identical: compare the final positions of all ranks
        t = torch.tensor(a1.positions, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            out["replica_spread"] = float((hi - lo).abs().max())
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
