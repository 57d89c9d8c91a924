"""N>1 host logic on CPU: two gloo ranks evaluate their centre blocks (with the oracle standing in for the GPU
kernels), all-reduce the packed [E | virial | F] buffer exactly as ShardedPotential does, and must reproduce the
single-rank result (the reference's MPI scheme, IPModel_GAP.f95:373-391, 538-556)."""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from quip_b200 import read_xyz
from quip_b200.potential import broadcast_comm_id, pack_results, partition_bounds, reduce_packed, unpack_results
from tests.models import GOLDEN, si_two_descriptor_model


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, xml, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = read_xyz(os.path.join(GOLDEN, "Si.np1.xyz"))[8]
    first, last = partition_bounds(rank, world, len(a))
    o = orc.Model(xml).calc(a, first=first, last=last, nthreads=1)
    t = torch.from_numpy(pack_results(o["energy"], o["virial"], o["force"]))
    reduce_packed(t)
    # the host's share of the in-library reduction: rank 0 creates the 128-byte communicator id, torch.distributed hands it to every rank
    cid = broadcast_comm_id()
    assert isinstance(cid, bytes) and len(cid) == 128
    np.save(os.path.join(out_dir, "id%d.npy" % rank), np.frombuffer(cid, dtype=np.uint8))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partition_bounds_cover():
    for N in (0, 1, 7, 96, 4096):
        for w in (1, 2, 3, 8):
            b = [partition_bounds(r, w, N) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[k][1] == b[k + 1][0] for k in range(w - 1))


def test_two_rank_gloo_allreduce_matches_single_rank():
    with tempfile.TemporaryDirectory() as tmp:
        xml = si_two_descriptor_model(tmp)
        a = read_xyz(os.path.join(GOLDEN, "Si.np1.xyz"))[8]
        full = orc.Model(xml).calc(a)
        mp.spawn(_worker, args=(2, _free_port(), xml, tmp), nprocs=2, join=True)
        r0 = np.load(os.path.join(tmp, "r0.npy"))
        r1 = np.load(os.path.join(tmp, "r1.npy"))
        id0, id1 = np.load(os.path.join(tmp, "id0.npy")), np.load(os.path.join(tmp, "id1.npy"))
    assert np.array_equal(r0, r1)
    assert np.array_equal(id0, id1) and id0.any()  # both ranks hold the same, non-trivial communicator id
    u = unpack_results(r0, len(a))
    assert abs(u["energy"] - full["energy"]) < 1e-9 * abs(full["energy"])
    assert np.abs(u["force"] - full["force"]).max() < 1e-10
    assert np.abs(u["virial"] - full["virial"]).max() < 1e-9
