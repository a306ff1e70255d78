"""CPU: the N>1 path (sharding of the (instance, l) list + the single eigenpair gather) with
world_size 2 over gloo.  The per-rank solve is replaced by the oracle here ONLY because this
container has no GPU; on the box the same plumbing runs over NCCL in bench.py / test_gpu_*."""
import os
import socket
import sys

import numpy as np
import pytest

from bspatom_b200.parallel import gather_eigenpairs, shard_items

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_items_is_a_partition():
    for n in (1, 7, 51, 4096):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in shard_items(n, r, world))
            assert got == list(range(n))
            sizes = [len(shard_items(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_items(4, 2, 2)


def _worker(rank, world, port, nitems, nfun, nvec, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard_items(nitems, rank, world)
    E = np.array([[100.0 * i + j for j in range(nfun)] for i in ids])
    Cm = np.array([[1000.0 * i + j for j in range(nfun * nvec)] for i in ids])
    Eg, Cg = gather_eigenpairs(E, ids, nitems, nfun, C_local=Cm, nvec=nvec, dst=0)
    if rank == 0:
        q.put((Eg, Cg))
    else:
        assert Eg is None and Cg is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    nitems, nfun, nvec, world = 5, 6, 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nitems, nfun, nvec, q)) for r in range(world)]
    for p in procs:
        p.start()
    Eg, Cg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(Eg, np.array([[100.0 * i + j for j in range(nfun)] for i in range(nitems)]))
    assert np.array_equal(Cg, np.array([[1000.0 * i + j for j in range(nfun * nvec)] for i in range(nitems)]))


def _worker_dev(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from bspatom_b200.parallel import gather_eigenpairs_device

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E = torch.arange(12, dtype=torch.float64).view(3, 4) + 100 * rank
    Cc = torch.arange(24, dtype=torch.float64).view(3, 8) + 1000 * rank
    Eg, Cg, sent = gather_eigenpairs_device(E, Cc, dst=0)
    if rank == 0:
        q.put((Eg.numpy(), Cg.numpy()))
    else:
        assert Eg is None and Cg is None and sent == 8 * (12 + 24)
    dist.barrier()
    dist.destroy_process_group()


def test_tensor_gather_world2_gloo():
    """gather_eigenpairs_device (the NCCL path of bench.py / tests/gather_worker.py) with CPU tensors over gloo"""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_dev, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    Eg, Cg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert Eg.shape == (2, 3, 4) and Cg.shape == (2, 3, 8)
    for r in range(2):
        assert np.array_equal(Eg[r], np.arange(12.0).reshape(3, 4) + 100 * r)
        assert np.array_equal(Cg[r], np.arange(24.0).reshape(3, 8) + 1000 * r)


def test_gather_single_process():
    E = np.arange(12.0).reshape(3, 4)
    Eg, Cg = gather_eigenpairs(E, [0, 1, 2], 3, 4)
    assert np.array_equal(Eg, E) and Cg is None


def test_numa_binding_helper_is_best_effort():
    """bind_host_memory_to_gpu never raises: unknown devices / single-node hosts leave everything as it is."""
    import os

    from bspatom_b200.parallel import bind_host_memory_to_gpu

    before = os.sched_getaffinity(0)
    out = bind_host_memory_to_gpu("0000:ff:1f.0")          # no such device here
    assert out["node"] is None and out["mempolicy"] is False
    assert os.sched_getaffinity(0) == before


def _worker_dev_ragged(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from bspatom_b200.parallel import gather_eigenpairs_device

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nloc = 3 - rank                      # 3 and 2 rows: the 21 l values of cfg4 over 8 ranks give 3,3,3,3,3,2,2,2
    E = torch.arange(4 * nloc, dtype=torch.float64).view(nloc, 4) + 100 * rank
    Cc = torch.arange(8 * nloc, dtype=torch.float64).view(nloc, 8) + 1000 * rank
    Eg, Cg, sent = gather_eigenpairs_device(E, Cc, dst=0)
    if rank == 0:
        q.put(([e.numpy() for e in Eg], [c.numpy() for c in Cg]))
    else:
        assert Eg is None and Cg is None and sent == 8 * 3 * (4 + 8)     # padded to the largest share on the wire
    dist.barrier()
    dist.destroy_process_group()


def test_tensor_gather_ragged_shares_world2_gloo():
    """unequal shares (an 8-GPU cfg4 run hung in NCCL on exactly this before the row counts were exchanged first)"""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_dev_ragged, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    Eg, Cg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(2):
        nloc = 3 - r
        assert np.array_equal(Eg[r], np.arange(4.0 * nloc).reshape(nloc, 4) + 100 * r)
        assert np.array_equal(Cg[r], np.arange(8.0 * nloc).reshape(nloc, 8) + 1000 * r)
