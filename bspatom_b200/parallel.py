"""Multi-GPU plumbing: one process per GPU, static partition of the (instance, l) work list, no
collective on the compute path, ONE gather of eigenpairs at the end.

In the reference the l-loop bodies are independent (they share only the read-only Sij, Tij, Vij,
matrices.f90:242-248), so the work list shards with no exchange step; S/H0/Q are cheap enough to be
re-assembled on every rank that needs them."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def shard_items(nitems: int, rank: int, world: int) -> List[int]:
    """item i -> rank (i mod world): every item costs the same for equal (N, k)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return list(range(rank, nitems, world))


def gather_eigenpairs(E_local: np.ndarray, idx_local: Sequence[int], nitems: int, nfun: int,
                      C_local: Optional[np.ndarray] = None, nvec: int = 0, dst: int = 0, device=None):
    """Collect the per-rank results on rank ``dst`` for the host writers (Enl.dat,
    Eigenvec_All.dat, WRITE_WF: matrices.f90:261-265,366-378).

    E_local: (len(idx_local), nfun); C_local: (len(idx_local), nfun*nvec) or None.
    Uses the default torch.distributed group: NCCL (tensors staged on ``device``) when the
    backend is nccl, gloo on CPU otherwise.  Returns (E, C) on dst, (None, None) elsewhere.
    Equal-count all_gather with padding: at most one padded item per rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        E = np.asarray(E_local).reshape(len(idx_local), nfun)
        Cm = None if C_local is None else np.asarray(C_local).reshape(len(idx_local), nfun * nvec)
        Eo = np.empty((nitems, nfun)); Eo[list(idx_local)] = E
        Co = None
        if Cm is not None:
            Co = np.empty((nitems, nfun * nvec)); Co[list(idx_local)] = Cm
        return Eo, Co
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    dev = device if backend == "nccl" else torch.device("cpu")
    per = (nitems + world - 1) // world
    width = nfun + (nfun * nvec if C_local is not None else 0)
    buf = torch.zeros((per, width), dtype=torch.float64, device=dev)
    nloc = len(idx_local)
    if nloc:
        buf[:nloc, :nfun] = torch.from_numpy(np.ascontiguousarray(E_local).reshape(nloc, nfun)).to(dev)
        if C_local is not None:
            buf[:nloc, nfun:] = torch.from_numpy(np.ascontiguousarray(C_local).reshape(nloc, nfun * nvec)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != dst:
        return None, None
    E = np.empty((nitems, nfun))
    Cm = np.empty((nitems, nfun * nvec)) if C_local is not None else None
    for r in range(world):
        ids = shard_items(nitems, r, world)
        blk = out[r][: len(ids)].cpu().numpy()
        E[ids] = blk[:, :nfun]
        if Cm is not None:
            Cm[ids] = blk[:, nfun:]
    return E, Cm


def gather_eigenpairs_device(E_dev, C_dev=None, dst: int = 0):
    """The single collective of the path on DEVICE buffers: every rank contributes E_dev (nloc, nfun) and,
    optionally, C_dev (nloc, ncols * nfun: the selected eigenvector columns of its pencils, e.g. filled by
    BspAtom.batch_download_ptrs straight from the solver's resident blocks) -- torch tensors on the rank's GPU
    (CPU tensors under gloo); rank ``dst`` receives them over NCCL / NVLink (ncclSend/Recv under dist.gather) and
    keeps them on its GPU for the writers.  nloc may differ between ranks (21 l values over 8 ranks, cfg4): the
    row counts are exchanged first (one small all_gather) and the blocks are padded to the largest one on the wire,
    so every rank issues the same collectives whatever its share.  Returns (E_all, C_all, sent_bytes): on dst lists
    with one (nloc_r, ...) tensor per rank -- stacked into one (world, nloc, ...) tensor when all shares are equal --
    and (None, None, sent_bytes) elsewhere."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return E_dev.unsqueeze(0), (None if C_dev is None else C_dev.unsqueeze(0)), 0
    world, rank = dist.get_world_size(), dist.get_rank()
    nloc = torch.tensor([int(E_dev.shape[0])], dtype=torch.int64, device=E_dev.device)
    counts = [torch.zeros_like(nloc) for _ in range(world)]
    dist.all_gather(counts, nloc)
    counts = [int(c[0]) for c in counts]
    nmax = max(counts)
    sent = 0
    outs = []
    for t in (E_dev, C_dev):
        if t is None:
            outs.append(None)
            continue
        t = t.contiguous()
        if t.shape[0] < nmax:      # ragged share: pad the block on the wire
            pad = torch.zeros((nmax - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], dim=0)
        recv = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, recv, dst=dst)
        if rank != dst:
            sent += t.numel() * t.element_size()
        if rank != dst:
            outs.append(None)
        elif min(counts) == nmax:
            outs.append(torch.stack(recv))
        else:
            outs.append([recv[r][: counts[r]] for r in range(world)])
    return outs[0], outs[1], sent


def bind_host_memory_to_gpu(pci_bus_id: str) -> dict:
    """One process per GPU: place this process (CPU affinity, as far as the cpuset allows) and its future host
    allocations (memory policy MPOL_PREFERRED) on the NUMA node the GPU hangs off.  The end-to-end path returns
    8 MB of eigenvectors per solve into pinned host memory; with 8 GPUs on a two-socket box the copies that have
    to cross the socket interconnect are what limits it.  Best effort: returns what was done, never raises."""
    import ctypes
    import os

    out = {"pci": pci_bus_id, "node": None, "cpus": None, "mempolicy": False}
    try:
        dev = pci_bus_id.lower()
        if len(dev.split(":")) == 2:
            dev = "0000:" + dev
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % dev).read().strip())
        if node < 0:
            return out
        out["node"] = node
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            out["cpus"] = len(allowed)
        # set_mempolicy(MPOL_PREFERRED = 1, nodemask, maxnode): x86_64 syscall 238
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, ctypes.byref(mask), 16 * 64)
        out["mempolicy"] = (rc == 0)
    except Exception as exc:      # unknown topology, no permission, non-Linux: leave everything as it is
        out["error"] = repr(exc)
    return out
