"""GPU, N > 1: the sharded work list + the single NCCL gather of eigenpairs from device buffers (SURVEY.md 8(e):
`ncclSend/ncclRecv` gather of E and C to the rank that runs the writers).  Needs >= 2 GPUs on the box
(`gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_nccl_gather_of_selected_eigenpairs_is_bit_identical_to_one_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "gather_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bit-identical to one GPU: True" in out.stdout


def test_one_process_two_devices_equals_one_device():
    """bspatom_solve_batch_multi on two different GPUs from ONE process (what a single-process Fortran host binds,
    SURVEY.md 8(b)): contiguous ranges per device, results bit-identical to one handle on one device"""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    sys.path.insert(0, HERE)
    import bspatom_b200 as bsp
    from cases import host_basis

    a = host_basis(kind_grid=0, k=7, nfun=300, rb=150.0)
    items = [(a.problem(), l) for l in range(9)]
    one = bsp.BspAtom(device=0)
    E1, C1, i1 = one.solve_batch(items, nvec=40)
    multi = bsp.BspAtomMulti([0, 1])
    E2, C2, i2 = multi.solve_batch(items, nvec=40)
    multi.close()
    one.close()
    assert not i1.any() and not i2.any()
    for l in range(9):
        assert np.array_equal(E1[l], E2[l]) and np.array_equal(C1[l], C2[l])
