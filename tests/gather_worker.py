"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU, NCCL): shard a work list,
solve the shard with device-side selection, gather E and the selected eigenvector columns to rank 0 from DEVICE
buffers (bspatom_batch_download into torch CUDA tensors -> dist.gather), and check on rank 0 against solving the
whole list on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspatom_b200 as bsp  # noqa: E402
from bspatom_b200.parallel import gather_eigenpairs_device, shard_items  # noqa: E402
from cases import host_basis  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    a = host_basis(kind_grid=0, k=7, nfun=200, rb=100.0)
    n = a.nfun
    probs = [bsp.Problem(k=a.k, nfun=n, nkp=a.nkp, ka=a.ka, rt=a.rt, pot_kind=bsp.POT_COULOMB, pot_par=(1.0 + 0.5 * z,))
             for z in range(2 * world)]
    items_all = [(p, l) for p in probs for l in range(3)]          # groups of 3 contiguous l per charge
    # shard whole charges (a selection group stays on one rank): charge z -> rank z mod world
    mine = [i for i, (p, l) in enumerate(items_all) if (i // 3) % world == rank]
    items = [items_all[i] for i in mine]
    sel = bsp.Selection.from_kind_pi(0.4, 3)
    atom = bsp.BspAtom(device=local)
    atom.batch_upload(items, select=sel)
    atom.batch_run()
    nsel = atom.selection()
    nloc = len(items)
    E_dev = torch.empty((nloc, n), dtype=torch.float64, device="cuda")
    C_dev = torch.zeros((nloc, n, n), dtype=torch.float64, device="cuda")      # cap layout: block p = (column, row)
    atom.batch_download_ptrs(E_dev.data_ptr(), C_dev.data_ptr(), None)
    ms = torch.tensor([int(nsel.max())], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    maxsel = int(ms[0])
    C_cols = C_dev[:, :maxsel, :].contiguous().view(nloc, maxsel * n)
    ns_dev = torch.from_numpy(nsel.astype(np.float64)).cuda().view(nloc, 1)
    Eg, Cg, sent = gather_eigenpairs_device(torch.cat([E_dev, ns_dev], dim=1), C_cols, dst=0)
    ok = 1
    if rank == 0:
        assert Eg.is_cuda and Cg.is_cuda and Eg.shape == (world, nloc, n + 1)
        ref = bsp.BspAtom(device=local)
        Es, Cs, info = ref.solve_batch(items_all, select=sel)
        nref = ref.selection()
        for r in range(world):
            ids = [i for i in range(len(items_all)) if (i // 3) % world == r]
            for j, i in enumerate(ids):
                e = Eg[r, j, :n].cpu().numpy()
                k = int(Eg[r, j, n].item())
                c = Cg[r, j].view(maxsel, n)[:k].cpu().numpy().T
                if not (np.array_equal(e, Es[i]) and k == nref[i] and np.array_equal(c, np.asarray(Cs[i])[:, :k])):
                    ok = 0
        ref.close()
        print("gather_worker: world %d, %d items, maxsel %d, bit-identical to one GPU: %s" % (world, len(items_all), maxsel, bool(ok)))
    else:
        assert Eg is None and Cg is None and sent == E_dev.numel() * 8 + nloc * 8 + C_cols.numel() * 8
    t = torch.tensor([ok], device="cuda")
    dist.broadcast(t, src=0)
    atom.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t[0]) == 1 else 1)


if __name__ == "__main__":
    main()
