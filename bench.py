#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json metric).

metric   : solves/sec, batched FP64 B-spline generalized eigenproblems N=1000 (+ max eig rel err)
workload : BASELINE.json configs[1]: Coulomb, l = 0..50, N = 1000 B-splines of order k = 7,
           Rmax = 500 a.u., ALL eigenpairs.  One "solve" = one (instance, l) pencil: its share of the
           assembly + all eigenvalues + all eigenvectors.  One "step" = one batch of ZREP nuclear
           charges x 51 values of l per GPU (weak scaling: every rank owns its own charges
           Z_i = 1 + i/64).
value    : whole-job solves/s with the batch resident in HBM (bspatom_batch_run only), device time
           from CUDA events on the library's stream, max over ranks.
e2e      : same metric through the reference-facing call bspatom_solve_batch with HOST buffers:
           H2D of knots/parameters and D2H of E and C (pinned memory) inside the timed region.
--impl reference : the reference's own CPU algorithm (oracle restatement of MATRIX_SVT + LAPACK dsygv
           with the arguments of matrices.f90:248) on the box's host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NFUN, K, RB, LMAX = 1000, 7, 500.0, 50
ZREP_DEFAULT = 8
EPS = np.finfo(float).eps


# ------------------------------------------------------------------------------------------
def workload_items(bsp, rank: int, zrep: int, grid: str):
    """(Problem, l) list of this rank: zrep charges x (LMAX+1) angular momenta."""
    if grid == "lin":
        inp = bsp.BspInputs.from_values(kind_grid=0, k=K, nfun=NFUN, rb=RB)
    else:  # exp-lin knots that land on exactly N=1000 with monotone knots
        inp = bsp.BspInputs.from_values(kind_grid=2, k=K, nfun=782, rb=RB, rmax=70.0)
    assert inp.nfun == NFUN
    items = []
    for iz in range(zrep):
        z = 1.0 + (rank * zrep + iz) / 64.0
        p = bsp.Problem(k=inp.k, nfun=inp.nfun, nkp=inp.nkp, ka=inp.ka, rt=inp.rt, pot_kind=bsp.POT_COULOMB,
                        pot_par=(z,))
        items += [(p, l) for l in range(LMAX + 1)]
    return inp, items


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons while the timed region runs (NVML, 100 ms)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
def cpu_reference_sample(grid: str, ls, threads: int):
    """The reference's algorithm on the host: MATRIX_SVT restatement (oracle, single thread like the
    shipped Makefile without -qopenmp, src/Makefile:21-23) + DSYGV(1,'V','U') from host LAPACK
    (OpenBLAS of the scipy wheel, `threads` threads) for the given l values.
    Returns (solves_per_s, t_assembly, t_solve_per_l, {l: eigenvalues})."""
    from oracle import oracle as O

    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    if grid == "lin":
        b = O.make_basis(kind_grid=0, k=K, nfun=NFUN, rb=RB)
    else:
        b = O.make_basis(kind_grid=2, k=K, nfun=782, rb=RB, rmax=70.0)
    t0 = time.perf_counter()
    m = O.matrix_svt(b, lmax=LMAX)          # the reference assembles U(:,:,0:lmax) once per run
    t_asm = time.perf_counter() - t0
    eig = {}
    t_solve = 0.0
    ctx = threadpool_limits(limits=threads) if threadpool_limits else None
    try:
        for l in ls:
            t0 = time.perf_counter()
            w, v = O.solve_system(m, l)
            t_solve += time.perf_counter() - t0
            eig[l] = w
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None) if hasattr(ctx, "__exit__") else None
    per_l = t_solve / len(ls)
    # one full reference run = 1 assembly + (LMAX+1) solves; a sample of len(ls) solves carries its share
    t_per_solve = t_asm / (LMAX + 1) + per_l
    return 1.0 / t_per_solve, t_asm, per_l, eig


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    ls = [0, 25, 50][: max(1, args.ref_ls)]
    for _ in range(args.warmup if args.warmup < 2 else 1):   # one warm-up pass is enough for a CPU baseline
        cpu_reference_sample(args.grid, ls[:1], threads)
    vals, t_all = [], 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, t_asm, per_l, _ = cpu_reference_sample(args.grid, ls, threads)
        t_all += time.perf_counter() - t0
        vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "solves/sec, batched FP64 B-spline gen. eigenproblems N=1000",
        "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 Coulomb l=0..50 N=1000 k=7 Rmax=500 all eigenpairs (%s knots)" % args.grid,
                   "sample": "1 assembly (lmax=50) + dsygv for l=%s per step" % ls},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": "oracle MATRIX_SVT restatement (1 thread) + scipy/OpenBLAS dsygv(1,'V','U') "
                                   "on %d threads, l=%s of 0..50; no Fortran compiler in the image, the "
                                   "reference itself cannot be built" % (threads, ls)},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", default="lin", choices=["lin", "explin"])
    ap.add_argument("--zrep", type=int, default=ZREP_DEFAULT, help="nuclear charges per GPU per step")
    ap.add_argument("--ref-ls", type=int, default=3, help="l values per reference step")
    ap.add_argument("--workers", type=int, default=0, help="chunk streams (0 = library default)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (experiments), repeatable")
    ap.add_argument("--no-numa", action="store_true", help="do not bind host memory / CPU affinity to the GPU's NUMA node")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import bspatom_b200 as bsp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    numa = None
    if world > 1 and not args.no_numa:
        # before any pinned allocation: host buffers on the GPU's own NUMA node (N = 1 keeps every core for the
        # CPU-baseline leg)
        from bspatom_b200.parallel import bind_host_memory_to_gpu
        prop = torch.cuda.get_device_properties(local)
        numa = bind_host_memory_to_gpu("%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id))
    atom = bsp.BspAtom(device=local)
    if args.workers:
        atom.set_option("workers", args.workers)
    for kv in args.opt:
        atom.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    inp, items = workload_items(bsp, rank, args.zrep, args.grid)
    nsolve = len(items)
    n_e = nsolve * NFUN
    n_c = nsolve * NFUN * NFUN

    # ---------------- device-resident throughput ("value") ----------------
    atom.batch_upload(items)
    for _ in range(max(args.warmup, 3)):
        atom.batch_run()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms, launches = 0.0, 0
    kms = np.zeros(4)
    kcnt = np.zeros(4)
    stage_ms = np.zeros(4)
    gaps = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        atom.batch_run()
        st = atom.stats()
        dev_ms += st["ms_total"]
        launches += int(st["launches"])
        kms += [st["ms_k_round"], st["ms_k_factor"], st["ms_k_back"], st["ms_k_assembly"]]
        kcnt += [st["n_k_round"], st["n_k_factor"], st["n_k_back"], st["n_k_assembly"]]
        stage_ms += [st["ms_assembly"], st["ms_eigenvalues"], st["ms_eigenvectors"], st["ms_finalize"]]
        gaps.append(st["ms_total"])
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    last_stats = st
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    total_solves = nsolve * world * args.steps
    value = total_solves / (dev_ms_max * 1e-3)

    # results of the last resident step (for the accuracy figure), eigenvalues only
    E_host = torch.empty(n_e, dtype=torch.float64).pin_memory()
    info = np.zeros(nsolve, dtype=np.int32)
    atom.batch_download(E_host.numpy(), None, info)
    E_last = E_host.numpy().reshape(nsolve, NFUN).copy()

    # ---------------- end to end through the C-ABI with host buffers ----------------
    # Every step = one bspatom_solve_batch call: H2D of knots/parameters, all kernels, D2H of E and C into
    # pinned host memory.  Headline mode: two handles alternate over the K steps (BspAtomPipeline, the
    # double-buffered sweep a user of the library runs), so the eigenvector D2H of one step (3.3 GB, PCIe
    # bound) overlaps the kernels of the next.  The strictly serial single-handle figure is reported too.
    e2e = None
    if not args.no_e2e:
        h2d = sum(8 * (p.nkp + 8) + 64 for p, _ in items)     # knots + parameters per problem struct
        d2h = 8 * (n_e + n_c) + 4 * nsolve
        def pinned_pair():
            return (torch.empty(n_e, dtype=torch.float64).pin_memory().numpy(),
                    torch.empty(n_c, dtype=torch.float64).pin_memory().numpy())

        bufs = [pinned_pair()]

        def gather_E(Eh):
            if world > 1:   # the single gather of the path: eigenvalues to rank 0 over NCCL / NVLink
                Eg = torch.from_numpy(Eh).cuda(non_blocking=True)
                out = [torch.empty_like(Eg) for _ in range(world)] if rank == 0 else None
                dist.gather(Eg, out, dst=0)

        # (a) serial, one handle
        for _ in range(2):
            atom.solve_batch(items, out_E=bufs[0][0], out_C=bufs[0][1])
            gather_E(bufs[0][0])
        barrier()
        t0 = time.perf_counter()
        serial_steps = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            _, _, inf = atom.solve_batch(items, out_E=bufs[0][0], out_C=bufs[0][1])
            gather_E(bufs[0][0])
            serial_steps.append(round(1e3 * (time.perf_counter() - ts), 2))
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        serial_s = float(tt[0])
        serial_stats = {k_: atom.stats()[k_] for k_ in ("wall_ms_upload", "wall_ms_run", "wall_ms_download", "wall_ms_copy_tail", "ms_total")}
        bad = int(np.count_nonzero(inf))
        # (b) two alternating handles
        pipe_s, pipe, pipe_err = None, None, None
        nb = args.steps
        try:
            bufs.append(pinned_pair())      # second result buffer: batch i+1 lands while batch i is in use
            pipe = bsp.BspAtomPipeline(device=local, depth=2)
            if args.workers:
                pipe.set_option("workers", args.workers)
            oE = [bufs[i % 2][0] for i in range(max(nb, 2))]
            oC = [bufs[i % 2][1] for i in range(max(nb, 2))]
            pipe.solve_batches([items] * 2, oE[:2], oC[:2])
        except Exception as exc:      # keep the serial figure if a second handle / buffer does not fit
            pipe_err = repr(exc)
        ok = torch.tensor([0 if pipe_err else 1], dtype=torch.int32, device="cuda")
        if world > 1:                 # every rank takes the same branch (the timed part has barriers)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]):
            barrier()
            t0 = time.perf_counter()
            infos = pipe.solve_batches([items] * nb, oE[:nb], oC[:nb])
            for i in range(nb):
                gather_E(oE[i])
            barrier()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            pipe_s = float(tt[0])
            bad += sum(int(np.count_nonzero(i_)) for i_ in infos)
        if pipe is not None:
            pipe.close()
        best_s = min(serial_s, pipe_s) if pipe_s else serial_s
        e2e = {"value": total_solves / best_s, "unit": "solves/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * best_s / args.steps,
               "mode": "two handles alternating over the steps (BspAtomPipeline)" if (pipe_s and pipe_s <= serial_s)
                       else "one handle, steps strictly serial",
               "pipelined": None if not pipe_s else {"value": total_solves / pipe_s, "ms_per_step": 1e3 * pipe_s / args.steps},
               "serial": {"value": total_solves / serial_s, "ms_per_step": 1e3 * serial_s / args.steps,
                          "ms_each_step": serial_steps, "wall_ms_last_step": serial_stats},
               "bad_info": bad}
        if pipe_err:
            e2e["pipelined_unavailable"] = pipe_err
        del bufs

    # ---------------- CPU baseline + accuracy, rank 0 only ----------------
    cpu_baseline, acc = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ls = [0, 1, 10, 25, 50]
        v, t_asm, per_l, eig = cpu_reference_sample(args.grid, ls, threads)
        cpu_baseline = {"value": v, "unit": "solves/s", "cores": threads, "kind": "port",
                        "sample": "oracle assembly lmax=50 (%.2f s, 1 thread) + OpenBLAS dsygv for l=%s (%.2f s each, "
                                  "%d threads); Z=1" % (t_asm, ls, per_l, threads)}
        rel, worst_tol = 0.0, 0.0
        for l in ls:                       # items[0..50] are Z = 1 + rank*zrep/64 = 1 for rank 0
            w = eig[l]
            e = E_last[l]
            rel = max(rel, float(np.max(np.abs(e - w) / np.maximum(np.abs(w), 1e-2))))
            tol = np.maximum(np.maximum(1e-12 * np.abs(w), 1e-10), 32 * EPS * np.abs(w).max())
            worst_tol = max(worst_tol, float(np.max(np.abs(e - w) / tol)))
        acc = {"max_eig_rel_err_vs_dsygv": rel, "max_err_over_tolerance": worst_tol,
               "tolerance": "max(1e-12|E|, 1e-10, 32 eps |E_max|)", "l_checked": ls, "info_nonzero": int(np.count_nonzero(info))}

    # ---------------- roofline of the dominant kernel ----------------
    # The value above runs two chunk streams concurrently, so a kernel's event-bracketed duration there
    # contains time shared with the other stream.  For the roofline the same batch is run `steps` more
    # times on ONE stream (every kernel alone on the GPU) and each launch is bracketed by CUDA events on
    # that stream (bspatom_get_stats out[8..15]).
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured copy bandwidth (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    atom.set_option("workers", 1)
    atom.batch_upload(items)
    atom.batch_run()
    k1ms, k1cnt, one_ms, it1 = np.zeros(4), np.zeros(4), 0.0, 0
    for _ in range(args.steps):
        atom.batch_run()
        s1 = atom.stats()
        k1ms += [s1["ms_k_round"], s1["ms_k_factor"], s1["ms_k_back"], s1["ms_k_assembly"]]
        k1cnt += [s1["n_k_round"], s1["n_k_factor"], s1["n_k_back"], s1["n_k_assembly"]]
        one_ms += s1["ms_total"]
        it1 = int(s1["iters"])
    if args.workers:
        atom.set_option("workers", args.workers)
    B = K - 1
    npad = ((NFUN + B) // (B + 1)) * (B + 1)
    per_pair = {   # algorithmic bytes per (pencil, eigenpair) and launch, DESIGN.md section 6
        "bsp_factor_kernel": 8 * (npad * (B + 1) + NFUN),                 # write (zd,l_1..l_B) rows, read rhs
        "bsp_back_kernel": 8 * (npad * (B + 1) + 3 * NFUN),               # read factors + x_old; write x, rhs
    }
    names = ["bsp_round_kernel", "bsp_factor_kernel", "bsp_back_kernel", "bsp_assemble_kernel"]
    share = k1ms / max(one_ms, 1e-9)
    dom = 1 + int(np.argmax(k1ms[1:3]))          # the HBM-bound sweep with the larger share
    dom_name = names[dom]
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "roofline_r1.json")))
    except Exception:
        pass
    dom_launches = k1cnt[dom] / args.steps
    avg_ms = k1ms[dom] / max(k1cnt[dom], 1)
    # the first min_iters (3) launches sweep every eigenpair; later ones only touch the few that missed conv_tol
    # (their bytes are not counted, their time is: the figure is a lower bound)
    min_iters = 3
    for kv in args.opt:
        if kv.split("=")[0] == "min_iters":
            min_iters = int(float(kv.split("=")[1]))
    full_launches = min(dom_launches, min_iters)
    bytes_per_launch = nsolve * NFUN * per_pair[dom_name]
    ach = bytes_per_launch * full_launches / (k1ms[dom] / args.steps * 1e-3) / 1e9
    traffic = None
    nk = ncu.get(dom_name)
    if nk:
        traffic = nk["dram_bytes"] * (nsolve / nk["pencils_per_launch"])
    roofline = {"kernel": dom_name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                "launches_per_step": dom_launches, "share_of_step": float(share[dom]),
                "how": "single chunk stream, %d pencils per launch, CUDA events around every launch; traffic = "
                       "dram__bytes_read+write of one `ncu --set full` launch of %s pencils scaled to %d"
                       % (nsolve, nk["pencils_per_launch"] if nk else "n/a", nsolve),
                "other_kernels": {
                    "bsp_factor_kernel" if dom == 2 else "bsp_back_kernel": {
                        "achieved_GBs": nsolve * NFUN * per_pair[names[3 - dom]] * full_launches
                        / (k1ms[3 - dom] / args.steps * 1e-3) / 1e9, "share_of_step": float(share[3 - dom])},
                    "bsp_round_kernel": {"share_of_step": float(share[0]), "bound": "fp64 pipe (serial pivot recurrence), "
                                         "sm__pipe_fp64_cycles_active %s %% in profiles/"
                                         % ncu.get("bsp_round_kernel", {}).get("fp64_pipe_active_pct", "n/a")}},
                "single_stream_ms_per_step": one_ms / args.steps,
                "note": "peak = measured device-to-device COPY rate (half reads, half writes); the back sweep is 80 % reads, "
                        "so its fraction of that figure can come out slightly above 1 (nominal HBM3e: 7.7 TB/s)"}
    b_alg = 8 * ((NFUN + K) + 4 * K * NFUN + NFUN + NFUN * NFUN)       # 8.24 MB per solve
    step_roof = {"bytes_per_solve": b_alg, "achieved_gbs": value / world * b_alg / 1e9,
                 "frac_of_hbm": value / world * b_alg / 1e9 / hbm_peak,
                 "note": "SURVEY 8(d) whole-solve figure; the sweeps move ~2x57 MB of factor per pencil and iteration"}

    if rank == 0:
        line = {
            "metric": "solves/sec, batched FP64 B-spline gen. eigenproblems N=1000", "value": value,
            "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2 Coulomb l=0..50 N=1000 k=7 Rmax=500 all eigenpairs (%s knots)" % args.grid,
                       "solves_per_step_per_gpu": nsolve, "charges_per_gpu": args.zrep,
                       "l2": "no flush: each chunk's factor workspace (~74 MB per pencil, ~11 GB per chunk) is far "
                             "larger than the 126 MB L2", "parallelism": "shard (Z, l) list over %d rank(s), no "
                             "collective on the compute path" % world},
            "clocks": sampler.summary(),
            "host_numa": numa,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": roofline,
            "step_roofline": step_roof,
            "cpu_baseline": cpu_baseline,
            "accuracy": acc,
            "kernel_ms_per_step_single_stream": {n_: float(m_) / args.steps for n_, m_ in zip(names, k1ms)},
            "stage_ms_per_step": dict(zip(("assembly", "eigenvalues", "eigenvectors", "finalize"),
                                          (stage_ms / args.steps).tolist())),
            "rounds": int(last_stats["rounds"]), "iters": int(last_stats["iters"]),
            "ms_each_step": gaps, "selected_third_solve_per_step": int(last_stats["selected_third_solve"]),
            "wall_ms_per_step": wall_ms_max / args.steps,
        }
        print(json.dumps(line))
    atom.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
