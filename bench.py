#!/usr/bin/env python
"""bench.py -- CRM cell-updates/s of the B200-native crm() (BASELINE.json metric) + roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2] [--ncrms n]

A "step" is one crm() call (one GCM time step, dt_gl = 1800 s => nstop = 180 CRM steps, each split into ncycle
sub-steps by the batch-global CFL) over the whole batch of synthetic aquaplanet columns.  Workload at N=1: BASELINE
config 2 (48,602 CRMs of 64x1x58, sam1mom, dt 10 s).  N>1: the same per-GPU batch on every rank (weak scaling; GCM
columns partition with no inter-GPU traffic, so there is no collective on the data path).
value  = cell-updates/s with the CRM state resident in HBM (device pointers through the C ABI)
e2e    = same through the C ABI with HOST buffers (pinned): H2D of crm_state/crm_rad + D2H of state and all outputs
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# SURVEY.md 8(d) / App. C algorithmic traffic model (fixed constants)
BYTES_PER_CELL_UPDATE = {"2d_sam1mom": 872.0, "3d_sam1mom": 880.0, "2d_m2005": 1456.0}
CONFIG_CLASS = {"cfg1": "2d_sam1mom", "cfg2": "2d_sam1mom", "cfg3": "2d_m2005", "cfg4": "3d_sam1mom", "cfg5": "2d_sam1mom"}
# distinct 3-D arrays read + written per cell by each kernel of this implementation (same table there)
KERNEL_WORDS = {
    "2d": {"k_p1_forcing": 14, "k_p2_sgs": 14, "k_p3_mom": 7, "k_p4_press_fwd": 9, "k_p5_press_bwd": 15,
           "k_mpdata": 4, "k_mpdata3": 8, "k_precip_a": 3, "k_precip_b": 7, "k_p7_diffuse": 7, "k_p8_micro": 13, "k_stats": 9.5,
           "k_kurant": 3, "k_m2005_column_p1": 13, "k_m2005_column_p2": 0, "k_m2005_column_p3": 0, "k_m2005_scatter": 42, "k_diffuse_set": 7},
    "3d": {"k_p1_forcing": 16, "k3_p2_sgs": 14, "k3_p3_adv": 9, "k3_p3_dif": 10, "k3_p4_press_fwd": 13, "k3_p5_press_bwd": 23,
           "k3_mpdata": 5, "k_precip_a": 3, "k_precip_b": 7, "k_p7_diffuse": 7, "k_p8_micro": 14, "k_stats": 9.5, "k_kurant": 4},
}


CHUNKED_KERNELS = ("k_m2005_column_p1", "k_m2005_column_p2", "k_m2005_column_p3", "k_m2005_scatter")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        sm, mx, reasons = [], [], set()
        with open(self.f.name) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}
        return out


def ncu_traffic(kernel, ncrms):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of the same
    batch size (profiles/*ncu_traffic*.json, written by tools/ncu_traffic.py from tools/gpu_traffic.sh); None if no
    capture of this batch size exists."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_traffic*.json"))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("ncrms") == ncrms and kernel in d.get("kernels", {}):
            best = d["kernels"][kernel]["dram_bytes_per_launch"]
    return best


def build_case(pkg, cfg, ncrms, seed):
    kw = dict(pkg.synthetic.CONFIGS[cfg])
    if ncrms:
        kw["ncrms"] = ncrms
    if kw["ny"] > 1:
        kw.update(nx_rad=kw["nx"] // 4, ny_rad=kw["ny"] // 4)
    dims = pkg.CrmDims(dt_gl=1800.0, igstep=1, **kw)
    return pkg.synthetic.make_case(dims, seed=seed)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_ncrms(args):
    """The CPU arm parallelises over CRMs with a barrier per sub-step: at least 8 CRMs per host thread."""
    return args.cpu_ncrms if args.cpu_ncrms > 0 else max(128, 8 * host_threads())


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The Fortran cannot be compiled here (no
    Fortran compiler), so this is the restated reference (oracle/, kind "port") with all host threads, on a bounded
    sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = importlib.import_module("acme-ecp_b200")
    import oracle_binding as ob
    ncrms, nstop = cpu_sample_ncrms(args), args.cpu_nstop
    dims, inp, st, rad = build_case(pkg, args.config, ncrms, 20260101 + int(args.config[3:]))
    dims = dims.copy(dt_gl=dims.dt * nstop)
    # all host threads this process may use (torchrun sets OMP_NUM_THREADS=1 for its ranks; only rank 0 works here)
    navail = host_threads()
    ob.lib().crm_oracle_set_num_threads(int(navail))
    cores = ob.lib().crm_oracle_num_threads()
    for w in range(args.warmup):
        ob.run(dims.copy(igstep=1 if w == 0 else 3), inp, st, rad)
    t0 = time.time()
    upd = 0
    for _ in range(args.steps):
        _, s = ob.run(dims.copy(igstep=3), inp, st, rad)
        upd += dims.cells * s["ncycle_sum"]
    dt = time.time() - t0
    val = upd / dt
    sample = "%d CRMs x %d CRM steps per step of the %s geometry (same seed), %d host threads (%d CRMs per thread)" % (
        ncrms, nstop, args.config, cores, ncrms // max(1, cores))
    line = {"impl": "reference", "metric": "CRM cell-updates/s", "value": val, "unit": "cell-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, dims) + "; CPU arm: %d host threads" % cores, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args, dims):
    return "%s: %d CRMs/GPU of %dx%dx%d %s, crm_dt %g s, dt_gl %g s (nstop %d), synthetic aquaplanet columns" % (
        args.config, dims.ncrms, dims.nx, dims.ny, dims.nzm, "m2005" if dims.micro else "sam1mom", dims.dt, dims.dt_gl, int(dims.dt_gl / dims.dt))


def run_e2e(args, pkg, abi, h, dims, inp, ci, st, rad, dstate, big, world, barrier, torch, dist, state_shape, st_names):
    """End-to-end arm: HOST (pinned) buffers through the same C-ABI call (H2D of crm_state/crm_rad, D2H of the state
    and of every output inside the timed region)."""
    def pinned_like(a):
        tns = torch.empty(a.size, dtype=torch.float64).pin_memory()
        v = tns.numpy().reshape(a.shape, order="F")
        v[...] = a
        return tns, v
    keep = []
    # host RAM is shared by all local ranks: the pinned boundary buffers of the full batch are ~31 GB per rank at cfg2.  If
    # they do not fit this rank's share, the end-to-end arm runs the leading CRMs that do (same call, same arrays with a
    # smaller ncrms; the throughput of a batch of many thousand CRMs does not depend on its size) and says so.
    lworld = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cell3 = dims.nx * dims.ny * dims.nzm * 8.0
    per_crm = cell3 * (len(st_names) + len(big) + 6.0 * dims.nx_rad * dims.ny_rad / (dims.nx * dims.ny)) + 45.0 * dims.plev * 8.0
    try:
        avail = float([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024.0
    except Exception:
        avail = 64e9 * lworld
    n_e2e = dims.ncrms
    budget = 0.6 * avail / max(1, lworld)
    if os.environ.get("CRM_BENCH_HOST_GB"):   # test hook: pretend this rank's share of the host RAM is this many GB
        budget = float(os.environ["CRM_BENCH_HOST_GB"]) * 1e9
    if per_crm * dims.ncrms > budget:
        n_e2e = max(64, int(budget / per_crm) // 64 * 64)
    n_e2e = min(n_e2e, dims.ncrms)
    d_e = dims.copy(ncrms=n_e2e, ld=n_e2e) if n_e2e < dims.ncrms else dims
    hst = pkg.crm_state_type(); hrad = pkg.crm_rad_type(); hout = pkg.crm_output_type().initialize(d_e)
    hst.initialize(dims.copy(ncrms=1, ld=1)); hrad.initialize(dims.copy(ncrms=1, ld=1))   # creates the member set of this microphysics
    for n in st_names:
        full = dstate[n].view(-1)
        src = full.view(dims.nx * dims.ny * dims.nzm, dims.ld)[:, :n_e2e].contiguous().cpu().numpy().reshape(-1)   # (ld fastest) -> leading CRMs
        tns, v = pinned_like(src.reshape((n_e2e,) + tuple(state_shape[1:]), order="F")); keep.append(tns); setattr(hst, n, v)
        del src
    for n in rad.names():
        tns, v = pinned_like(np.asfortranarray(getattr(rad, n)[:n_e2e])); keep.append(tns); setattr(hrad, n, v)
    for n in big:
        tns, v = pinned_like(getattr(hout, n)); keep.append(tns); setattr(hout, n, v)
    hcs, hcr, hco = hst.c_struct(), hrad.c_struct(), hout.c_struct()
    if n_e2e < dims.ncrms:
        hin = pkg.crm_input_type().initialize(d_e)
        for n, _ in abi.INPUT_FIELDS:
            getattr(hin, n)[...] = getattr(inp, n)[:n_e2e]
        hin.gcol[...] = inp.gcol[:n_e2e]
        keep.append(hin)
        ci = hin.c_struct()
    dims_full, dims = dims, d_e
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h.run_structs(dims.copy(igstep=3), ci, hcs, hcr, hco)   # warm-up (page-locks / first touch)
    barrier()
    t0 = time.time()
    upd_e, h2d, d2h = 0, 0, 0
    for _ in range(e2e_steps):
        h.run_structs(dims.copy(igstep=3), ci, hcs, hcr, hco)
        s = h.stats()
        upd_e += dims.cells * s["ncycle_sum"]; h2d = s["h2d_bytes"]; d2h = s["d2h_bytes"]
    barrier()
    te = time.time() - t0
    t = torch.tensor([te, float(upd_e)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        te, upd_e = float(tmax[0]), float(tsum[1])
    res = {"value": upd_e / te, "unit": "cell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "host_buffers": "pinned", "ncrms_per_gpu": int(n_e2e)}
    if n_e2e < dims_full.ncrms:
        res["note"] = "host RAM share of this rank holds the pinned boundary arrays of %d of %d CRMs" % (n_e2e, dims_full.ncrms)
    if args.e2e_resident:
        # Same call with crm_state / crm_rad resident in HBM keyed by the physics chunk (crm_b200_resident_*):
        # per step only crm_input + qrad go in and crm_output + the crm_rad results come out (host buffers, pinned).  The one-off
        # upload of the chunk's state (crm_b200_resident_put, the first GCM step) is outside the timed region.
        import ctypes as C
        dstate.clear()                 # the device-pointer arm is finished: give its state (10-22 GB) back before the resident copy is made
        torch.cuda.empty_cache()
        d3 = dims.copy(igstep=3)
        h.resident_put(0, d3, hst, hrad)
        cd = d3.c_struct()

        def rrun():
            rc = h.lib.crm_b200_resident_run(h.h, 0, C.byref(cd), C.byref(ci), C.byref(hcr), C.byref(hco))
            if rc != 0:
                raise RuntimeError("crm_b200_resident_run failed (%d): %s" % (rc, h.lib.crm_b200_last_error(h.h).decode()))
        rrun()
        barrier()
        t0 = time.time()
        upd_r, h2d, d2h = 0, 0, 0
        for _ in range(e2e_steps):
            rrun()
            s = h.stats()
            upd_r += dims.cells * s["ncycle_sum"]; h2d = s["h2d_bytes"]; d2h = s["d2h_bytes"]
        barrier()
        tr = time.time() - t0
        t = torch.tensor([tr, float(upd_r)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            tr, upd_r = float(tmax[0]), float(tsum[1])
        res["resident"] = {"value": upd_r / tr, "unit": "cell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                           "steps": e2e_steps, "resident_bytes": h.resident_bytes(),
                           "what": "crm_state/crm_rad kept in HBM by crm_b200_resident_run; crm_input, qrad in, crm_output and crm_rad results out"}
        h.resident_drop(-1)
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("acme-ecp_b200")
    libmod = importlib.import_module("acme-ecp_b200.lib")
    abi = pkg.abi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lrank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    partition = args.partition
    if partition == "auto":   # cfg5 is THE column-partition case (one global batch over the GPUs); the others weak-scale
        partition = "roundrobin" if (args.config == "cfg5" and world > 1) else "replicate"
    if partition == "replicate":
        dims, inp, st, rad = build_case(pkg, args.config, args.ncrms, 20260101 + int(args.config[3:]) + 1000 * rank)
    else:   # one global batch, ordered by latitude, partitioned with shard.shard_indices
        dims, inp, st, rad, _ = pkg.synthetic.make_sharded_case(args.config, rank, world, partition, ncrms=args.ncrms or None, dt_gl=1800.0, igstep=1)
    if args.nstop:
        dims = dims.copy(dt_gl=dims.dt * args.nstop)
    h = libmod.CrmHandle(dims, device=lrank)
    out = pkg.crm_output_type().initialize(dims)
    big = ("qcl", "qci", "qpl", "qpi", "tk", "tkh") + (("wvar", "aut", "acc", "evpc", "evpr", "mlt", "sub", "dep", "con") if dims.micro == 1 else ())
    st_names = st.names()

    # ---- device-resident arm: crm_state, crm_rad and the 3-D outputs live in HBM
    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda()
    dstate = {}
    state_shape = getattr(st, st_names[0]).shape
    for n in st_names:     # upload field by field and drop the host copy (8 ranks share one host's RAM)
        dstate[n] = to_dev(getattr(st, n))
        setattr(st, n, None)
    drad = {n: to_dev(getattr(rad, n)) for n in rad.names()}
    dout = {n: torch.zeros(dstate["qt"].numel(), dtype=torch.float64, device="cuda") for n in big}
    cs = abi.crm_state_t(); cr = abi.crm_rad_t(); co = out.c_struct(); ci = inp.c_struct()
    for n, tns in dstate.items():
        setattr(cs, n, abi.dptr(tns.data_ptr()))
    for n, tns in drad.items():
        setattr(cr, n, abi.dptr(tns.data_ptr()))
    for n, tns in dout.items():
        setattr(co, n, abi.dptr(tns.data_ptr()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def call(igstep):
        h.run_structs(dims.copy(igstep=igstep), ci, cs, cr, co)
        return h.stats()

    for w in range(args.warmup):
        call(1 if w == 0 else 3)
    barrier()
    sampler = ClockSampler(lrank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    upd, launches, ncyc, kms = 0, 0, 0, 0.0
    ev0.record()
    for _ in range(args.steps):
        s = call(3)
        upd += dims.cells * s["ncycle_sum"]; launches += s["kernel_launches"]; ncyc += s["ncycle_sum"]; kms += s["ms_kernels"]
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, float(upd), float(launches), float(ncyc), float(dims.ncrms)], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        allt = torch.stack(allt).cpu().numpy()
        ms_r = allt[:, 0]
        ms, upd_all, launches_all = float(ms_r.max()), float(allt[:, 1].sum()), int(allt[:, 2].sum())
        # load balance of the column partition: device time, sub-step count and idle share (waiting for the slowest rank) per rank
        per_rank = {"ms": [round(float(x), 1) for x in ms_r], "ncycle_sum": [int(x) for x in allt[:, 3]], "ncrms": [int(x) for x in allt[:, 4]],
                    "idle_frac": [round(float(1.0 - x / ms_r.max()), 4) for x in ms_r],
                    "balance_efficiency": float(ms_r.mean() / ms_r.max())}
    else:
        upd_all, launches_all = float(upd), int(launches)
    value = upd_all / (ms * 1e-3)

    # ---- per-kernel device times of one more call (events around every launch) -> roofline of the dominant kernel
    peak, peak_src = peaks()
    roof = None
    if rank == 0:
        h.set_kernel_timing(True)
        s = call(3)
        kt = h.kernel_times()
        h.set_kernel_timing(False)
        tot = sum(v[0] for v in kt.values())
        kwords = KERNEL_WORDS["3d" if dims.ny > 1 else "2d"]
        bpcu = BYTES_PER_CELL_UPDATE[CONFIG_CLASS.get(args.config, "2d_sam1mom")]
        name = max((k for k in kt if k in kwords), key=lambda k: kt[k][0])
        ms_k, n_k = kt[name]
        # cells one launch processes: the whole batch, except for the m2005 column kernels, which sweep the batch once per
        # sub-step in chunks of tiles (the column work space is sized for a chunk)
        def cells_per_launch(k):
            return dims.cells * s["ncycle_sum"] / kt[k][1] if k in CHUNKED_KERNELS else float(dims.cells)
        alg = kwords[name] * 8.0 * cells_per_launch(name)            # bytes per launch
        ach = alg / (ms_k / n_k * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic(name, dims.ncrms),
                "peak_source": peak_src, "kernel_share_of_step": ms_k / tot,
                "algorithmic_bytes_per_launch": alg,
                "whole_step": {"bytes_per_cell_update": bpcu, "achieved": bpcu * (value / world) / 1e9, "frac": bpcu * (value / world) / 1e9 / peak},
                "kernel_frac": {k: round(kwords[k] * 8.0 * cells_per_launch(k) / (v[0] / v[1] * 1e-3) / 1e9 / peak, 3) for k, v in kt.items() if k in kwords},
                "kernel_ms": {k: round(v[0], 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])}}

    e2e = None
    if args.e2e_steps > 0:
        e2e = run_e2e(args, pkg, abi, h, dims, inp, ci, st, rad, dstate, big, world, barrier, torch, dist, state_shape, st_names)

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_binding as ob
        try:
            ob.lib().crm_oracle_set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        d2, i2, s2, r2 = build_case(pkg, args.config, cpu_sample_ncrms(args), 20260101 + int(args.config[3:]))
        d2 = d2.copy(dt_gl=d2.dt * args.cpu_nstop)
        ob.run(d2.copy(igstep=1), i2, s2, r2)
        t0 = time.time()
        _, so = ob.run(d2.copy(igstep=3), i2, s2, r2)
        tc = time.time() - t0
        cpu = {"value": d2.cells * so["ncycle_sum"] / tc, "unit": "cell-updates/s", "cores": ob.lib().crm_oracle_num_threads(),
               "kind": "port", "sample": "%d CRMs x %d CRM steps of the %s geometry after a %d-step spin-up, OpenMP over CRMs (%d per thread)" %
               (d2.ncrms, args.cpu_nstop, args.config, args.cpu_nstop, d2.ncrms // max(1, ob.lib().crm_oracle_num_threads()))}
    if rank == 0:
        line = {"metric": "CRM cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak" if partition == "replicate" else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, dims), "l2": "inputs larger than L2 (44 GB of state per GPU)" if dims.cells > 2e7 else "small batch",
                           "mean_ncycle": ncyc / max(1, args.steps * int(dims.dt_gl / dims.dt)),
                           "partition": ("every GPU runs its own batch of this size (weak scaling), no collective" if partition == "replicate" else
                                         "one global batch of %d columns ordered by latitude, %s-partitioned over %d GPUs (shard.shard_indices), no collective" %
                                         (int(sum(per_rank["ncrms"])) if per_rank else dims.ncrms, partition, world)),
                           "per_rank": per_rank},
                "gpu_launches": launches_all, "kernel_ms_per_step": kms / max(args.steps, 1),
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--ncrms", type=int, default=0, help="override the batch size (testing only)")
    ap.add_argument("--nstop", type=int, default=0, help="override CRM steps per call (testing only)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--e2e-resident", type=int, default=1, help="also time the end-to-end call with the state resident in HBM (0 = skip)")
    ap.add_argument("--cpu-ncrms", type=int, default=0, help="CRMs of the CPU sample (default: max(128, 8 per host thread))")
    ap.add_argument("--partition", default="auto", choices=["auto", "replicate", "block", "roundrobin"],
                    help="replicate: every rank runs the config's batch (weak scaling); block/roundrobin: the config's batch is ONE global "
                         "batch partitioned over the ranks (auto: roundrobin for cfg5 on several GPUs, else replicate)")
    ap.add_argument("--cpu-nstop", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
