#!/usr/bin/env python
"""Small-pcols regime: H independent handles (one per physics chunk, each with its own stream) driven by H host threads at the same time,
against ONE handle running the same number of CRMs in one call.  ctypes releases the GIL during crm_b200_run, so the calls overlap on
the GPU.  python tools/multi_handle.py [--config cfg2] [--ncrms 256] [--handles 8] [--nstop 30]"""
import argparse, importlib, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2"); ap.add_argument("--ncrms", type=int, default=256)
ap.add_argument("--handles", type=int, default=8); ap.add_argument("--nstop", type=int, default=30)
a = ap.parse_args()
pkg = importlib.import_module("acme-ecp_b200"); libmod = importlib.import_module("acme-ecp_b200.lib")
kw = dict(pkg.synthetic.CONFIGS[a.config])


def case(n, seed):
    k = dict(kw); k["ncrms"] = n
    d = pkg.CrmDims(dt_gl=k["dt"] * a.nstop, igstep=1, **k)
    return pkg.synthetic.make_case(d, seed=seed)


def timed(handles, cases, reps=2):
    """every handle runs its case `reps` times on its own thread; returns (wall seconds of the timed rep, sub-steps taken)"""
    upd = [0] * len(handles)

    def work(i, igstep):
        d, inp, st, rad = cases[i]
        handles[i].run(d.copy(igstep=igstep), inp, st, rad)
        upd[i] = d.cells * handles[i].stats()["ncycle_sum"]
    wall = 0.
    for rep in range(reps + 1):                      # rep 0: spin-up with setperturb, untimed
        th = [threading.Thread(target=work, args=(i, 1 if rep == 0 else 3)) for i in range(len(handles))]
        t0 = time.time()
        for t in th: t.start()
        for t in th: t.join()
        wall = time.time() - t0
    return wall, sum(upd)


res = {}
cases = [case(a.ncrms, 20260200 + i) for i in range(a.handles)]
hs = [libmod.CrmHandle(c[0]) for c in cases]
w1, u1 = timed(hs[:1], cases[:1])
res["one_handle"] = {"ncrms": a.ncrms, "ms_per_call": 1e3 * w1, "cell_updates_per_s": u1 / w1}
wh, uh = timed(hs, cases)
res["concurrent_handles"] = {"handles": a.handles, "ncrms_each": a.ncrms, "ms_per_round": 1e3 * wh, "cell_updates_per_s": uh / wh}
for h in hs: h.close()
big = case(a.ncrms * a.handles, 20260300)
hb = libmod.CrmHandle(big[0])
wb, ub = timed([hb], [big])
res["one_big_call"] = {"ncrms": a.ncrms * a.handles, "ms_per_call": 1e3 * wb, "cell_updates_per_s": ub / wb}
hb.close()
res["config"] = {"geometry": a.config, "nstop": a.nstop, "host_buffers": "pageable numpy arrays (boundary copies included)"}
print(json.dumps(res))
