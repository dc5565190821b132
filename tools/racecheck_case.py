#!/usr/bin/env python
"""Small developed-state crm() calls for compute-sanitizer (racecheck / memcheck):
   compute-sanitizer --tool racecheck --racecheck-report analysis python tools/racecheck_case.py [2d|3d|m2005|all]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm
which = sys.argv[1] if len(sys.argv) > 1 else "2d"
cases = []
if which in ("2d", "all"):
    cases.append(("2d", cm.small_dims(ncrms=8, nx=64, nzm=58, plev=72, nsteps=2, use_crm_accel=False), 20))
if which in ("3d", "all"):
    cases.append(("3d", cm.pkg.CrmDims(ncrms=2, nx=32, ny=32, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=40.0, igstep=3, nx_rad=8, ny_rad=8), 6))
    cases.append(("3d-16x8", cm.pkg.CrmDims(ncrms=5, nx=16, ny=8, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=40.0, igstep=3, nx_rad=4, ny_rad=2), 6))
if which in ("2drt", "all"):   # runtime-geometry kernel variants (nx = 128: CRM tile of 2)
    cases.append(("2d-runtime-geometry", cm.small_dims(ncrms=3, nx=128, nsteps=2), 6))
if which in ("esmt", "all"):
    cases.append(("esmt", cm.small_dims(ncrms=8, nx=32, nsteps=2, esmt=True), 10))
    cases.append(("esmt-3d", cm.pkg.CrmDims(ncrms=3, nx=16, ny=16, nzm=28, plev=30, dx=4000.0, dt=20.0, dt_gl=40.0, igstep=3, nx_rad=4, ny_rad=4, esmt=True), 6))
if which in ("m2005", "all"):
    cases.append(("m2005", cm.small_dims(ncrms=8, nx=32, nsteps=2, micro=1), 10))
if which in ("resident", "all"):   # crm_b200_resident_* + crm_b200_history_field (k_hist_qv), 2-D and m2005
    for tag, d in (("resident", cm.small_dims(ncrms=7, nx=32, nsteps=2, ld=9)), ("resident-m2005", cm.small_dims(ncrms=5, nx=16, nsteps=1, micro=1))):
        d, inp, st, rad = cm.make_case(d, 10)
        h = cm.libmod.CrmHandle(d)
        h.resident_put(3, d, st, rad)
        h.resident_run(3, d, inp, rad)
        h.resident_run(3, d, inp, None)
        for n in ("CRM_U", "CRM_QV", "CRM_QC", "CRM_TKH"):
            h.history_field(3, d, n)
        h.resident_get(3, d, st, rad)
        h.resident_drop(-1)
        print(tag, "ran", h.stats()["kernel_launches"], "launches"); h.close()
for name, dims, spin in cases:
    dims, inp, st, rad = cm.make_case(dims, spin)
    h = cm.libmod.CrmHandle(dims); h.run(dims, inp, st, rad); print(name, "ran", h.stats()["kernel_launches"], "launches"); h.close()
