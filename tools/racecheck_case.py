#!/usr/bin/env python
"""Small developed-state crm() call for compute-sanitizer (racecheck / memcheck):
   compute-sanitizer --tool racecheck --racecheck-report analysis python tools/racecheck_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm
dims = cm.small_dims(ncrms=8, nx=64, nzm=58, plev=72, nsteps=2, use_crm_accel=False)
dims, inp, st, rad = cm.make_case(dims, 20)
h = cm.libmod.CrmHandle(dims); h.run(dims, inp, st, rad); print("ran", h.stats()["kernel_launches"], "launches"); h.close()
