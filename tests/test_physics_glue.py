"""crm_physics_tend glue around crm() (crm_b200_physics_pre / crm_b200_physics_post; reference
PHYS/crm_physics.F90:1006-1011, 1160-1209): the device path against the oracle restatement and against the numpy
statement of the same lines.  The CPU test runs the device code under the dev emulator; the gpu test goes through the C ABI
with host arrays and with device pointers."""
import os
import subprocess

import numpy as np
import pytest

import common as cm


def _case(ncrms=11, ld=13, seed=3):
    dims = cm.small_dims(ncrms=ncrms, nx=16, nsteps=1, ld=ld)
    rng = np.random.default_rng(seed)
    pdel = np.asfortranarray(rng.uniform(500., 3000., (dims.ld, dims.plev)))
    qrad = np.asfortranarray(rng.normal(0., 2e-5, cm.pkg.abi.shape_of("rad", dims)))
    out = cm.pkg.crm_output_type().initialize(dims)
    for n, kind in cm.pkg.abi.OUTPUT_FIELDS:
        if kind in ("lev", "col"):
            getattr(out, n)[...] = rng.normal(0., 1.0, getattr(out, n).shape)
    return dims, pdel, qrad, out


def _copy_out(dims, out):
    o = cm.pkg.crm_output_type().initialize(dims)
    for n, _ in cm.pkg.abi.OUTPUT_FIELDS:
        getattr(o, n)[...] = getattr(out, n)
    return o


def _check(handle_factory):
    dims, pdel, qrad, out = _case()
    n = dims.ncrms
    q_o, q_g, out_o, out_g = qrad.copy(order="F"), qrad.copy(order="F"), _copy_out(dims, out), _copy_out(dims, out)
    h = handle_factory(dims)
    # ---- pre
    cm.ob.physics_pre(dims, pdel, q_o)
    h.physics_pre(dims, pdel, q_g)
    ref = qrad / pdel[:, ::-1][:, None, None, :dims.nzm]              # qrad(i,:,:,m) / pdel(i, plev-m+1)
    assert np.array_equal(q_o[:n], ref[:n]) and np.array_equal(q_g[:n], ref[:n])
    assert np.array_equal(q_g[n:], qrad[n:])                          # rows ncrms..ld-1 untouched
    # ---- post
    pt_o = cm.ob.physics_post(dims, pdel, q_o, out_o)
    pt_g = h.physics_post(dims, pdel, q_g, out_g)
    h.close()
    for name, src in (("s", "sltend"), ("q_vapor", "qltend"), ("q_cldliq", "qcltend"), ("q_cldice", "qiltend"), ("cld", "cld")):
        assert np.array_equal(pt_g[name][:n], getattr(out, src)[:n]), name
        assert np.array_equal(pt_g[name][:n], pt_o[name][:n]) and np.all(pt_g[name][n:] == 0.0)
    assert np.array_equal(out_g.precc[:n], (out.precc + out.precl)[:n]) and np.array_equal(out_g.precsc[:n], (out.precsc + out.precsl)[:n])
    assert np.all(out_g.precl[:n] == 0.0) and np.all(out_g.precsl[:n] == 0.0)
    assert np.array_equal(pt_g["prec_dp"][:n], out_g.precc[:n]) and np.array_equal(pt_g["snow_dp"][:n], out_g.precsc[:n])
    for name in ("precc", "precl", "precsc", "precsl"):
        assert np.array_equal(getattr(out_g, name)[:n], getattr(out_o, name)[:n]) and np.array_equal(getattr(out_g, name)[n:], getattr(out, name)[n:])
    # qrad is restored to Q*dp (bit-identical to the oracle; within 1 ulp of the input because (q/p)*p rounds twice)
    assert np.array_equal(q_g[:n], q_o[:n]) and np.abs(q_g - qrad).max() <= 2.3e-16 * np.abs(qrad).max()


def test_physics_glue_emulated_device_path():
    subprocess.check_call(["make", "-C", os.path.join(cm.ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    os.environ.setdefault("CRM_B200_CT", "4")
    _check(lambda dims: cm.libmod.CrmHandle(dims, lib_path=cm.EMU_LIB))


@pytest.mark.gpu
def test_physics_glue_on_device_host_arrays():
    _check(lambda dims: cm.libmod.CrmHandle(dims))


@pytest.mark.gpu
def test_physics_glue_on_device_resident_arrays():
    """qrad and pdel as device pointers (the resident-state use case of SURVEY 8(f) item 1): no staging copies."""
    import torch
    dims, pdel, qrad, out = _case(ncrms=40, ld=40, seed=9)
    h = cm.libmod.CrmHandle(dims)
    dq = torch.from_numpy(np.ascontiguousarray(qrad.ravel(order="F"))).cuda()
    dp = torch.from_numpy(np.ascontiguousarray(pdel.ravel(order="F"))).cuda()
    d = dims.c_struct()
    import ctypes as C
    rc = h.lib.crm_b200_physics_pre(h.h, C.byref(d), cm.pkg.abi.dptr(dp.data_ptr()), cm.pkg.abi.dptr(dq.data_ptr()))
    assert rc == 0
    torch.cuda.synchronize()
    got = dq.cpu().numpy().reshape(qrad.shape, order="F")
    ref = qrad / pdel[:, ::-1][:, None, None, :dims.nzm]
    assert np.array_equal(got, ref)
    h.close()
