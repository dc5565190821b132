"""The plain-C host of the C ABI (examples/crm_replay.c): compiled with gcc against include/crm_b200.h, it replays a recorded crm()
call from a file -- the reference's hot path entered from a compiled host language, without Python or torch in the process.
CPU tests run it against the dev-only emulation build of the device sources (tests/emu); the GPU test against libcrm_b200.so.
Results must be bit-identical to the ctypes path (same library, same inputs)."""
import os
import re
import subprocess

import numpy as np
import pytest

import common as cm

SRC = os.path.join(cm.ROOT, "examples", "crm_replay.c")
EXE = os.path.join(cm.ROOT, "examples", "crm_replay")
HDR = os.path.join(cm.ROOT, "include", "crm_b200.h")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(cm.ROOT, "include"), SRC, "-o", EXE, "-ldl"])
    return EXE


def test_member_tables_match_the_header():
    """Every pointer member of crm_input_t / crm_state_t / crm_rad_t / crm_output_t appears in the C host's name tables, in header order."""
    import test_shim_static as ts
    structs = ts._c_structs()
    src = open(SRC).read()
    for tab, sname in (("k_input", "crm_input_t"), ("k_state", "crm_state_t"), ("k_rad", "crm_rad_t"), ("k_output", "crm_output_t")):
        body = src[src.index("static const member_t %s[]" % tab):]
        body = body[:body.index("};")]
        names = re.findall(r"M\(%s,\s*(\w+)\)" % sname, body)
        assert names == [m for m, _ in structs[sname]], tab


def _replay(exe, lib_path, dims, tmp_path, ncalls=2):
    dims, inp, st, rad = cm.make_case(dims, 90)
    fin, fout = str(tmp_path / "in.crmr"), str(tmp_path / "out.crmr")
    cm.pkg.replay.write_replay(fin, dims, inp, st, rad)
    log = subprocess.run([exe, lib_path, fin, fout, str(ncalls)], check=True, capture_output=True, text=True).stdout
    assert log.count("kernel launches") == ncalls
    d2, got = cm.pkg.replay.read_replay(fout)
    assert (d2.ncrms, d2.nx, d2.nzm, d2.ld) == (dims.ncrms, dims.nx, dims.nzm, dims.ld)
    h = cm.libmod.CrmHandle(dims, lib_path=lib_path)
    for call in range(ncalls):
        out = h.run(dims if call == 0 or dims.igstep > 1 else dims.copy(igstep=2), inp, st, rad)
    h.close()
    for n in st.names():
        assert np.array_equal(got["state." + n], getattr(st, n).ravel(order="F")), n
    for n in rad.names():
        assert np.array_equal(got["rad." + n], getattr(rad, n).ravel(order="F")), n
    nout = 0
    for n, _ in cm.pkg.abi.OUTPUT_FIELDS:
        assert np.array_equal(got["output." + n], getattr(out, n).ravel(order="F")), n
        nout += 1
    assert nout > 60 and float(np.abs(got["output.sltend"]).max()) > 0.
    return log


def test_c_host_replays_a_call_on_the_emulated_device(exe, tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(cm.ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    os.environ.setdefault("CRM_B200_CT", "4")
    _replay(exe, cm.EMU_LIB, cm.small_dims(ncrms=5, nx=8, nsteps=2, ld=6), tmp_path)


def test_c_host_reports_library_errors(exe, tmp_path):
    """A geometry the library refuses: the C host prints crm_b200_last_error and exits non-zero (the shim calls endrun there)."""
    subprocess.check_call(["make", "-C", os.path.join(cm.ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    dims = cm.small_dims(ncrms=2, nx=8, nsteps=1)
    dims, inp, st, rad = cm.pkg.synthetic.make_case(dims)
    bad = dims.copy()
    bad.nx = 12                                     # not a power of two
    fin = str(tmp_path / "bad.crmr")
    cm.pkg.replay.write_replay(fin, bad, inp, st, rad)
    r = subprocess.run([exe, cm.EMU_LIB, fin, str(tmp_path / "o.crmr")], capture_output=True, text=True)
    assert r.returncode != 0 and "crm_b200_init failed" in r.stderr and "power of two" in r.stderr


@pytest.mark.gpu
def test_c_host_replays_calls_on_the_gpu(exe, tmp_path):
    log = _replay(exe, cm.libmod.DEFAULT_LIB, cm.small_dims(ncrms=37, nx=32, nsteps=4, ld=40), tmp_path)
    assert "sub-steps" in log
