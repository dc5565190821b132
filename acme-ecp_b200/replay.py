"""Replay files: the arguments of one crm() call on disk (BASELINE config 1 is specified as "crm() I/O replayed").

Written here, consumed by the plain-C host ``examples/crm_replay.c`` (and readable back here).  Format (little endian):
``"CRMRPLY1"``, int32 sizeof(crm_dims_t), crm_dims_t, int32 nrec, then records of char name[40] (``input.tl``, ``state.qt``,
``rad.qrad``, ``output.sltend`` ...), int32 dtype (0 = f64, 1 = i32), int32 has_data, int64 count, data.  Arrays are stored exactly as
the Fortran caller owns them (column-major, CRM index fastest, leading dimension ld).
"""
import ctypes as C
import struct

import numpy as np

from . import abi
from .crm_types import CrmDims, crm_output_type

MAGIC = b"CRMRPLY1"


def _rec(f, name, arr, has_data=True):
    a = np.asarray(arr)
    dtype = 1 if a.dtype == np.int32 else 0
    assert dtype == 1 or a.dtype == np.float64, (name, a.dtype)
    f.write(name.encode().ljust(40, b"\0"))
    f.write(struct.pack("<iiq", dtype, 1 if has_data else 0, a.size))
    if has_data:
        f.write(a.ravel(order="F").tobytes())


def write_replay(path, dims, inp, state, rad, out=None):
    """One crm() call: inputs with data, crm_state / crm_rad with data (in/out), crm_output as shapes only."""
    if out is None:
        out = crm_output_type().initialize(dims)
    recs = [("input." + n, getattr(inp, n), True) for n in inp.names()]
    recs += [("state." + n, getattr(state, n), True) for n in state.names()]
    recs += [("rad." + n, getattr(rad, n), True) for n in rad.names()]
    recs += [("output." + n, getattr(out, n), False) for n, _ in abi.OUTPUT_FIELDS + abi.OUTPUT_FIELDS_M2005 + abi.OUTPUT_FIELDS_ESMT
             if getattr(out, n, None) is not None]
    d = dims.c_struct()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<i", C.sizeof(d)))
        f.write(bytes(d))
        f.write(struct.pack("<i", len(recs)))
        for name, arr, has in recs:
            _rec(f, name, arr, has)


def read_replay(path):
    """Returns (CrmDims, {record name: flat numpy array})."""
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC, "not a replay file"
        (sz,) = struct.unpack("<i", f.read(4))
        d = abi.crm_dims_t.from_buffer_copy(f.read(sz))
        (nrec,) = struct.unpack("<i", f.read(4))
        out = {}
        for _ in range(nrec):
            name = f.read(40).split(b"\0")[0].decode()
            dtype, has, count = struct.unpack("<iiq", f.read(16))
            np_t = np.int32 if dtype == 1 else np.float64
            out[name] = np.frombuffer(f.read(count * np.dtype(np_t).itemsize), dtype=np_t).copy() if has else np.zeros(count, np_t)
    dims = CrmDims(ncrms=d.ncrms, nx=d.nx, ny=d.ny, nzm=d.nzm, plev=d.plev, dx=d.dx, dt=d.dt, dt_gl=d.dt_gl, ld=d.ld, nx_rad=d.nx_rad,
                   ny_rad=d.ny_rad, igstep=d.igstep, micro=d.micro, use_crm_accel=bool(d.use_crm_accel), crm_accel_uv=bool(d.crm_accel_uv),
                   crm_accel_factor=d.crm_accel_factor, twiddle_f32=bool(d.twiddle_f32), esmt=bool(d.esmt))
    return dims, out
