"""Host-side mirror of the reference's CRM interface types.

Same names, members and allocation semantics as the Fortran derived types the one call site of ``crm()`` uses
(reference PHYS/crm_physics.F90:753-756, 1145-1156):

  crm_input_type   CRM/crm_input_module.F90:10-47    (initialize(ncol, nlev) allocates + zero-fills)
  crm_state_type   CRM/crm_state_module.F90:12-56    (pointers into the physics buffer, shape (pcols,nx,ny,nzm))
  crm_rad_type     CRM/crm_rad_module.F90:17-34
  crm_output_type  CRM/crm_output_module.F90:7-133   (crm_output_initialize(output, ncol, nlev))

Arrays are numpy float64 in Fortran order so that the memory layout is exactly what the Fortran caller owns.
"""
import numpy as np
from . import abi


class CrmDims:
    """Compile-time constants of the reference (PHYS/crmdims.F90:13-22) + hidden module inputs of crm()."""

    def __init__(self, ncrms, nx, ny, nzm, plev, dx, dt, dt_gl=1800.0, ld=None, nx_rad=None, ny_rad=None,
                 igstep=3, micro=0, use_crm_accel=False, crm_accel_uv=False, crm_accel_factor=0.0, twiddle_f32=False, esmt=False):
        self.ncrms, self.nx, self.ny, self.nzm, self.plev = int(ncrms), int(nx), int(ny), int(nzm), int(plev)
        self.ld = int(ld if ld is not None else ncrms)
        self.nx_rad = int(nx_rad if nx_rad is not None else max(1, nx // 4))
        self.ny_rad = int(ny_rad if ny_rad is not None else max(1, ny // 4))
        self.yes3d = 1 if ny > 1 else 0
        self.dx, self.dt, self.dt_gl = float(dx), float(dt), float(dt_gl)
        self.igstep, self.micro = int(igstep), int(micro)
        self.use_crm_accel, self.crm_accel_uv = bool(use_crm_accel), bool(crm_accel_uv)
        self.crm_accel_factor = float(crm_accel_factor)
        self.twiddle_f32 = bool(twiddle_f32)
        self.esmt = bool(esmt)   # -DSP_ESMT: explicit scalar momentum transport (scalar_momentum.F90)

    def c_struct(self):
        d = abi.crm_dims_t()
        for n, _ in abi.crm_dims_t._fields_:
            setattr(d, n, type(getattr(d, n))(getattr(self, n)))
        return d

    def copy(self, **kw):
        import copy
        o = copy.copy(self)
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    @property
    def cells(self):
        return self.ncrms * self.nx * self.ny * self.nzm


class _Bag:
    _fields = ()

    def arrays(self):
        return {n: getattr(self, n) for n, _ in self._fields if getattr(self, n, None) is not None}


class crm_input_type(_Bag):
    _fields = abi.INPUT_FIELDS + [("gcol", "icol")]

    def initialize(self, dims):
        for n, kind in abi.INPUT_FIELDS:
            setattr(self, n, np.zeros(abi.shape_of(kind, dims), order="F"))
        self.gcol = np.zeros(dims.ld, dtype=np.int32)
        if getattr(dims, "esmt", False):   # crm_input_module.F90:110-113 (allocated only in SP_ESMT builds)
            for n, kind in abi.INPUT_FIELDS_ESMT:
                setattr(self, n, np.zeros(abi.shape_of(kind, dims), order="F"))
        return self

    def names(self):
        return [n for n, _ in abi.INPUT_FIELDS] + ["gcol"] + [n for n, _ in abi.INPUT_FIELDS_ESMT if getattr(self, n, None) is not None]

    def c_struct(self):
        s = abi.crm_input_t()
        for n, _ in abi.INPUT_FIELDS:
            setattr(s, n, abi.dptr(getattr(self, n)))
        s.gcol = abi.iptr(self.gcol)
        for n, _ in abi.INPUT_FIELDS_ESMT:
            setattr(s, n, abi.dptr(getattr(self, n, None)))
        return s


class crm_state_type(_Bag):
    """crm_state_type; the m2005 members (nc, qr, nr, qi, ni, qs, ns, qg, ng, qc: crm_state_module.F90:34-45) exist when
    initialize() is given dims with micro = 1 (qp and qn are then unused, as in the reference's m2005 build)."""
    _fields = [(n, "crm3d") for n in abi.STATE_FIELDS_SAM1MOM]

    def names(self):
        return list(abi.STATE_FIELDS_SAM1MOM) + [n for n in abi.STATE_FIELDS_M2005 if getattr(self, n, None) is not None]

    def initialize(self, dims):
        for n in abi.STATE_FIELDS_SAM1MOM + (abi.STATE_FIELDS_M2005 if getattr(dims, "micro", 0) == 1 else []):
            setattr(self, n, np.zeros(abi.shape_of("crm3d", dims), order="F"))
        return self

    def c_struct(self):
        s = abi.crm_state_t()
        for n in self.names():
            setattr(s, n, abi.dptr(getattr(self, n)))
        return s

    def copy(self):
        o = crm_state_type()
        for n in self.names():
            setattr(o, n, getattr(self, n).copy(order="F"))
        return o


class crm_rad_type(_Bag):
    _fields = [(n, "rad") for n in abi.RAD_FIELDS_SAM1MOM]

    def names(self):
        return [n for n in abi.RAD_FIELDS if getattr(self, n, None) is not None]

    def initialize(self, dims):
        for n in (abi.RAD_FIELDS if getattr(dims, "micro", 0) == 1 else abi.RAD_FIELDS_SAM1MOM):
            setattr(self, n, np.zeros(abi.shape_of("rad", dims), order="F"))
        return self

    def c_struct(self):
        s = abi.crm_rad_t()
        for n in self.names():
            setattr(s, n, abi.dptr(getattr(self, n)))
        return s

    def copy(self):
        o = crm_rad_type()
        for n in self.names():
            setattr(o, n, getattr(self, n).copy(order="F"))
        return o


class crm_output_type(_Bag):
    _fields = abi.OUTPUT_FIELDS

    def initialize(self, dims):
        """crm_output_initialize(output, ncol, nlev): allocate + zero (CRM/crm_output_module.F90:138-330)."""
        for n, kind in (abi.OUTPUT_FIELDS + (abi.OUTPUT_FIELDS_M2005 if getattr(dims, "micro", 0) == 1 else []) +
                        (abi.OUTPUT_FIELDS_ESMT if getattr(dims, "esmt", False) else [])):
            setattr(self, n, np.zeros(abi.shape_of(kind, dims), order="F"))
        return self

    def c_struct(self):
        s = abi.crm_output_t()
        for n, _ in abi.OUTPUT_FIELDS + abi.OUTPUT_FIELDS_M2005 + abi.OUTPUT_FIELDS_ESMT:
            setattr(s, n, abi.dptr(getattr(self, n, None)))
        return s


class crm_ecpp_output_type:
    """Opaque pass-through: untouched unless -DECPP (reference crm_module.F90:123, 1600-1657)."""
    pass
