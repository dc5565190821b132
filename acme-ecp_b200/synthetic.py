"""Synthetic aquaplanet-like inputs for crm(); numpy only, seeded, no files besides data/vcoord.json.

The generator produces what PHYS/crm_physics.F90:1051-1130 would hand to crm(): GCM column profiles on the real E3SM
hybrid levels (L72 / L30 tables shipped with the reference, see tools/make_vcoord.py), an APE-"control"-like
sounding per column latitude, and a CRM state seeded exactly like the model's first step
(crm_physics.F90:912-939): u=ul, v=vl, w=0, T=tl, qt=ql+qccl+qiil, qp=0, qn=qccl+qiil.
"""
import json
import os
import numpy as np
from .crm_types import CrmDims, crm_input_type, crm_state_type, crm_rad_type

_VC = None
G = 9.80616
RD = 6.02214e26 * 1.38065e-23 / 28.966
CONFIGS = {
    # name: ncrms, nx, ny, nzm, plev, dx, dt   (BASELINE.json configs)
    "cfg1": dict(ncrms=866, nx=32, ny=1, nzm=28, plev=30, dx=4000.0, dt=20.0),
    "cfg2": dict(ncrms=48602, nx=64, ny=1, nzm=58, plev=72, dx=1000.0, dt=10.0),
    "cfg3": dict(ncrms=48602, nx=64, ny=1, nzm=58, plev=72, dx=1000.0, dt=10.0, micro=1),
    "cfg4": dict(ncrms=4096, nx=32, ny=32, nzm=58, plev=72, dx=1000.0, dt=10.0),
    "cfg5": dict(ncrms=194402, nx=32, ny=1, nzm=58, plev=72, dx=1000.0, dt=10.0),
}


def vcoord(plev):
    global _VC
    if _VC is None:
        with open(os.path.join(os.path.dirname(__file__), "data", "vcoord.json")) as f:
            _VC = json.load(f)
    t = _VC["L%d" % plev]
    return (np.array(t["hyai"]), np.array(t["hybi"]), np.array(t["hyam"]), np.array(t["hybm"]))


def _qsatw(T, p_pa):
    es = 610.78 * np.exp(17.269 * (T - 273.16) / (T - 35.86))   # Tetens, Pa (input generation only)
    es = np.minimum(es, 0.5 * p_pa)
    return 0.622 * es / (p_pa - es)


def make_inputs(dims, seed=20260101, lat_range=(-1.0, 1.0), rows=None, nglobal=None, sort_by_lat=False):
    """Return (crm_input, crm_state, crm_rad) for ``dims``; rows ncrms..ld-1 are filled with copies of valid columns.

    Column sharding (bench.py --config cfg5 --gpus N): with ``rows`` (global column indices owned by this rank) and
    ``nglobal`` (size of the global batch) the per-column random draws are made for the whole global batch - every rank
    draws the same numbers - optionally ordered by latitude like a GCM's column numbering, and only the columns ``rows``
    are built (dims.ld == len(rows)); the 3-D state is never materialised for columns of other ranks."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n, ld, plev, nzm = dims.ncrms, dims.ld, dims.plev, dims.nzm
    hyai, hybi, hyam, hybm = vcoord(plev)
    inp = crm_input_type().initialize(dims)
    ng = ld if rows is None else int(nglobal)
    ps = 101325.0 + 300.0 * rng.standard_normal(ng)
    sinlat = rng.uniform(lat_range[0], lat_range[1], ng)
    lon = rng.uniform(0.0, 360.0, ng)
    cloud_u = rng.uniform(0.0, 1.0, ng)
    noise_u = rng.standard_normal((ng, plev))
    noise_v = rng.standard_normal((ng, plev))
    gcol = np.arange(1, ng + 1, dtype=np.int32)
    if sort_by_lat:
        order = np.argsort(sinlat, kind="stable")
        ps, sinlat, lon, cloud_u, noise_u, noise_v = ps[order], sinlat[order], lon[order], cloud_u[order], noise_u[order], noise_v[order]
    if rows is not None:
        rows = np.asarray(rows)
        assert len(rows) == ld, "dims.ld must equal the number of owned columns"
        ps, sinlat, lon, cloud_u, noise_u, noise_v, gcol = ps[rows], sinlat[rows], lon[rows], cloud_u[rows], noise_u[rows], noise_v[rows], gcol[rows]
    lat = np.arcsin(sinlat)
    pint = hyai[None, :] * 1.0e5 + hybi[None, :] * ps[:, None]
    pmid = hyam[None, :] * 1.0e5 + hybm[None, :] * ps[:, None]
    pdel = np.diff(pint, axis=1)
    Ts = np.where(np.abs(lat) < np.pi / 3, 27.0 * (1.0 - np.sin(1.5 * lat) ** 2), 0.0) + 273.15
    sig = pmid / ps[:, None]
    T = np.maximum(200.0, (Ts[:, None] - 1.0) * sig ** (RD * 6.5e-3 / G))
    # moist boundary layer + drier free troposphere; a GCM-scale cloud deck (qccl, qiil) in the (sub)tropics makes
    # the large-scale forcing supply condensate so that every microphysics branch is exercised from the first call
    # (deviation from SURVEY.md 8(d), which has qccl=qiil=0 and leaves the CRM cloud-free for many calls).
    RH = np.clip(0.35 + 0.63 * ((sig - 0.3) / 0.6), 0.3, 0.98)
    ql = RH * _qsatw(T, pmid)
    cloudy = (cloud_u < 0.7)[:, None]
    qccl = np.where(cloudy & (sig > 0.55) & (sig < 0.9) & (T > 263.0), 1.6e-3 * np.sin(np.pi * (sig - 0.55) / 0.35) ** 2, 0.0)
    qiil = np.where(cloudy & (sig > 0.25) & (sig < 0.6) & (T < 258.0), 3.0e-4 * np.sin(np.pi * (sig - 0.25) / 0.35) ** 2, 0.0)
    Tv = T * (1.0 + 0.608 * ql)
    # hypsometric integration from the surface (phis = 0) upward
    zint = np.zeros((ld, plev + 1))
    for l in range(plev - 1, -1, -1):
        zint[:, l] = zint[:, l + 1] + RD * Tv[:, l] / G * np.log(pint[:, l + 1] / pint[:, l])
    zmid = np.zeros((ld, plev))
    for l in range(plev):
        zmid[:, l] = zint[:, l + 1] + RD * Tv[:, l] / G * np.log(pint[:, l + 1] / pmid[:, l])
    ul = 30.0 * np.sin(2 * lat)[:, None] ** 2 * (1.0 - sig) + 2.0 * noise_u
    vl = 2.0 * noise_v
    inp.zmid[...] = zmid; inp.zint[...] = zint; inp.tl[...] = T; inp.ql[...] = ql
    inp.qccl[...] = qccl; inp.qiil[...] = qiil
    inp.ps[...] = ps; inp.pmid[...] = pmid; inp.pint[...] = pint; inp.pdel[...] = pdel
    inp.ul[...] = ul; inp.vl[...] = vl
    inp.ocnfrac[...] = 1.0; inp.tau00[...] = 0.05; inp.bflxls[...] = 5.0e-3 * np.cos(lat)
    inp.wndls[...] = np.abs(ul[:, plev - 1])
    inp.lat_deg[...] = np.degrees(lat); inp.lon_deg[...] = lon
    inp.gcol[...] = gcol
    if getattr(dims, "esmt", False):   # crm_physics.F90:1126-1127: the transported scalar momentum starts from the GCM wind
        inp.ul_esmt[...] = ul; inp.vl_esmt[...] = vl

    st = crm_state_type().initialize(dims)
    rad = crm_rad_type().initialize(dims)
    for k in range(nzm):
        m = plev - 1 - k
        st.u_wind[:, :, :, k] = ul[:, m][:, None, None]
        st.v_wind[:, :, :, k] = vl[:, m][:, None, None]
        st.temperature[:, :, :, k] = T[:, m][:, None, None]
        if getattr(dims, "micro", 0) == 1:   # m2005 seeding (crm_physics.F90:926-938): qt = qv + cloud liquid, qi = cloud ice, qc = cloud liquid
            st.qt[:, :, :, k] = (ql[:, m] + qccl[:, m])[:, None, None]
            st.qi[:, :, :, k] = qiil[:, m][:, None, None]
            st.qc[:, :, :, k] = qccl[:, m][:, None, None]
        else:
            st.qt[:, :, :, k] = (ql[:, m] + qccl[:, m] + qiil[:, m])[:, None, None]
            st.qn[:, :, :, k] = (qccl[:, m] + qiil[:, m])[:, None, None]
        rad.qrad[:, :, :, k] = np.where(pmid[:, m] > 2.0e4, -1.5 / 86400.0, 0.0)[:, None, None]
        rad.temperature[:, :, :, k] = T[:, m][:, None, None]
        rad.qv[:, :, :, k] = ql[:, m][:, None, None]
    return inp, st, rad


def make_case(name_or_dims, seed=None, **over):
    """Build (dims, crm_input, crm_state, crm_rad) for a named BASELINE config ("cfg1".."cfg5") or explicit CrmDims."""
    if isinstance(name_or_dims, str):
        kw = dict(CONFIGS[name_or_dims]); kw.update(over)
        dims = CrmDims(**kw)
        if seed is None:
            seed = 20260101 + int(name_or_dims[3:])
    else:
        dims = name_or_dims
        if seed is None:
            seed = 20260101
    inp, st, rad = make_inputs(dims, seed)
    return dims, inp, st, rad


def make_sharded_case(name, rank, world, mode="roundrobin", seed=None, ncrms=None, **over):
    """This rank's share of a named config whose GLOBAL batch (ordered by latitude) is partitioned over ``world`` GPUs with
    shard.shard_indices(mode) - "block" (contiguous gcol ranges, the reference's chunking) or "roundrobin"."""
    from .shard import shard_indices
    kw = dict(CONFIGS[name]); kw.update(over)
    nglobal = int(ncrms or kw["ncrms"])
    idx = shard_indices(nglobal, rank, world, mode)
    kw["ncrms"] = len(idx)
    dims = CrmDims(**kw)
    if seed is None:
        seed = 20260101 + int(name[3:])
    inp, st, rad = make_inputs(dims, seed, rows=idx, nglobal=nglobal, sort_by_lat=True)
    return dims, inp, st, rad, idx
