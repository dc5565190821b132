"""Column partition of a crm() batch across GPUs.

CRMs are independent, so GCM columns shard with NO collective on the data path: every rank owns a handle on its GPU and
runs crm() on its own columns, exactly like the real model runs one crm() per chunk per MPI rank
(reference PHYS/crm_physics.F90:1145).  The only cross-CRM coupling of the reference is the batch-global max behind
ncycle (CRM/kurant.F90:54-64) and nprec (MICRO_SAM1MOM/microphysics.F90:603,646-647); evaluating it per shard is what
chunking does in the reference too.  torch.distributed (gloo on CPU, nccl on GPUs) is used only by the harness to gather
results / timings, never between sub-steps.
"""
import numpy as np

from . import abi
from .crm_types import crm_input_type, crm_state_type, crm_rad_type


def shard_indices(ncrms, rank, world, mode="block"):
    """Column indices owned by ``rank``.  "block": contiguous gcol ranges of ceil(ncrms/world) columns (the reference's
    chunking); "roundrobin": gcol interleaved across ranks, which balances the per-shard ncycle when the batch is
    ordered by latitude (tropical columns sub-cycle more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if mode == "block":
        per = -(-ncrms // world)
        return np.arange(min(ncrms, rank * per), min(ncrms, (rank + 1) * per))
    if mode == "roundrobin":
        return np.arange(rank, ncrms, world)
    raise ValueError("mode must be 'block' or 'roundrobin'")


def take_shard(dims, inp, state, rad, idx):
    """Copies of the boundary arrays restricted to the columns ``idx`` (leading dimension = len(idx))."""
    n = len(idx)
    d = dims.copy(ncrms=n, ld=n)
    si, ss, sr = crm_input_type().initialize(d), crm_state_type().initialize(d), crm_rad_type().initialize(d)
    for name, _ in abi.INPUT_FIELDS:
        getattr(si, name)[...] = getattr(inp, name)[idx]
    si.gcol[...] = inp.gcol[idx]
    for name in abi.STATE_FIELDS_SAM1MOM:
        getattr(ss, name)[...] = getattr(state, name)[idx]
    for name in abi.RAD_FIELDS_SAM1MOM:
        getattr(sr, name)[...] = getattr(rad, name)[idx]
    return d, si, ss, sr


def gather_rows(local, idx, ncrms, world, dist=None):
    """All-gather per-column results (first axis = column) back into global column order.  ``dist`` is
    torch.distributed (initialised) or None for world == 1."""
    local = np.ascontiguousarray(local)
    if world == 1 or dist is None:
        out = np.zeros((ncrms,) + local.shape[1:], dtype=local.dtype)
        out[idx] = local
        return out
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(idx), local))
    out = np.zeros((ncrms,) + local.shape[1:], dtype=local.dtype)
    for ii, a in parts:
        out[ii] = a
    return out
