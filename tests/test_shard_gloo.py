"""CPU-only: the N>1 path.  Two processes (torch.distributed, gloo, 127.0.0.1) each take their shard of one synthetic
batch, run crm() on it through the C ABI (here: the dev-only emulator build of the device code, since there is no GPU),
and the gathered result must equal what the reference's chunking gives: one crm() call per shard (ncycle/nprec are
batch-global maxima over the columns of ONE call, kurant.F90:54-64).  No collective is used on the data path."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import common as cm

WORKER = r'''
import os, sys, importlib, pickle
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["CRM_ROOT"]); sys.path.insert(0, os.path.join(os.environ["CRM_ROOT"], "tests"))
import common as cm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
mode = os.environ["CRM_SHARD_MODE"]
dims = cm.small_dims(ncrms=7, nx=8, nsteps=3, igstep=1)
dims, inp, st, rad = cm.pkg.synthetic.make_case(dims, seed=4242)
idx = cm.pkg.shard.shard_indices(dims.ncrms, rank, world, mode)
d, si, ss, sr = cm.pkg.shard.take_shard(dims, inp, st, rad, idx)
h = cm.libmod.CrmHandle(d, lib_path=cm.EMU_LIB)
out = h.run(d, si, ss, sr)
stats = h.stats()
h.close()
res = {n: cm.pkg.shard.gather_rows(getattr(out, n), idx, dims.ncrms, world, dist) for n in ("sltend", "qltend", "precc", "cld")}
res["qt"] = cm.pkg.shard.gather_rows(ss.qt, idx, dims.ncrms, world, dist)
ncyc = [None] * world
dist.all_gather_object(ncyc, int(stats["ncycle_sum"]))
if rank == 0:
    res["ncycle_sum"] = ncyc
    with open(os.environ["CRM_OUT"], "wb") as f:
        pickle.dump(res, f)
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_indices_cover_and_disjoint():
    for n, w in [(7, 2), (48602, 8), (5, 8), (194402, 8)]:
        for mode in ("block", "roundrobin"):
            parts = [cm.pkg.shard.shard_indices(n, r, w, mode) for r in range(w)]
            allidx = np.concatenate(parts)
            assert len(allidx) == n and len(np.unique(allidx)) == n
            assert max(len(q) for q in parts) - min(len(q) for q in parts) <= (-(-n // w) if mode == "block" else 1)


@pytest.mark.parametrize("mode", ["block", "roundrobin"])
def test_two_rank_sharded_run_matches_per_shard_oracle(tmp_path, mode):
    subprocess.check_call(["make", "-C", os.path.join(cm.ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    port = _free_port()
    out = tmp_path / "res.pkl"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CRM_ROOT=cm.ROOT,
                   CRM_OUT=str(out), CRM_SHARD_MODE=mode, CRM_B200_CT="4", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    import pickle
    res = pickle.load(open(out, "rb"))
    # reference behaviour: one crm() per shard
    dims = cm.small_dims(ncrms=7, nx=8, nsteps=3, igstep=1)
    dims, inp, st, rad = cm.pkg.synthetic.make_case(dims, seed=4242)
    want = {n: np.zeros_like(res[n]) for n in res if n != "ncycle_sum"}
    for r in range(2):
        idx = cm.pkg.shard.shard_indices(dims.ncrms, r, 2, mode)
        d, si, ss, sr = cm.pkg.shard.take_shard(dims, inp, st, rad, idx)
        o, stats = cm.ob.run(d, si, ss, sr)
        assert stats["ncycle_sum"] == res["ncycle_sum"][r]
        for n in ("sltend", "qltend", "precc", "cld"):
            want[n][idx] = getattr(o, n)
        want["qt"][idx] = ss.qt
    for n in want:
        scale = max(np.abs(want[n]).max(), 1e-300)
        assert np.abs(want[n] - res[n]).max() / scale < 1e-10, n
