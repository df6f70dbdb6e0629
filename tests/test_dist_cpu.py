"""Host-side multi-GPU logic on CPU: gloo backend, world_size 2 (no GPU, no kernels).
Covers channel/PRN sharding, the packed-key MAX all-reduce that replaces the cross-GPU peak
search, and the result-table gather."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["B200_ROOT"])
from gnss_sdr_b200 import dist as bd
from gnss_sdr_b200.capi import ACQ_RESULT_DTYPE

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_prn = 32
owned = bd.shard_round_robin(n_prn, world, rank)
assert all(bd.owner_of(s, world) == rank for s in owned)
rng = np.random.default_rng(123)                 # same table on all ranks: the "truth"
truth = np.zeros(n_prn, ACQ_RESULT_DTYPE)
truth["grid_maximum"] = rng.uniform(1, 1e9, n_prn).astype(np.float32)
truth["index_time"] = rng.integers(0, 25000, n_prn)
truth["index_doppler"] = rng.integers(0, 81, n_prn)
local = truth[owned]
# what b200_acq_sweep_best_dev leaves on each rank: its own best record
mine = np.zeros(1, bd.PEAK_DTYPE)
k = int(np.argmax(local["grid_maximum"]))
mine[0] = (local["grid_maximum"][k], owned[k] + 1, local["index_doppler"][k], local["index_time"][k])
allp = bd.allgather_peaks(torch.from_numpy(mine.view(np.int32).copy()))
assert tuple(allp.shape) == (world, 4)
best = bd.best_peak(allp.numpy().view(bd.PEAK_DTYPE))
w = int(np.argmax(truth["grid_maximum"]))
assert int(best["prn"]) == w + 1 and int(best["index_doppler"]) == int(truth["index_doppler"][w]) and int(best["index_time"]) == int(truth["index_time"][w])
assert best["test_statistics"] == truth["grid_maximum"][w]
full = bd.gather_results(local, owned, n_prn)
assert np.array_equal(full, truth)
# band fan-out: rank 0 holds the IQ block, everybody ends up with it
blk = torch.arange(2 * 4096, dtype=torch.float32).reshape(4096, 2) if rank == 0 else torch.zeros((4096, 2))
bd.broadcast_band(blk, src=0)
assert float(blk[-1, 1]) == 2 * 4096 - 1 and float(blk[17, 0]) == 34.0
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gloo_world2_sharding_and_peak_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, B200_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIST_OK" in r.stdout


def test_best_peak_breaks_ties_like_the_reference_scan():
    """Equal statistics: the lowest PRN wins, then the lowest Doppler bin, then the lowest code phase (the reference's
    strict '>' keeps the first maximum); code phases beyond 2^15 (two-level FFT sizes) survive the exchange."""
    from gnss_sdr_b200 import dist as bd
    r = np.zeros(5, bd.PEAK_DTYPE)
    r[0] = (7.5, 9, 40, 199999)
    r[1] = (7.5, 4, 41, 5)
    r[2] = (7.5, 4, 40, 120000)
    r[3] = (7.5, 4, 40, 119999)
    r[4] = (9.0, 0, 0, 0)          # a rank that searched nothing
    b = bd.best_peak(r)
    assert (int(b["prn"]), int(b["index_doppler"]), int(b["index_time"])) == (4, 40, 119999)
    r[0]["test_statistics"] = 7.6
    b = bd.best_peak(r)
    assert int(b["prn"]) == 9 and int(b["index_time"]) == 199999
