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
keys = bd.pack_peak_key(local["grid_maximum"], np.array(owned) + 1, local["index_doppler"], local["index_time"])
best = torch.tensor([keys.max()], dtype=torch.int64)
bd.allreduce_best_peak(best)
peak, prn, d, t = bd.unpack_peak_key(best.numpy())
w = int(np.argmax(truth["grid_maximum"]))
assert int(prn[0]) == w + 1 and int(d[0]) == int(truth["index_doppler"][w]) and int(t[0]) == int(truth["index_time"][w])
assert peak[0] == truth["grid_maximum"][w]
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


def test_key_order_preserving():
    from gnss_sdr_b200 import dist as bd
    peaks = np.array([0.0, 1e-30, 1.0, 1.0000001, 3.5e9, 3.4e38], np.float32)
    keys = bd.pack_peak_key(peaks, [1] * 6, [0] * 6, [0] * 6)
    assert np.all(np.diff(keys) > 0)
    p, prn, d, t = bd.unpack_peak_key(bd.pack_peak_key([2.5], [17], [80], [24999]))
    assert (p[0], prn[0], d[0], t[0]) == (2.5, 17, 80, 24999)
