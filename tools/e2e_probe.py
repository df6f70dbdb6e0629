"""Where does the cf32 e2e step lose time against a plain 200 MB host->device copy?  Variants of the bench's e2e leg."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
import gnss_sdr_b200.capi as capi  # noqa: E402
import gnss_synth  # noqa: E402

dev = torch.device("cuda", 0)
N_CH, EPOCH, N_EPOCHS, TAPS = bench.N_CH, bench.EPOCH, bench.N_EPOCHS, bench.TAPS
codes = {p: gnss_synth.gps_ca_code(p) for p in range(1, N_CH + 1)}
svs = bench.svs_for_rank(0)
n_iq = EPOCH * N_EPOCHS
host_iq = torch.empty((n_iq, 2), dtype=torch.float32, pin_memory=True)
host_iq.normal_()
eng = capi.Engine(0)
eng.iq_create(1, 2 * n_iq)
cids = []
for sv in svs:
    cid = eng.channel_create(1, TAPS)
    eng.channel_set_code(cid, codes[sv["prn"]], bench.SHIFTS)
    cids.append(cid)
items = bench.build_items(capi, svs, cids, 0)
base_idx = items["sample_index"].copy()
items_v = items.reshape(N_EPOCHS, N_CH)


def run(chunks, submit, steps=20, pipelined=True):
    ep = N_EPOCHS // chunks

    def enqueue():
        t = []
        first0 = None
        for c in range(chunks):
            a, b = c * ep, (c + 1) * ep if c < chunks - 1 else N_EPOCHS
            first = eng.iq_push_ptr(1, host_iq.data_ptr() + a * EPOCH * 8, (b - a) * EPOCH)
            if first0 is None:
                first0 = first
                items["sample_index"] = base_idx + np.uint64(first0)
            if submit:
                t.append(eng.trk_submit(items_v[a:b].reshape(-1), TAPS))
        return t

    def collect(t):
        for x in t:
            eng.trk_wait(x)

    for _ in range(2):
        collect(enqueue())
    eng.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host_t = 0.0
    prev = None
    for _ in range(steps):
        h0 = time.perf_counter()
        cur = enqueue()
        host_t += time.perf_counter() - h0
        if pipelined:
            if prev is not None:
                collect(prev)
            prev = cur
        else:
            collect(cur)
    if pipelined and prev is not None:
        collect(prev)
    eng.sync()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"chunks={chunks} submit={submit} pipelined={pipelined}: {dt * 1e3:.3f} ms/step ({n_iq * 8 / dt / 1e9:.1f} GB/s), host enqueue {host_t / steps * 1e3:.3f} ms/step", flush=True)


run(8, True, steps=4)   # touches all 16 slots (allocations) before anything is timed
for chunks in (1, 8):
    run(chunks, False)
for chunks in (1, 8):
    run(chunks, True)
