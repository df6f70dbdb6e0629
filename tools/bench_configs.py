#!/usr/bin/env python
"""Device-resident throughput of the tracking correlator on the other BASELINE.json tracking configs
(bench.py's headline line is configs[1] = C2).  One JSON line per workload.

  C3  Galileo E1, 64 channels, 50 Msps, N = 200000 (4 ms), sinBOC(1,1) table of 8184 values,
      5 taps VE/E/P/L/VL (dll_pll_veml_tracking.cc:632-636), 1 s of signal per step
  C5p per-GPU share of C5 (256 channels over 8 GPUs = 32 per GPU): 12 GPS L1 (N=50000, L=1023, 3 taps),
      12 Galileo E1 (N=200000, L=8184, 5 taps), 8 GPS L5 (N=50000, L=10230, 3 taps) at 50 Msps
      (mixed tap counts => the per-item kernel's run-time tap dispatch)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(name, fs, groups, seconds, steps=20, warmup=3):
    """groups: list of dict(n_ch, N, L, shifts, table_rate) ; table_rate = table values per second"""
    import torch
    import gnss_sdr_b200.capi as capi
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    eng = capi.Engine(0, st.cuda_stream)
    n_iq = int(fs * seconds)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    iq = torch.randn((n_iq + 16, 2), generator=g, device=dev, dtype=torch.float32)
    eng.iq_attach_dev(0, iq.data_ptr(), n_iq + 16, 0)
    rng = np.random.default_rng(3)
    items = []
    max_taps = max(len(gr["shifts"]) for gr in groups)
    ch_samples = 0
    for gr in groups:
        for c in range(gr["n_ch"]):
            cid = eng.channel_create(0, len(gr["shifts"]))
            eng.channel_set_code(cid, rng.choice([-1.0, 1.0], gr["L"]).astype(np.float32), gr["shifts"])
            n_ep = n_iq // gr["N"]
            doppler = rng.uniform(-5000, 5000)
            step = gr["table_rate"] * (1 + doppler / 1575.42e6) / fs
            k = np.arange(n_ep)
            arr = np.zeros(n_ep, capi.TRK_ITEM_DTYPE)
            arr["channel"] = cid
            arr["n"] = gr["N"]
            arr["sample_index"] = k * gr["N"]
            arr["rem_carrier_phase_rad"] = np.mod(2 * np.pi * doppler / fs * k * gr["N"], 2 * np.pi)
            arr["phase_step_rad"] = 2 * np.pi * doppler / fs
            arr["rem_code_phase_chips"] = -np.mod(rng.uniform(0, gr["L"]) + step * k * gr["N"], gr["L"])
            arr["code_phase_step_chips"] = step
            items.append(arr)
            ch_samples += n_ep * gr["N"]
    items = np.concatenate(items)
    items = items[np.argsort(items["sample_index"], kind="stable")]     # group-friendly order
    it_dev = torch.from_numpy(items.view(np.uint8)).to(dev)
    out = torch.zeros((items.size, max_taps, 2), dtype=torch.float32, device=dev)
    res = {}
    for _ in range(warmup):
        eng.trk_batch_dev(it_dev.data_ptr(), items.size, out.data_ptr(), max_taps, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.trk_batch_dev(it_dev.data_ptr(), items.size, out.data_ptr(), max_taps, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    line = {"workload": name, "kernel_mode": os.environ.get("B200_TRK_SHARED", "auto"), "items": int(items.size),
            "channel_samples_per_step": int(ch_samples), "ms_per_step": ms, "Msamples_per_s": ch_samples / (ms * 1e-3) / 1e6,
            "algorithmic_GBps": ch_samples * 8 / (ms * 1e-3) / 1e9, "iq_bytes": n_iq * 8}
    print(json.dumps(line), flush=True)
    eng.close()
    return line


if __name__ == "__main__":
    e1 = np.array([-0.6, -0.15, 0.0, 0.15, 0.6], np.float32) * 2
    which = sys.argv[1:] or ["C3", "C5p"]
    if "C3" in which:
        run("C3: Galileo E1 64 ch x 50 Msps x 1 s, N=200000, L=8184, 5 taps", 50e6,
            [dict(n_ch=64, N=200000, L=8184, shifts=e1, table_rate=2 * 1.023e6)], 1.0)
    if "C5p" in which:
        run("C5 per-GPU share: 12 GPS L1 + 12 Galileo E1 + 8 GPS L5 at 50 Msps x 1 s", 50e6,
            [dict(n_ch=12, N=50000, L=1023, shifts=[-0.5, 0.0, 0.5], table_rate=1.023e6),
             dict(n_ch=12, N=200000, L=8184, shifts=e1, table_rate=2 * 1.023e6),
             dict(n_ch=8, N=50000, L=10230, shifts=[-0.5, 0.0, 0.5], table_rate=10.23e6)], 1.0)
