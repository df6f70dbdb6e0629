"""Small closed-loop run for ncu launch lists: 32 channels of the C2 band, 40 epochs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
import gnss_sdr_b200.capi as capi  # noqa: E402
import gnss_synth  # noqa: E402

n_ep = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
codes = {p: gnss_synth.gps_ca_code(p) for p in range(1, bench.N_CH + 1)}
svs = bench.svs_for_rank(0)
n_iq = bench.EPOCH * (n_ep + 4)
iq = bench.synth_iq_device(torch, codes, svs, n_iq + 16, 2, dev)
eng = capi.Engine(0)
eng.iq_attach_dev(0, iq.data_ptr(), n_iq + 16, 0)
cids = []
for sv in svs:
    cid = eng.channel_create(0, 3)
    eng.channel_set_code(cid, codes[sv["prn"]], bench.SHIFTS)
    cids.append(cid)
print(bench.bench_closed_loop(capi, eng, svs, cids, n_epochs=n_ep, repeats=1))
