#!/usr/bin/env python
"""Device-resident throughput of the tracking correlator on the other BASELINE.json tracking configs (C3 with and without
the pilot's data tap, the per-GPU share of C5, C2 with distinct samples per channel-epoch).  One JSON line per workload;
bench.py reports the same figures under `other_configs`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    import torch
    import bench
    import gnss_sdr_b200.capi as capi
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    for k, v in bench.other_configs(torch, capi, dev, st, steps=20, only=only).items():
        print(json.dumps({k: v}), flush=True)
