#!/bin/bash
# A/B of the per-item tracking kernel's build variants (gnss_sdr_b200/variants/*.so) against the production build on ONE box:
# ms per step of C3 (five taps, long tables), the C5 share (mixed) and the distinct-IQ C2 leg (three taps, HBM-bound).
for v in gnss_sdr_b200/libb200gnss.so $(ls gnss_sdr_b200/variants/libb200gnss_u*.so 2>/dev/null) gnss_sdr_b200/libb200gnss.so; do
  B200_LIB=$PWD/$v timeout 200 python tools/bench_configs.py --only C3,C3_track_pilot,C5_per_gpu_share,C2_distinct_iq 2>/dev/null | python -c "
import json,sys
out=[]
for l in sys.stdin:
    d=json.loads(l)
    for k,v in d.items(): out.append(k+' '+str(round(v.get('ms_per_step',0),4)))
print('$v', ' | '.join(out))"
done
