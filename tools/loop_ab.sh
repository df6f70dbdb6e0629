#!/bin/bash
# A/B of the persistent DLL/PLL kernel on ONE box: production build vs variants (loopserial = -DB200_LOOP_SERIAL_UPDATE=1, the whole update on thread 0).
for v in gnss_sdr_b200/libb200gnss.so $(ls gnss_sdr_b200/variants/libb200gnss_loop*.so 2>/dev/null) gnss_sdr_b200/libb200gnss.so; do
  [ -f $v ] || continue
  B200_LIB=$PWD/$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-acq --no-e2e --no-extra 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['closed_loop']; print('$v', 'us/epoch', round(c['us_per_epoch'],2), 'x8', round(c['x8_channels']['us_per_epoch'],2), 'locked', c['channels_locked'])"
done
