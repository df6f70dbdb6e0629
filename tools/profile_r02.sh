#!/bin/bash
# Round-2 ncu evidence (run under gpurun, ONE GPU): launch list of the bench command + one --set full capture per main kernel.
# Outputs under gpurun_out/; summaries are copied into profiles/ by hand (profiles/README.md).
set -u
mkdir -p gpurun_out
K='regex:trk_|acq_|convert_|loop'
# (1) launch list of the default bench command (value + e2e legs; no extras, no CPU leg: same step, shorter run)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r02_launch_list.csv \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_launch_bench.log 2>&1
# (2) the dominant tracking kernel (C2 step), the per-item kernel on the distinct layout, the acquisition row kernel, the loop kernel
ncu --set full --clock-control none --import-source on -k regex:trk_shared_kernel -s 4 -c 1 -o gpurun_out/r02_trk_shared -f \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:trk_correlate_kernel -s 2 -c 1 -o gpurun_out/r02_trk_item_distinct -f \
    python tools/bench_configs.py --only C2_distinct_iq > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:acq_corr_kernel -s 4 -c 1 -o gpurun_out/r02_acq_corr -f \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:trk_loop_persistent -c 1 -o gpurun_out/r02_loop -f \
    python tools/loop_profile.py 200 > /dev/null 2>&1
for r in r02_trk_shared r02_trk_item_distinct r02_acq_corr r02_loop; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page details > gpurun_out/${r}_ncu_details.txt 2>&1
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>&1
done
ls -la gpurun_out/*.ncu-rep
