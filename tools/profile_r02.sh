#!/bin/bash
# Round-2 ncu evidence (one GPU).  Each capture runs only after the same command has exited 0 without ncu.
#   bash tools/profile_r02.sh            -> gpurun_out/r02_*.ncu-rep, r02_launches.csv
# Read the reports on the build box with tools/ncu_export.sh, which also regenerates profiles/ncu_kernel_stats.json.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-secondary"
set -x
$B --steps 2 --warmup 3 > gpurun_out/r02_plain_k1.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_plain_full.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_batch -s 3 -c 1 -o gpurun_out/r02_k1 -f $B --steps 1 --warmup 3 > gpurun_out/r02_ncu_k1.log 2>&1
$B --steps 1 --warmup 3 --workload batch256 --triples 4096 > gpurun_out/r02_plain_k2s.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fast_stream_kernel -s 3 -c 1 -o gpurun_out/r02_k2s -f $B --steps 1 --warmup 3 --workload batch256 --triples 4096 > gpurun_out/r02_ncu_k2s.log 2>&1
$B --steps 1 --warmup 3 --workload long1024 > gpurun_out/r02_plain_k2l1024.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:long2_kernel -s 3 -c 1 -o gpurun_out/r02_k2l_1024 -f $B --steps 1 --warmup 3 --workload long1024 > gpurun_out/r02_ncu_k2l1024.log 2>&1
$B --steps 1 --warmup 3 --workload long8192 > gpurun_out/r02_plain_k2l8192.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:long2_kernel -s 3 -c 1 -o gpurun_out/r02_k2l_8192 -f $B --steps 1 --warmup 3 --workload long8192 > gpurun_out/r02_ncu_k2l8192.log 2>&1
set +x
ls -la gpurun_out/r02_*.ncu-rep
tail -n 2 gpurun_out/r02_ncu_k1.log gpurun_out/r02_ncu_k2s.log gpurun_out/r02_ncu_k2l1024.log gpurun_out/r02_ncu_k2l8192.log | cut -c1-200
