# round-end ncu evidence (one GPU): launch list of the default bench command, full captures of the dominant kernels
mkdir -p gpurun_out
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k1_r01e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_batch -s 3 -c 1 -o gpurun_out/prof_k1_r1e -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_k1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_stream -s 2 -c 1 -o gpurun_out/prof_k2s_r1e -f python bench.py --steps 1 --warmup 2 --no-cpu --workload batch256 --triples 4096 > gpurun_out/ncu_k2se.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:long2_kernel -s 1 -c 1 -o gpurun_out/prof_k2l_1024_r1e -f python bench.py --steps 1 --warmup 1 --no-cpu --workload long1024 > gpurun_out/ncu_k2l1024e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:long2_kernel -s 1 -c 1 -o gpurun_out/prof_k2l_8192_r1e -f python bench.py --steps 1 --warmup 1 --no-cpu --workload long8192 > gpurun_out/ncu_k2l8192e.log 2>&1
set +x
tail -n 2 gpurun_out/ncu_k1e.log gpurun_out/ncu_k2se.log gpurun_out/ncu_k2l1024e.log gpurun_out/ncu_k2l8192e.log | cut -c1-300
