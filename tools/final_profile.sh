# round-end ncu evidence (one GPU): launch list of the default bench command, full capture of the dominant kernel
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k1_r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_batch -s 3 -c 1 -o gpurun_out/prof_k1_r1d -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_k1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_batch -s 2 -c 1 -o gpurun_out/prof_k2batch_r1d -f python bench.py --steps 1 --warmup 2 --no-cpu --workload batch256 --triples 4096 > gpurun_out/ncu_k2bd.log 2>&1
tail -n 2 gpurun_out/ncu_k1d.log gpurun_out/ncu_k2bd.log | cut -c1-300
