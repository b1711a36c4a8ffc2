mkdir -p gpurun_out
python bench.py --no-cpu --steps 3 --warmup 3 --workload long1024 > gpurun_out/lp_bench.json 2> gpurun_out/lp_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:long2_kernel -c 1 -o gpurun_out/prof_k2l_1024 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload long1024 > gpurun_out/ncu_k2l.log 2>&1
tail -3 gpurun_out/ncu_k2l.log
cat gpurun_out/lp_bench.json | head -c 400
