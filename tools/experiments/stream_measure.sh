# K2S (streaming block sweep) against the per-block kernel on config 4 (reduced), then tests, then one ncu capture
mkdir -p gpurun_out
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/s_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-10s %-50s %8.1f GCUPS %9.3f ms  %s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/s_err.log
{
python tools/tiny_cases.py 2>&1 | tail -3
TAG=stream one --steps 3 --warmup 3 --workload batch256 --triples 8192
TAG=sweep TRIALIGN_K2=sweep one --steps 3 --warmup 3 --workload batch256 --triples 8192
TAG=stream one --steps 2 --warmup 3 --workload batch256 --triples 32768
} > gpurun_out/stream_perf.txt 2>&1
cat gpurun_out/stream_perf.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/stream_perf.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_stream -c 1 -o gpurun_out/prof_k2s_r1 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload batch256 --triples 4096 > gpurun_out/ncu_k2s.log 2>&1
tail -3 gpurun_out/ncu_k2s.log
tail -5 gpurun_out/s_err.log
