# K2L (second-generation long-triple kernel) against the first generation on configs 3 and 5 (one GPU)
mkdir -p gpurun_out
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/l_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-12s %-44s %8.1f GCUPS %9.3f ms  %s cks=%s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path'], d['config']['scores_checksum']))" "$@"; }
: > gpurun_out/l_err.log
{
python tools/tiny_cases.py k2l 2>&1 | tail -2
TAG=old TRIALIGN_K2L=old one --steps 5 --warmup 3 --workload long1024
TAG=new-lag1 one --steps 5 --warmup 3 --workload long1024
TAG=new-lag1-g2 TRIALIGN_K2_GATE=2 one --steps 5 --warmup 3 --workload long1024
TAG=new-lag1-g8 TRIALIGN_K2_GATE=8 one --steps 5 --warmup 3 --workload long1024
TAG=new-lag2 TRIALIGN_K2_LAG=2 one --steps 5 --warmup 3 --workload long1024
TAG=old TRIALIGN_K2L=old one --steps 2 --warmup 2 --workload long8192
TAG=new-lag2 one --steps 2 --warmup 2 --workload long8192
TAG=new-lag1 TRIALIGN_K2_LAG=1 one --steps 2 --warmup 2 --workload long8192
TAG=new-lag2-g8 TRIALIGN_K2_GATE=8 one --steps 2 --warmup 2 --workload long8192
TAG=new-lag2-g32 TRIALIGN_K2_GATE=32 one --steps 2 --warmup 2 --workload long8192
} > gpurun_out/long_perf.txt 2>&1
cat gpurun_out/long_perf.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/long_perf.txt
tail -5 gpurun_out/pytest_gpu.log
tail -5 gpurun_out/l_err.log
