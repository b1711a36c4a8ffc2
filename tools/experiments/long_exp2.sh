mkdir -p gpurun_out
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/q_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-14s %-44s %8.1f GCUPS %9.3f ms  %s cks=%s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path'], d['config']['scores_checksum']))" "$@"; }
: > gpurun_out/q_err.log
{
TAG=lag2 TRIALIGN_K2_LAG=2 one --steps 2 --warmup 2 --workload long8192
TAG=gate8 TRIALIGN_K2_GATE=8 one --steps 2 --warmup 2 --workload long8192
TAG=gate32 TRIALIGN_K2_GATE=32 one --steps 2 --warmup 2 --workload long8192
TAG=gate8 TRIALIGN_K2_GATE=8 one --steps 5 --warmup 3 --workload long1024
TAG=gate2 TRIALIGN_K2_GATE=2 one --steps 5 --warmup 3 --workload long1024
} > gpurun_out/long_exp2.txt 2>&1
cat gpurun_out/long_exp2.txt; tail -3 gpurun_out/q_err.log
