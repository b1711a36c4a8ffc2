mkdir -p gpurun_out
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/u_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-10s %-50s %8.1f GCUPS %9.3f ms  %s cks=%s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path'], d['config']['scores_checksum']))" "$@"; }
: > gpurun_out/u_err.log
{
TAG=k1 one --steps 3 --warmup 3 --triples 262144
TAG=k2s@64 TRIALIGN_K1=stream one --steps 3 --warmup 3 --triples 262144
} > gpurun_out/k1s_perf.txt 2>&1
cat gpurun_out/k1s_perf.txt; tail -3 gpurun_out/u_err.log
