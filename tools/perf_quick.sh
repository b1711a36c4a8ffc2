# quick development measurements (one gpurun call): parity on tiny cases, then workloads, device-resident value only
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/p_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-6s %-44s %8.1f GCUPS %9.3f ms  %s' % (os.environ.get('TRIALIGN_K1',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/p_err.log
for v in 256m 256l; do
  TRIALIGN_K1=$v python tools/sanitize_case.py k1 2>&1 | tail -1
  TRIALIGN_K1=$v one --steps 3 --warmup 3 --triples 131072
done
tail -5 gpurun_out/p_err.log
