# quick development measurements (one gpurun call): parity on tiny cases, then workloads, device-resident value only
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/p_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-8s %-44s %8.1f GCUPS %9.3f ms  %s' % (os.environ.get('TRIALIGN_K2_GATE',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/p_err.log
python tools/tiny_cases.py all 2>&1 | tail -4
for g in 2 4; do
TRIALIGN_K2_GATE=$g one --steps 5 --warmup 3 --workload long1024
done
python tools/time_traceback.py 64 256 512 1024 | tee gpurun_out/traceback_r01.jsonl
tail -5 gpurun_out/p_err.log
