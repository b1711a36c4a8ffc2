# quick development measurements (one gpurun call): parity on tiny cases, then the four workloads, device-resident value only
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/p_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-40s %8.1f GCUPS %9.3f ms  %s' % (' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/p_err.log
python tools/sanitize_case.py all 2>&1 | tail -5
one --steps 3 --warmup 3 --triples 131072
TRIALIGN_K1_BAR=sync one --steps 3 --warmup 3 --triples 131072
one --steps 2 --warmup 3 --workload long8192
one --steps 5 --warmup 3 --workload long1024
one --steps 2 --warmup 3 --workload batch256 --triples 4096
tail -5 gpurun_out/p_err.log
