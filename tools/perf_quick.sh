# quick development measurements (one gpurun call): parity on tiny cases, then workloads, device-resident value only
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/p_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-8s %-44s %8.1f GCUPS %9.3f ms  %s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/p_err.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
one --steps 3 --warmup 3 --triples 131072
one --steps 2 --warmup 2 --workload long8192
one --steps 5 --warmup 3 --workload long1024
one --steps 2 --warmup 3 --workload batch256 --triples 4096
tail -5 gpurun_out/p_err.log
