# quick development measurements (one gpurun call): parity on tiny cases, then workloads, device-resident value only
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/p_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-8s %-44s %8.1f GCUPS %9.3f ms  %s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/p_err.log
python tools/sanitize_case.py k1 2>&1 | tail -1
TAG=perm one --steps 3 --warmup 3 --triples 131072
TAG=perm-m TRIALIGN_K1=256m one --steps 3 --warmup 3 --triples 131072
ALT=$PWD/hw-accelerator-three-sequence-alignment_b200/lib_alt/libtrialign_diag.so
TRIALIGN_LIB=$ALT python tools/sanitize_case.py k1 2>&1 | tail -1
TAG=diag TRIALIGN_LIB=$ALT one --steps 3 --warmup 3 --triples 131072
TAG=diag-m TRIALIGN_K1=256m TRIALIGN_LIB=$ALT one --steps 3 --warmup 3 --triples 131072
tail -5 gpurun_out/p_err.log
