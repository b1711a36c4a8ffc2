# quick development measurements (one gpurun call)
one() { python bench.py --no-cpu "$@" 2>> gpurun_out/p_err.log | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('%-18s %-44s %8.1f GCUPS %9.3f ms  %s' % (os.environ.get('TAG',''), ' '.join(sys.argv[1:]), d['value'], d['ms_per_step'], d['config']['kernel_path']))" "$@"; }
: > gpurun_out/p_err.log
D=$PWD/hw-accelerator-three-sequence-alignment_b200/lib_alt
for v in hint_oldborder hint_newborder; do
TRIALIGN_LIB=$D/libtrialign_$v.so python tools/tiny_cases.py k1 | tail -1
TAG=$v TRIALIGN_LIB=$D/libtrialign_$v.so one --steps 3 --warmup 3 --triples 131072
TAG=$v-256m TRIALIGN_K1=256m TRIALIGN_LIB=$D/libtrialign_$v.so one --steps 3 --warmup 3 --triples 131072
done
tail -5 gpurun_out/p_err.log
