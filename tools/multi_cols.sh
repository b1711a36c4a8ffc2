# column pipeline vs i-slab pipeline on N GPUs: bash tools/multi_cols.sh N  (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
TRIALIGN_MULTI=cols run --steps 3 --warmup 3 --no-cpu --workload long8192 > gpurun_out/mc_${N}gpu_cols.json 2> gpurun_out/mc_${N}gpu_cols.err
if [ "$2" != "colsonly" ]; then
TRIALIGN_MULTI=slab run --steps 3 --warmup 3 --no-cpu --workload long8192 > gpurun_out/mc_${N}gpu_slab.json 2> gpurun_out/mc_${N}gpu_slab.err
fi
for f in gpurun_out/mc_${N}gpu_*.json; do python -c "import sys,json; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],1), round(d.get('ms_per_step',0),2), d.get('scaling'), d['config'].get('scores_checksum'), d['config']['parallelism'])" $f; done
tail -n 4 gpurun_out/mc_${N}gpu_cols.err
