# multi-GPU measurements: bash tools/final_multi.sh N   (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 3 --warmup 3 --no-cpu --workload long8192 > gpurun_out/final_${N}gpu_long8192.json 2> gpurun_out/final_${N}gpu_long8192.err
run --steps 3 --warmup 3 --no-cpu > gpurun_out/final_${N}gpu_batch64.json 2> gpurun_out/final_${N}gpu_batch64.err
run --steps 2 --warmup 3 --no-cpu --workload batch256 --triples 32768 > gpurun_out/final_${N}gpu_batch256.json 2> gpurun_out/final_${N}gpu_batch256.err
for f in gpurun_out/final_${N}gpu_*.json; do python -c "import sys,json; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],1), round(d.get('ms_per_step',0),2), d.get('scaling'), d['config'].get('scores_checksum'), d.get('e2e',{}).get('value'))" $f; done
tail -n 3 gpurun_out/final_${N}gpu_long8192.err
