#!/bin/bash
# How the round-2 numbers in profiles/ were produced (each block is one gpurun call; N = GPUs of the box).
#   1 GPU :  bash tools/measure_r02.sh 1
#   N GPUs:  bash tools/measure_r02.sh N         (inside  gpurun --gpus N)
# Outputs land in gpurun_out/ (r02_bench_<N>gpu.json ...); copy what is to be kept to profiles/.
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -3
  python __graft_entry__.py smoke 2>&1 | tail -1
  python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> /dev/null
  if [ "${FULL:-0}" = 1 ]; then   # the slower side measurements (kept in profiles/ from earlier calls of the round)
    python tools/time_literal.py > gpurun_out/r02_fallbacks.jsonl 2>&1
    python tools/time_traceback.py 256 512 1024 > gpurun_out/r02_traceback.jsonl 2>&1
    python tools/time_slabs.py > gpurun_out/r02_time_slabs.jsonl 2>&1
    python tools/time_cols.py > gpurun_out/r02_time_cols.jsonl 2>&1
  fi
  python tools/stress_gpu.py ${STRESS:-30} 2>&1 | tail -1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
  python -m pytest tests/test_gpu_long.py -m gpu -q -k "multi or mask or split or pipelines" 2>&1 | tail -3
  $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
  for m in slab cols; do   # the two decompositions side by side
    $TR bench.py --gpus $N --steps 3 --warmup 3 --workload long8192 --opt multi=$m > gpurun_out/r02_long8192_${m}_${N}gpu.json 2> /dev/null
  done
fi
