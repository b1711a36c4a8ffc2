# round-end measurement set on one GPU: tests, smoke, default bench line, reference arm, the other BASELINE configs at full size
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_batch64.json 2> gpurun_out/final_batch64.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
python bench.py --no-cpu --steps 5 --warmup 3 --workload long1024 > gpurun_out/final_long1024.json 2> gpurun_out/final_long1024.err
python bench.py --no-cpu --steps 3 --warmup 3 --workload long8192 > gpurun_out/final_long8192.json 2> gpurun_out/final_long8192.err
python bench.py --no-cpu --steps 2 --warmup 3 --workload batch256 > gpurun_out/final_batch256.json 2> gpurun_out/final_batch256.err
set +x
for f in gpurun_out/final_*.json; do python -c "import sys,json; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], d['value'], d.get('ms_per_step'), d.get('e2e',{}).get('value'))" $f; done
tail -n 3 gpurun_out/final_*.err
