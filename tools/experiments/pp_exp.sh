export SHAPES=short
{
TRIALIGN_K2_PP=0 python tools/time_long_shapes.py
TRIALIGN_K2_PP=1 python tools/time_long_shapes.py
} > gpurun_out/pp_exp.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/pp_exp.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print("%-28s %-20s %8.4f ms %7.1f GCUPS" % (d['tag'], d['shape'], d['ms'], d['gcups']))
PY
TRIALIGN_K2_PP=0 bash tools/long_quick.sh pp0
