export SHAPES=short
{
python tools/time_long_shapes.py
TRIALIGN_K2_EARLY=1 python tools/time_long_shapes.py
TRIALIGN_K2_GATETID=160 python tools/time_long_shapes.py
TRIALIGN_K2_GATETID=224 python tools/time_long_shapes.py
TRIALIGN_K2_GATETID=0 python tools/time_long_shapes.py
TRIALIGN_K2_GATETID=160 TRIALIGN_K2_GATE=2 python tools/time_long_shapes.py
TRIALIGN_K2_GATETID=160 TRIALIGN_K2_GATE=8 python tools/time_long_shapes.py
} > gpurun_out/long_exp.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/long_exp.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print("%-28s %-20s %8.4f ms %7.1f GCUPS" % (d['tag'], d['shape'], d['ms'], d['gcups']))
PY
