#!/bin/bash
# full GPU check of the round: every -m gpu test, smoke, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; rc=$?
tail -5 gpurun_out/gpu_tests.log
[ $rc -ne 0 ] && exit $rc
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'] if d.get('e2e') else None, d['clocks'])
for k,v in (d.get('extra') or {}).items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)
print(d.get('config5_ofdm'))
PY
tail -3 gpurun_out/bench_default.err
