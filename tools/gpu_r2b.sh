#!/bin/bash
# round 2, second half: composite radices in the any-length FFT, QPSK fast path of the fused modem, tensor-path operand-prep ubench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; rc=$?
tail -5 gpurun_out/gpu_tests.log
[ $rc -ne 0 ] && exit $rc
{ echo "composite radices"; timeout 200 python tools/fft_odd_quick.py; echo "prime radices only"; AE_FFT_NO_COMPOSITE=1 timeout 200 python tools/fft_odd_quick.py; } > gpurun_out/fft_odd.log 2>&1
cat gpurun_out/fft_odd.log
timeout 200 build/tc_operand_prep > gpurun_out/tc_operand_prep.txt 2>&1; cat gpurun_out/tc_operand_prep.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'] if d.get('e2e') else None, d['clocks'])
for k,v in (d.get('extra') or {}).items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)
PY
tail -3 gpurun_out/bench_default.err
