#!/bin/bash
# 2-GPU check: torchrun-launched parity test (NCCL through the C ABI, FIR shards, chain shards) + bench at N=2
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_sharding.py -m gpu -x -q -k "two_ranks" > gpurun_out/gpu2_tests.log 2>&1; rc=$?
tail -4 gpurun_out/gpu2_tests.log
cat gpurun_out/dist_gpu_worker.json 2>/dev/null
[ $rc -ne 0 ] && exit $rc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['roofline']['frac'], d['e2e'])
for k,v in (d.get('extra') or {}).items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)
print(d.get('config5_ofdm'))
PY
tail -3 gpurun_out/bench_2gpu.err
