#!/bin/bash
# N-GPU bench (N = number of visible GPUs): every named shape at N ranks, config 3 as N stream shards, config 5 reduced through the C-ABI communicator
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
N=$N python - <<'PY'
import json, os
n=os.environ['N']
d=json.load(open('gpurun_out/bench_%sgpu.json'%n))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['roofline']['frac'], d['e2e'])
for k,v in (d.get('extra') or {}).items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)
print(d.get('config5_ofdm'))
PY
tail -3 gpurun_out/bench_${N}gpu.err
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
