#!/bin/bash
# evidence for profiles/: the default bench line, the reference arm, and (after each plain run exited 0) ncu: the launch
# list of the bench command, --set full of the headline kernel and of every kernel at 2^26.  The .ncu-rep files are
# summarised on the box (they exceed the 64 MiB that travel back).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "ref rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'])
e=d.get('extra') or {}
print(e.get('error'))
for k in ('fir64_overlap_save','fir64_stream_sharded','correlator1024','modem_fused_1000000sym','modem_fused_1000000sym_graph16'): print(k, e.get(k))
"
timeout 300 python tools/chain_quick.py 262144 > gpurun_out/chain_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_x2 -s 3 -c 1 -o gpurun_out/chain_x2_final python tools/chain_quick.py 262144 > gpurun_out/ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/chain_x2_final.ncu-rep > gpurun_out/chain_x2_final_ncu_summary.txt 2>&1
python tools/ncu_hot.py gpurun_out/chain_x2_final.ncu-rep 25 > gpurun_out/chain_x2_final_hotspots.txt 2>&1
timeout 300 python tools/profile_all.py > gpurun_out/profile_all_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -o gpurun_out/all_kernels python tools/profile_all.py > gpurun_out/ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/all_kernels.ncu-rep > gpurun_out/all_kernels_ncu_summary.txt 2>&1
rm -f gpurun_out/all_kernels.ncu-rep
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_short.json 2>/dev/null &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu3.log 2>&1
du -sh gpurun_out; ls gpurun_out | head -40
