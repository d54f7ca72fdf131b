#!/bin/bash
# first GPU pass of a build: fan-out tests (guarded against a hung cooperative kernel), the whole GPU suite, the bench
# line, and the serialised ncu launch list of a short bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== fanout tests"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fanout or config3 or watch" 2>&1 | tail -15 | tee gpurun_out/t_fanout.log
echo "== all gpu tests"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/t_gpu.log
echo "== bench"; timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_checked','gpu_launches')}, d['e2e'], d.get('latency'), d.get('fanout_alone'))
    print('compaction', {k:v for k,v in d['extra']['compaction'].items() if k!='kernels'})
    for k in d['kernels'][:12]: print(k['name'], round(k['avg_us'],1), round(k['achieved_gbs'] or 0,1), round(k['share'],3))
    print(d['host_call_us']); print(d['cpu_baseline'])
except Exception as e: print('bench parse failed', e)
PY
echo "== ncu launch list"
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
