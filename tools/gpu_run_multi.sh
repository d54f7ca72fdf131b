#!/bin/bash
# multi-GPU pass: usage  tools/gpu_run_multi.sh N [strong]   (run with gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
make -C kubebrain_b200/csrc 2>&1 | tail -1
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== 2-GPU cursor test"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -k two_gpu 2>&1 | tail -4 | tee gpurun_out/t_2gpu.log
fi
echo "== weak N=$N"
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_n$N.err
if [ "${2:-}" = "strong" ]; then
  echo "== strong N=$N"
  timeout -s KILL 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --mode strong > gpurun_out/bench_strong_n$N.json 2> gpurun_out/bench_strong_n$N.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_strong_n$N.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*_n$N.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','parity_checked','scaling','numa')}, d['e2e'], d.get('cursor_exchange_us'), d.get('parity'))
    except Exception as e: print(f, 'parse failed', e)
PY
