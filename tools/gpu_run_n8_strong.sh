make -C kubebrain_b200/csrc | tail -1
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --mode strong > gpurun_out/bench_strong_n8.json 2> gpurun_out/bench_strong_n8.err; echo rc=$?
python tools/bench_brief.py gpurun_out/bench_strong_n8.json
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_strong_n8.json') if l.startswith('{')][-1])
print(d['e2e'], d['parity'], d.get('cursor_exchange_us'))
PY
