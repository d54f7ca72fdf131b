#!/bin/bash
# priority split A/B (KB_PRIO_SPLIT=0/1), 2 batches in flight
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_cpp_host.py -m gpu -x -q > gpurun_out/pipe3_t.log 2>&1
tail -3 gpurun_out/pipe3_t.log
for s in 1 0 1; do
  KB_PRIO_SPLIT=$s timeout 900 python bench.py --in-flight 2 --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > gpurun_out/pipe3_bench_s$s.json
  python tools/bench_brief.py gpurun_out/pipe3_bench_s$s.json
  python - <<P
import json
d=json.load(open("gpurun_out/pipe3_bench_s$s.json"))
print("split$s value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d.get("parity_checked"))
print("latency", d["latency"]["device_us"], d["latency"]["e2e_us"])
P
done
