#!/bin/bash
# validation + A/B of the range pipeline: GPU suite, then the bench with the given numbers of batches in flight
# usage: tools/gpu_run_pipe.sh [in-flight ...]   (default: 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pipe_tall.log 2>&1
tail -3 gpurun_out/pipe_tall.log
for f in ${@:-2}; do
  timeout 900 python bench.py --in-flight $f --no-cpu-baseline --small-compaction 2>/dev/null | grep '^{' > gpurun_out/pipe_bench_if$f.json
  python tools/bench_brief.py gpurun_out/pipe_bench_if$f.json
  python - <<P
import json
d=json.load(open("gpurun_out/pipe_bench_if$f.json"))
print("if$f value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d.get("parity_checked"))
print("fanout_alone", d["fanout_alone"]["us_per_burst"], d["fanout_alone"]["parts_us"])
print("latency", d["latency"]["device_us"], d["latency"]["e2e_us"])
P
done
