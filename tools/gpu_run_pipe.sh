#!/bin/bash
# submit/collect validation: the new test, the whole GPU suite, then the bench with 1 and 2 batches in flight
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "submit_collect or prefetch" > gpurun_out/pipe_t1.log 2>&1
tail -5 gpurun_out/pipe_t1.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pipe_tall.log 2>&1
tail -3 gpurun_out/pipe_tall.log
for f in 1 2 3; do
  timeout 600 python bench.py --in-flight $f --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > gpurun_out/pipe_bench_if$f.json
  python tools/bench_brief.py gpurun_out/pipe_bench_if$f.json
  python - <<P
import json
d=json.load(open("gpurun_out/pipe_bench_if$f.json"))
print("if$f value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d.get("parity_checked"))
P
done
