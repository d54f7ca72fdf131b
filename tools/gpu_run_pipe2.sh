#!/bin/bash
# register-budget A/B: GPU suite, then the bench (2 in flight, with extras) and the joined loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pipe2_tall.log 2>&1
tail -3 gpurun_out/pipe2_tall.log
for f in 2 1; do
  timeout 900 python bench.py --in-flight $f --no-cpu-baseline --small-compaction 2>/dev/null | grep '^{' > gpurun_out/pipe2_bench_if$f.json
  python tools/bench_brief.py gpurun_out/pipe2_bench_if$f.json
  python - <<P
import json
d=json.load(open("gpurun_out/pipe2_bench_if$f.json"))
print("if$f value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d.get("parity_checked"))
print("fanout_alone", d["fanout_alone"]["us_per_burst"], d["fanout_alone"]["parts_us"])
print("latency", d["latency"]["device_us"], d["latency"]["e2e_us"])
print("wire", {k:round(v["avg_us"],1) for k,v in d["extra"]["wire"]["kernels"].items()})
print("compaction", d["extra"]["compaction"]["ms_per_sweep"], d["extra"]["compaction"]["records_per_s"])
P
done
