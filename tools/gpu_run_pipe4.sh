#!/bin/bash
# lanes A/B: batches in flight 2/3/4 on 4 lanes; co-resident decode + gather variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_cpp_host.py -m gpu -x -q > gpurun_out/pipe4_t.log 2>&1
tail -3 gpurun_out/pipe4_t.log
run() {  # name, env..., in-flight
  name=$1; shift; f=$1; shift
  env "$@" KB_LANES=4 timeout 900 python bench.py --in-flight $f --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > gpurun_out/pipe4_$name.json
  python tools/bench_brief.py gpurun_out/pipe4_$name.json
  python - <<P
import json
d=json.load(open("gpurun_out/pipe4_$name.json"))
print("$name value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d.get("parity_checked"))
P
}
run if3 3 X=1
run if4 4 X=1
run if2 2 X=1
run cores3 3 KB_GATHER_CTAS=1 KB_DECODE_WARPS=5
run cores3b 3 KB_GATHER_CTAS=1 KB_DECODE_WARPS=6
