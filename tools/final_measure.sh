#!/bin/bash
# round-end pass on the GPU box (outputs under gpurun_out/, copied to profiles/ by hand):
# pytest -m gpu, smoke(), the reference arm, the bench line, the serialised ncu launch list and one `--set full` capture
# of the three kernels that matter (k_gather, k_decode_lcp, k_fanout) + its stall summaries
set -u
mkdir -p gpurun_out
make -C kubebrain_b200/csrc 2>&1 | tail -1
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout -s KILL 1200 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout -s KILL 600 python bench.py --in-flight 1 --no-cpu-baseline --no-extras > gpurun_out/final_bench_if1.json 2> gpurun_out/final_bench_if1.err  # A/B: one kb_range_batch per step
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/final_ncu_launch.log 2>&1
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:"k_decode_lcp|k_gather$|k_fanout" -c 6 -o gpurun_out/final_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-parity --serial > gpurun_out/final_ncu_full.log 2>&1
for k in k_decode_lcp k_gather k_fanout; do python tools/ncu_stalls.py gpurun_out/final_prof.ncu-rep $k 25 > gpurun_out/final_ncu_stalls_$k.txt 2>&1; done
ncu -i gpurun_out/final_prof.ncu-rep --page raw --csv > gpurun_out/final_ncu_full.csv 2>/dev/null
tail -2 gpurun_out/final_pytest_gpu.log; tail -1 gpurun_out/final_smoke.log; head -c 300 gpurun_out/final_bench.json; echo; head -c 300 gpurun_out/final_bench_reference.json; echo; head -8 gpurun_out/final_ncu_stalls_k_fanout.txt
