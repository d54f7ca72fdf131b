set -x
timeout 1300 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/final_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_decode_lcp|k_gather$" -c 4 -o gpurun_out/final_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_pytest_gpu.log; cat gpurun_out/final_smoke.log | tail -2; head -c 400 gpurun_out/final_bench.json; echo; head -c 300 gpurun_out/final_bench_reference.json
