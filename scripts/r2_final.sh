#!/bin/bash
# final GPU evidence pass of round 2 (1 GPU): tests, bench lines, ncu launch lists.  Outputs -> gpurun_out/
mkdir -p gpurun_out
(timeout 200 python -m pytest tests -m gpu -q --timeout 100 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log)
timeout 300 python bench.py > gpurun_out/r2_bench_m7_q8_0_n1.json 2> gpurun_out/bench_q8.err
timeout 150 python bench.py --wtype q4_0 --no-prefill > gpurun_out/r2_bench_m7_q4_0_n1.json 2> gpurun_out/bench_q4.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec|attn|embed" -c 340 --csv --log-file gpurun_out/r2_launches_bench_m7_q8_0.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-prefill > /dev/null 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec|attn|embed" -c 340 --csv --log-file gpurun_out/r2_launches_bench_m7_q4_0.csv python bench.py --wtype q4_0 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-prefill > /dev/null 2>&1
tail -2 gpurun_out/r2_pytest_gpu.log; head -c 400 gpurun_out/r2_bench_m7_q8_0_n1.json; echo; head -c 300 gpurun_out/r2_bench_m7_q4_0_n1.json; echo; wc -l gpurun_out/r2_launches_*.csv
