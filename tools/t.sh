timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_final_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/r01_final_launches_bench.csv
