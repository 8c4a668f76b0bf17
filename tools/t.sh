timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 800 python bench.py --no-cpu > gpurun_out/bench_g3.json 2> gpurun_out/bench_g3.err; tail -2 gpurun_out/bench_g3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_g3.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e'])"
