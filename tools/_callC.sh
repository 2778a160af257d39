set -x
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_r02c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02c.log; tail -5 gpurun_out/pytest_r02c.log
CB_BENCH_SHAPES=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc1.json 2> gpurun_out/bench_c_clc1.err
CB_GEMM_CLC=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc0.json 2> gpurun_out/bench_c_clc0.err
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc1b.json 2> gpurun_out/bench_c_clc1b.err
CB_GEMM_CLC=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc0b.json 2> gpurun_out/bench_c_clc0b.err
for f in gpurun_out/bench_c_*.json; do python - <<PY
import json
try:
    d=json.load(open('$f')); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('$f ERR', e)
PY
done
