set -x
timeout 300 python tools/probe1.py gemm_clc > gpurun_out/probe_clc.log 2>&1; echo "clc rc=$?" >> gpurun_out/probe_clc.log; tail -25 gpurun_out/probe_clc.log
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_r02c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02c.log; tail -5 gpurun_out/pytest_r02c.log
CB_BENCH_SHAPES=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc1.json 2> gpurun_out/bench_c_clc1.err
CB_GEMM_CLC=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc0.json 2> gpurun_out/bench_c_clc0.err
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc1b.json 2> gpurun_out/bench_c_clc1b.err
CB_GEMM_CLC=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c_clc0b.json 2> gpurun_out/bench_c_clc0b.err
timeout 300 python tools/decode_bench.py --batch 1 > gpurun_out/decode_b1.json 2> gpurun_out/decode_b1.err; cat gpurun_out/decode_b1.json
timeout 300 python tools/decode_bench.py --batch 8 > gpurun_out/decode_b8.json 2> gpurun_out/decode_b8.err; cat gpurun_out/decode_b8.json
for f in gpurun_out/bench_c_*.json; do python - <<PY
import json
try:
    d=json.load(open('$f')); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('$f ERR', e)
PY
done
