set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_r02b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02b.log; tail -5 gpurun_out/pytest_r02b.log
CB_BENCH_SHAPES=1 timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b_default.json 2> gpurun_out/bench_b_default.err; tail -c 300 gpurun_out/bench_b_default.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --max-grad-norm 0 > gpurun_out/bench_b_noclip.json 2> gpurun_out/bench_b_noclip.err
CB_BACKGROUND_OPT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b_nobg.json 2> gpurun_out/bench_b_nobg.err
CB_BENCH_HINTS=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b_nohints.json 2> gpurun_out/bench_b_nohints.err
timeout 400 python bench.py --config 7b-clip-mlp --steps 5 --warmup 3 > gpurun_out/bench_b_7b.json 2> gpurun_out/bench_b_7b.err
for f in gpurun_out/bench_b_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['peak_mem_gb'], d['clocks']['sm_mhz'], d['roofline']['frac'], d.get('roofline_adamw'))
except Exception as e: print('ERR', e)
"; done
