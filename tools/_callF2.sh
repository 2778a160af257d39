# 2-GPU call: timeline at N=2, N=2 bench with the collectives serial (no overlap) for comparison
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tools/trace_step.py --gpus 2 > gpurun_out/trace_n2.log 2>&1; tail -2 gpurun_out/trace_n2.log | cut -c1-900
CB_OVERLAP=0 timeout 400 $TR --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_f2_n2_serial.json 2> gpurun_out/bench_f2_n2_serial.err
NCCL_MAX_CTAS=8 timeout 400 $TR --master-port 29523 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_f2_n2_maxctas8.json 2> gpurun_out/bench_f2_n2_maxctas8.err
timeout 400 $TR --master-port 29524 bench.py --gpus 2 --steps 8 --warmup 3 --bucket-mb 1024 > gpurun_out/bench_f2_n2_bucket1g.json 2> gpurun_out/bench_f2_n2_bucket1g.err
for f in gpurun_out/bench_f2_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('$f ERR', e)
PY
done
