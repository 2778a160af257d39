# 8-GPU call: config 3 DDP at N=8 + timeline, NCCL protocol variant, config 4 (13B ZeRO-2) at N=8, config 5 (34B ZeRO-3 generate --check)
set -x
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 400 $TR --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_e_n8.json 2> gpurun_out/bench_e_n8.err
timeout 400 $TR --master-port 29612 tools/trace_step.py --gpus 8 > gpurun_out/trace_n8.log 2>&1; tail -2 gpurun_out/trace_n8.log | cut -c1-900
NCCL_PROTO=Simple timeout 400 $TR --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_e_n8_simple.json 2> gpurun_out/bench_e_n8_simple.err
timeout 500 $TR --master-port 29614 bench.py --gpus 8 --config 13b-zero2 --micro-batch 4 --steps 6 --warmup 3 > gpurun_out/bench_e_13b_n8.json 2> gpurun_out/bench_e_13b_n8.err
timeout 500 $TR --master-port 29615 tools/zero3_generate.py --layers 60 --new-tokens 16 --check > gpurun_out/zero3_34b_n8_check.log 2>&1
tail -3 gpurun_out/zero3_34b_n8_check.log | cut -c1-600
for f in gpurun_out/bench_e_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['per_gpu'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), d['peak_mem_gb'])
except Exception as e: print('$f ERR', e)
PY
done
