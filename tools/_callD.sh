# 2-GPU call: NCCL ZeRO-2 test, config 3 at N=2 (CLC on / off), config 4 (13B ZeRO-2) at N=2
set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout=800 -p no:cacheprovider > gpurun_out/pytest_multi_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_gpu.log; tail -15 gpurun_out/pytest_multi_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_d_n2_clc1.json 2> gpurun_out/bench_d_n2_clc1.err
CB_GEMM_CLC=0 timeout 500 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_d_n2_clc0.json 2> gpurun_out/bench_d_n2_clc0.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --config 13b-zero2 --steps 5 --warmup 3 > gpurun_out/bench_d_13b_n2.json 2> gpurun_out/bench_d_13b_n2.err
for f in gpurun_out/bench_d_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), d['peak_mem_gb'])
except Exception as e: print('$f ERR', e)
PY
done
tail -5 gpurun_out/bench_d_13b_n2.err
