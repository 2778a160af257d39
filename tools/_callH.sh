# 2-GPU call: symmetric all-reduce kernel tests + engine parity, N=2 bench with the in-switch all-reduce vs NCCL
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s --timeout=800 -p no:cacheprovider > gpurun_out/pytest_multi_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_gpu.log; grep -v "Generation\|trust_remote\|owner of" gpurun_out/pytest_multi_gpu.log | tail -25 | cut -c1-700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
CB_COLLECTIVE=multimem timeout 400 $TR --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_h_n2_multimem.json 2> gpurun_out/bench_h_n2_multimem.err
CB_COLLECTIVE=multimem CB_AR_CTAS=4 timeout 400 $TR --master-port 29532 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_h_n2_multimem_c4.json 2> gpurun_out/bench_h_n2_multimem_c4.err
CB_COLLECTIVE=p2p timeout 400 $TR --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_h_n2_p2p.json 2> gpurun_out/bench_h_n2_p2p.err
NCCL_PROTO=Simple timeout 400 $TR --master-port 29534 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_h_n2_nccl_simple.json 2> gpurun_out/bench_h_n2_nccl_simple.err
for f in gpurun_out/bench_h_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('$f ERR', e)
PY
done
tail -4 gpurun_out/bench_h_n2_multimem.err | cut -c1-400
