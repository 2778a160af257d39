set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s --timeout=800 -p no:cacheprovider > gpurun_out/pytest_multi_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_gpu.log; grep -v "Generation\|trust_remote\|owner of" gpurun_out/pytest_multi_gpu.log | tail -8 | cut -c1-500
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
CB_COLLECTIVE=multimem timeout 400 $TR --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_i_n2_multimem.json 2> gpurun_out/bench_i_n2_multimem.err
CB_COLLECTIVE=multimem CB_AR_CTAS=48 timeout 400 $TR --master-port 29542 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_i_n2_multimem_c48.json 2> gpurun_out/bench_i_n2_multimem_c48.err
NCCL_PROTO=Simple timeout 400 $TR --master-port 29543 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_i_n2_nccl_simple.json 2> gpurun_out/bench_i_n2_nccl_simple.err
for f in gpurun_out/bench_i_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']; print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(r['frac'],3), round(r['share_of_step']*d['ms_per_step'],1))
except Exception as e: print('$f ERR', e)
PY
done
tail -3 gpurun_out/bench_i_n2_multimem.err | cut -c1-300
