# 8-GPU call: config 3 with the in-switch all-reduce vs NCCL (Simple) on the same box; config 5 (34B ZeRO-3 generate --check)
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
CB_COLLECTIVE=multimem timeout 300 $TR --master-port 29711 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/bench_j_n8_multimem.json 2> gpurun_out/bench_j_n8_multimem.err
NCCL_PROTO=Simple timeout 300 $TR --master-port 29712 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/bench_j_n8_nccl_simple.json 2> gpurun_out/bench_j_n8_nccl_simple.err
timeout 400 $TR --master-port 29713 tools/zero3_generate.py --layers 60 --new-tokens 16 --check > gpurun_out/zero3_34b_n8_check.log 2>&1
grep -v "Generation\|trust_remote\|owner of" gpurun_out/zero3_34b_n8_check.log | tail -3 | cut -c1-700
for f in gpurun_out/bench_j_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']; print('$f', round(d['value'],3), round(d['per_gpu'],3), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], round(r['frac'],3))
except Exception as e: print('$f ERR', e)
PY
done
tail -3 gpurun_out/bench_j_n8_multimem.err | cut -c1-400
