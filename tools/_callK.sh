# final 1-GPU call of round 2 (8 GPU-minutes left): new / previously unvalidated cases first, then the records still
# missing (graph + GEMV decode latency, the release-grid train step), then the rest of the suite
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
K="sep or bilinear or tower_combine or two_query or gemv or generate or greedy"
timeout 150 python -u -m pytest tests -m gpu -k "$K" -v --timeout=100 -p no:cacheprovider --tb=short > gpurun_out/pytest_k1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k1.log
grep -E "PASSED|FAILED|ERROR" gpurun_out/pytest_k1.log | cut -c1-160 | tail -40; tail -3 gpurun_out/pytest_k1.log
timeout 110 python tools/decode_bench.py --batch 1,8 --profile > gpurun_out/decode_graph_gemv.jsonl 2> gpurun_out/decode_graph_gemv.err; echo "decode rc=$?"; cut -c1-600 gpurun_out/decode_graph_gemv.jsonl; tail -3 gpurun_out/decode_graph_gemv.err | cut -c1-300
timeout 120 python bench.py --config 8b-release --steps 5 --warmup 3 > gpurun_out/bench_k_release.json 2> gpurun_out/bench_k_release.err; echo "release rc=$?"; cut -c1-400 gpurun_out/bench_k_release.json; tail -3 gpurun_out/bench_k_release.err | cut -c1-300
timeout 200 python -u -m pytest tests -m gpu -k "not ($K)" -v --timeout=100 -p no:cacheprovider --tb=short > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k2.log
grep -E "FAILED|ERROR" gpurun_out/pytest_k2.log | cut -c1-160 | tail -20; tail -3 gpurun_out/pytest_k2.log
timeout 60 python tools/probe1.py gemv > gpurun_out/probe_gemv.log 2>&1; tail -12 gpurun_out/probe_gemv.log
