set -x
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_r02g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02g.log; tail -4 gpurun_out/pytest_r02g.log
timeout 200 python tools/probe1.py gemv > gpurun_out/probe_gemv.log 2>&1; tail -42 gpurun_out/probe_gemv.log
timeout 300 python tools/decode_bench.py --batch 1 --profile > gpurun_out/decode_b1_graph.json 2> gpurun_out/decode_b1_graph.err; cat gpurun_out/decode_b1_graph.json
timeout 300 python tools/decode_bench.py --batch 8 > gpurun_out/decode_b8_graph.json 2> gpurun_out/decode_b8_graph.err; cat gpurun_out/decode_b8_graph.json
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log; tail -4 gpurun_out/sanitizer_synccheck.log
timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_g_default.json 2> gpurun_out/bench_g_default.err; tail -c 600 gpurun_out/bench_g_default.json
