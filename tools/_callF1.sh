# 1-GPU call: tests, decode latency (graph / eager loop), timeline at N=1, sanitizer, ncu launch list of one step
set -x
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/pytest_r02f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02f.log; tail -4 gpurun_out/pytest_r02f.log
timeout 300 python tools/decode_bench.py --batch 1 > gpurun_out/decode_b1_graph.json 2> gpurun_out/decode_b1_graph.err; cat gpurun_out/decode_b1_graph.json
timeout 300 python tools/decode_bench.py --batch 8 > gpurun_out/decode_b8_graph.json 2> gpurun_out/decode_b8_graph.err; cat gpurun_out/decode_b8_graph.json
timeout 300 python tools/trace_step.py --gpus 1 > gpurun_out/trace_n1.log 2>&1; tail -2 gpurun_out/trace_n1.log | cut -c1-600
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log; tail -6 gpurun_out/sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log; tail -4 gpurun_out/sanitizer_synccheck.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_b4_r02.csv python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/launches_step_b4_r02.csv
timeout 300 python bench.py --steps 8 --warmup 3 --micro-batch 8 --recompute 1 --no-cpu-baseline > gpurun_out/bench_f_b8_recompute.json 2> gpurun_out/bench_f_b8_recompute.err; tail -c 400 gpurun_out/bench_f_b8_recompute.json
