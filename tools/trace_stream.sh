timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
for plan in 8,40,40,40 32,32,32,32 16,38,37,37 64,64 128 16,16,16,16,16,16,16,16 4,12,28,28,28,28; do
  MCS_STREAM_CHUNKS=$plan MCS_TRACE_STREAM=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/trace_bench.json 2> gpurun_out/trace_stream_$plan.log
  echo "PLAN $plan e2e $(python -c "import json; print(json.loads(open('gpurun_out/trace_bench.json').read().strip().splitlines()[-1])['e2e']['value'])")" | tee -a gpurun_out/trace_plans.log
  tail -8 gpurun_out/trace_stream_$plan.log | tail -4 | tee -a gpurun_out/trace_plans.log
done
