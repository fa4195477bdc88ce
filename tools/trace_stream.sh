# e2e chunk plans of mcs_extract_match_stream_packed (needs the -DMCS_DEBUG_KNOBS build: multicol_slam_b200/libmcs_b200_knobs.so)
export MCS_B200_LIB=$PWD/multicol_slam_b200/libmcs_b200_knobs.so
for plan in "$@"; do
  MCS_STREAM_CHUNKS=$plan MCS_TRACE_STREAM=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/trace_bench.json 2> gpurun_out/trace_stream_$plan.log
  echo "PLAN $plan e2e $(python -c "import json; print(json.loads(open('gpurun_out/trace_bench.json').read().strip().splitlines()[-1])['e2e']['value'])")" | tee -a gpurun_out/trace_plans.log
  tail -8 gpurun_out/trace_stream_$plan.log | tail -4 | tee -a gpurun_out/trace_plans.log
done
