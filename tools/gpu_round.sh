# one GPU-box visit: parity tests, K3 wave sweep, sanitizer on the bag-of-words tests
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
for w in 8 4 16 32; do
  MCS_K3_WAVES=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WAVES $w', j['value'], j['e2e']['value'], j['roofline']['stage_ms'])" | tee -a gpurun_out/k3_waves.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bow_gpu.py -x -q -k "not big_node" > gpurun_out/sanitize_bow.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitize_bow.log; tail -4 gpurun_out/sanitize_bow.log
