# one GPU-box visit: baseline parity + bench, then the same with a K3 variant build (MCS_B200_LIB)
VAR=${1:-multicol_slam_b200/libmcs_b200_k3h.so}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
MCS_B200_LIB=$PWD/$VAR timeout 600 python -m pytest tests/test_extract_gpu.py -m gpu -x -q > gpurun_out/gpu_tests_var.log 2>&1; tail -2 gpurun_out/gpu_tests_var.log
MCS_B200_LIB=$PWD/$VAR timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
python - <<'PY'
import json
for n in ("base", "var"):
    try:
        j = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["stage_ms"])
    except Exception as e:
        print(n, "failed", e)
PY
