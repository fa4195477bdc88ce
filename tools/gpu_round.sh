# one GPU-box visit: parity tests, default bench, smoke
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("BENCH", j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["stage_ms"], j["cpu_baseline"]["value"], j["cpu_baseline"]["sample"])
PY
