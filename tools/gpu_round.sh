# one GPU-box visit: parity tests, smoke, default bench
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_check.json").read().strip().splitlines()[-1])
print("BENCH", j["value"], j["e2e"]["value"], j["ms_per_step"], j["cpu_baseline"]["value"], j["clocks"])
PY
