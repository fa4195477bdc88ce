# one GPU visit: smoke, all GPU tests, bench config 2, K1..K3 stage times of A/B libraries, chunk plans (knobs build)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || { echo "SMOKE FAILED OR HUNG"; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; tail -12 gpurun_out/gpu_tests.log | cut -c1-400
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
    print("BENCH c2 value", round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "ms", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j["roofline"]["stage_ms"].items()},
          {k: round(v, 3) for k, v in j["roofline"]["other_stages"]["m2_match_stream_greedy"].items() if k.endswith("_ms")})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_c2.err").read()[-1500:])
PY
echo "== stage probe default"; timeout 200 python tools/stage_probe.py 128 stats 2>&1 | grep -E "images:|tiers"
for v in $VARIANTS; do if [ -f multicol_slam_b200/libmcs_b200_$v.so ]; then echo "== stage probe $v"; MCS_B200_LIB=$PWD/multicol_slam_b200/libmcs_b200_$v.so timeout 200 python tools/stage_probe.py 128 2>&1 | grep -E "images:|tiers"; fi; done
if [ -n "$1" ]; then bash tools/trace_stream.sh "$@"; fi
