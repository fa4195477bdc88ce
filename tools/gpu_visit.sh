# one GPU visit: smoke, all GPU tests, the default bench line of config 2 (with the reference-code cpu baseline)
mkdir -p gpurun_out
VARIANTS=${VARIANTS:-"p2u4 p1u8 both"}
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || { echo "SMOKE FAILED OR HUNG"; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; tail -6 gpurun_out/gpu_tests.log | cut -c1-400
timeout 400 python bench.py > gpurun_out/r3_bench_c2.json 2> gpurun_out/r3_bench_c2.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r3_bench_c2.json").read().strip().splitlines()[-1])
    print("BENCH c2 value", round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "ms", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j["roofline"]["stage_ms"].items()},
          {k: round(v, 3) for k, v in j["roofline"]["other_stages"]["m2_match_stream_greedy"].items() if k.endswith("_ms")}, j["cpu_baseline"]["value"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r3_bench_c2.err").read()[-1500:])
PY
echo "== stage probe default"; timeout 100 python tools/stage_probe.py 128 2>&1 | grep -E "images:"
for v in $VARIANTS; do if [ -f multicol_slam_b200/libmcs_b200_$v.so ]; then echo "== stage probe $v"; MCS_B200_LIB=$PWD/multicol_slam_b200/libmcs_b200_$v.so timeout 100 python tools/stage_probe.py 128 2>&1 | grep -E "images:"; fi; done
