# one GPU visit: stage times of the K3 / K1 A/B builds, then an ncu capture of the default build's K3 and K1 (one launch each)
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || { echo "SMOKE FAILED OR HUNG"; exit 1; }
echo "== default"; timeout 200 python tools/stage_probe.py 128 stats 2>&1 | grep -E "images:|tiers"
for v in "$@"; do echo "== $v"; MCS_B200_LIB=$PWD/multicol_slam_b200/libmcs_b200_$v.so timeout 200 python tools/stage_probe.py 128 2>&1 | grep -E "images:"; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
for K in 4 8; do
MCS_BENCH_K=$K timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k$K.json 2> gpurun_out/bench_k$K.err
python - $K <<'PY'
import json, sys
try:
    j = json.loads(open(f"gpurun_out/bench_k{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("BENCH K", sys.argv[1], round(j["value"], 2), round(j["e2e"]["value"], 2), round(j["ms_per_step"], 3), {k: (round(v, 3) if v else v) for k, v in j["roofline"]["stage_ms"].items()}, j["run"].get("greedy_replay_redo_images"))
except Exception as e:
    print("bench failed", e)
PY
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:describe_kernel -c 1 -o gpurun_out/r2_k3c -f python tools/stage_probe.py 32 > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pyr_fast_kernel -c 1 -o gpurun_out/r2_k1b -f python tools/stage_probe.py 32 > gpurun_out/ncu_k1.log 2>&1; tail -2 gpurun_out/ncu_k1.log
