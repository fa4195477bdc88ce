# one GPU-box visit: parity (all GPU tests) + bench with the default build, then extractor parity + bench for every variant library given
# guard: a kernel that hangs (e.g. a TMA / mbarrier mistake) must not burn the visit -- one tiny extraction under a short timeout first
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || { echo "SMOKE FAILED OR HUNG: aborting this visit"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
timeout 120 python -m pytest tests/test_extract_gpu.py -m gpu -q -k tier -s 2>&1 | grep "K3 tiers"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
for VAR in "$@"; do
  N=$(basename $VAR .so)
  MCS_B200_LIB=$PWD/$VAR timeout 600 python -m pytest tests/test_extract_gpu.py tests/test_ref_pin_gpu.py -m gpu -x -q > gpurun_out/gpu_tests_$N.log 2>&1; tail -1 gpurun_out/gpu_tests_$N.log
  MCS_B200_LIB=$PWD/$VAR timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
done
python - "$@" <<'PY'
import json, sys, os
for n in ["base"] + [os.path.basename(v)[:-3] for v in sys.argv[1:]]:
    try:
        j = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"], 2), round(j["e2e"]["value"], 2), round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j["roofline"]["stage_ms"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
