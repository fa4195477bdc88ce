# bench config 2 with the default library and every A/B library given (multicol_slam_b200/libmcs_b200_<name>.so): value, step time, stage times
for v in default "$@"; do
  if [ $v = default ]; then unset MCS_B200_LIB; else export MCS_B200_LIB=$PWD/multicol_slam_b200/libmcs_b200_$v.so; fi
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - $v <<'PY'
import json, sys
try:
    j = json.loads(open(f"gpurun_out/bench_var_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    m = j["roofline"]["other_stages"]["m2_match_stream_greedy"]
    print("VARIANT", sys.argv[1], round(j["value"], 2), round(j["e2e"]["value"], 2), round(j["ms_per_step"], 3), "lists", round(m["lists_ms"], 3), "replay", round(m["replay_ms"], 3),
          {k: round(v, 3) for k, v in j["roofline"]["stage_ms"].items()})
except Exception as e:
    print("bench failed", sys.argv[1], e)
PY
done
