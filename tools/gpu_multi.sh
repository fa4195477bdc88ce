# multi-GPU visit: bench.py under torchrun at N ranks for the configs that shard over N (2: stream-sharded; 3: <= 4 cameras; 4: <= 8 cameras)
N=$1; shift
for CFG in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$CFG bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --config $CFG > gpurun_out/bench_n${N}_c$CFG.json 2> gpurun_out/bench_n${N}_c$CFG.err
  python - $N $CFG <<'PY'
import json, sys
n, c = sys.argv[1:3]
try:
    j = json.loads(open(f"gpurun_out/bench_n{n}_c{c}.json").read().strip().splitlines()[-1])
    print("BENCH N", n, "config", c, j["run"].get("matcher_stats"), round(j["value"], 3), round(j["e2e"]["value"], 3), round(j["ms_per_step"], 3), j["run"].get("allgather"))
except Exception as e:
    print("bench failed", n, c, e); print(open(f"gpurun_out/bench_n{n}_c{c}.err").read()[-1500:])
PY
done
