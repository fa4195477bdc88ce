set -x

for v in 5 4 15 3 16; do
  MCS_K3_MINB=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VAR $v', j['value'], j['e2e']['value'], j['roofline']['stage_ms'])" | tee -a gpurun_out/k3_variants.log
done
ncu --set full --clock-control none --import-source on -k regex:describe -s 3 -c 1 -f -o gpurun_out/k3_poly python bench.py --steps 1 --warmup 3 --frames 32 --no-cpu-baseline > gpurun_out/ncu_k3.log 2>&1
