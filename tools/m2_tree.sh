timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bow_profile.py 2>&1 | tail -1
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', j['value'], j['e2e']['value'], j['roofline']['stage_ms'], j['single_frame_latency_ms']['value'])"
