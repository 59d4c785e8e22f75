timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('stages_ms'))"
timeout 150 python tools/engine_breakdown.py 0 2>&1 | grep -E "^   \("
timeout 600 python -m pytest tests/test_gpu_query.py -x -q -k "sdf" 2>&1 | tail -2
