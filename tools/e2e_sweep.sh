for t in 4 8 12 16; do FLACB200_CHUNKS=$t python bench.py --no-cpu-baseline --steps 8 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks', $t, round(d['e2e']['value']), d['e2e']['last_call_breakdown_ms'])"; done
