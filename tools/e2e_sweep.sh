for t in 0 0; do FLACB200_MD5_THREADS=$t python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('threads', $t, round(d['e2e']['value']), d['e2e']['last_call_breakdown_ms'])"; done
