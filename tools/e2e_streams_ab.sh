for s in 2 3 4 6 8; do python bench.py --no-cpu-baseline --e2e-streams $s --streams $s --repeats 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('streams',$s,'value',round(d['value']),d['ms_per_step_runs'],'e2e',round(d['e2e']['value']),d['e2e']['ms_per_step_runs'])"; done
