for d in 0 1 2 4 3 7; do RAE_TC_DEBUG=$d timeout 100 python bench.py --workload T --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('dbg=$d', 'ms/step %.4f'%j['ms_per_step'], 'fwd %.4f bwd %.4f'%(j['phase_ms']['decoder_forward'], j['phase_ms']['decoder_backward']))"; done
