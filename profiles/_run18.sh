cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/t18_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/t18_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/t18_bench_T.json 2> gpurun_out/t18_bench_T.err; echo "bench T rc=$?"
tail -3 gpurun_out/t18_bench_T.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t18_bench_T.json').read().strip().splitlines()[-1])
print("T", d['value'], d['ms_per_step'], d['rounds'], d['e2e'], d.get('parity_ok'), d.get('max_rel_err'))
print(d['phase_ms'])
print(d['roofline'])
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'], d['extra']['cfg2']['phase_ms'])
PY
