cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_scale_fullsize.py -q -m gpu --timeout 300 -x > gpurun_out/t17_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t17_parity.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > gpurun_out/t17_bench_T.json 2> gpurun_out/t17_bench_T.err; echo "bench T rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t17_bench_T.json').read())
print("T", d['value'], d['ms_per_step'], d['rounds'])
print(d['phase_ms'])
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'], d['extra']['cfg2']['phase_ms'])
PY
