cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_scale_fullsize.py tests/test_driver.py -q -m gpu --timeout 300 > $OUT/t26_gpu.log 2>&1; echo "gpu rc=$?"; tail -4 $OUT/t26_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > $OUT/t26_bench_T.json 2> $OUT/t26_bench_T.err; echo "bench rc=$?"; tail -3 $OUT/t26_bench_T.err
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t26_bench_T.json').read().strip().splitlines()[-1])
print("T", d['value'], d['ms_per_step'], "e2e", d['e2e']['value'], "cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'])
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['timeline_us']))
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['extra']['cfg2']['timeline_us']))
print(d['phase_ms']); print(d['extra']['cfg2']['phase_ms'])
PY
