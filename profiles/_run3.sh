cd $GRAFT_REPO_ROOT
export RAE_PARITY_LOG=$GRAFT_REPO_ROOT/gpurun_out/parity_errors.jsonl
rm -f $RAE_PARITY_LOG
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scale_fullsize.py tests/test_golden.py -q -m gpu --timeout 300 > gpurun_out/t3_parity.log 2>&1; echo "parity rc=$?"
tail -8 gpurun_out/t3_parity.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > gpurun_out/t3_bench_T.json 2> gpurun_out/t3_bench_T.err; echo "bench T rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t3_bench_T.json').read())
print("T", d['value'], d['ms_per_step'], d['rounds'])
print(d['phase_ms']); print(d['e2e']['value'])
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'], d['extra']['cfg2']['phase_ms'])
PY
timeout 300 python profiles/trace_tc.py T > gpurun_out/t3_trace_T.txt 2>&1; echo "trace rc=$?"
cat gpurun_out/t3_trace_T.txt | tail -120
