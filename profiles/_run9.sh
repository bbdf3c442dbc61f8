cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "matches_oracle" > gpurun_out/t9_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t9_parity.log
TRACE_KNOCK=1 timeout 300 python profiles/trace_tc.py T > gpurun_out/t9_trace_T.txt 2>&1; echo "trace rc=$?"
grep -A8 "knock-out" gpurun_out/t9_trace_T.txt | head -12; grep -A16 "CTA 74" gpurun_out/t9_trace_T.txt | head -52
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check --no-extra > gpurun_out/t9_bench_T.json 2> gpurun_out/t9_bench_T.err; echo "bench T rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t9_bench_T.json').read())
print("T", d['value'], d['ms_per_step'], d['rounds'])
print(d['phase_ms'])
PY
