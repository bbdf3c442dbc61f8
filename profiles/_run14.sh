cd $GRAFT_REPO_ROOT
export RAE_PARITY_LOG=$GRAFT_REPO_ROOT/gpurun_out/parity_errors.jsonl
rm -f $RAE_PARITY_LOG
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu --timeout 300 -x > gpurun_out/t14_parity.log 2>&1; echo "parity rc=$?"; tail -12 gpurun_out/t14_parity.log
timeout 600 python -m pytest tests/test_scale_fullsize.py -q -m gpu --timeout 300 > gpurun_out/t14_full.log 2>&1; echo "full rc=$?"; tail -6 gpurun_out/t14_full.log
python - <<'PY'
import json, collections
worst = collections.defaultdict(float)
for l in open('gpurun_out/parity_errors.jsonl'):
    r = json.loads(l); worst[(r['test'], r['what'])] = max(worst[(r['test'], r['what'])], r['err'])
for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:8]: print("%.3g %s" % (v, k))
PY
TRACE_KNOCK=1 timeout 300 python profiles/trace_tc.py T > gpurun_out/t14_trace_T.txt 2>&1; echo "trace rc=$?"
grep -A8 "knock-out" gpurun_out/t14_trace_T.txt | head -12; grep -A16 "== dC" gpurun_out/t14_trace_T.txt | head -34
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > gpurun_out/t14_bench_T.json 2> gpurun_out/t14_bench_T.err; echo "bench T rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t14_bench_T.json').read())
print("T", d['value'], d['ms_per_step'], d['rounds'])
print(d['phase_ms'])
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'], d['extra']['cfg2']['phase_ms'])
PY
