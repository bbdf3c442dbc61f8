cd $GRAFT_REPO_ROOT
export RAE_PARITY_LOG=$GRAFT_REPO_ROOT/gpurun_out/parity_errors.jsonl
rm -f $RAE_PARITY_LOG
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t2_bench_T.json 2> gpurun_out/t2_bench_T.err; echo "bench T rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t2_bench_T.json').read())
print("T", d['value'], d['ms_per_step'], d['rounds'], d.get('parity_ok'), d.get('max_rel_err'), d.get('tf32_peak_tflops_measured'))
print(d['phase_ms']); print(d['e2e']); print(d['roofline']); print(d['setup_s'])
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'], d['extra']['cfg2']['phase_ms'])
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scale_fullsize.py tests/test_golden.py -x -q -m gpu --timeout 300 > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/t2_parity.log
python - <<'PY'
import json, collections
worst = collections.defaultdict(float)
for l in open('gpurun_out/parity_errors.jsonl'):
    r = json.loads(l); worst[r['test']] = max(worst[r['test']], r['err'])
for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:12]: print("%.3g %s" % (v, k))
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
