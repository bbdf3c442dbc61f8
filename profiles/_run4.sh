cd $GRAFT_REPO_ROOT
./profiles/microbench/mma_power > gpurun_out/t4_mma_power.txt 2>&1; echo "mma_power rc=$?"; cat gpurun_out/t4_mma_power.txt
timeout 300 python profiles/trace_tc.py T > gpurun_out/t4_trace_T.txt 2>&1; echo "trace rc=$?"
grep -E "^==|kernel wall" gpurun_out/t4_trace_T.txt
export RAE_PARITY_LOG=$GRAFT_REPO_ROOT/gpurun_out/parity_errors.jsonl
rm -f $RAE_PARITY_LOG
timeout 600 python -m pytest tests/test_scale_fullsize.py -q -m gpu --timeout 300 > gpurun_out/t4_full.log 2>&1; echo "full rc=$?"; tail -4 gpurun_out/t4_full.log
python - <<'PY'
import json, collections
worst = collections.defaultdict(float)
for l in open('gpurun_out/parity_errors.jsonl'):
    r = json.loads(l); worst[(r['test'], r['what'])] = max(worst[(r['test'], r['what'])], r['err'])
for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:12]: print("%.3g %s" % (v, k))
PY
