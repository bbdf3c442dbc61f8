cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_scale_fullsize.py -q -m gpu --timeout 300 -x > $OUT/t20_parity.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/t20_parity.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > $OUT/t20_bench_T.json 2> $OUT/t20_bench_T.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t20_bench_T.json').read().strip().splitlines()[-1])
print("T", d['value'], d['ms_per_step'])
for r in d['timeline_us']: print("  ", r)
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'])
for r in d['extra']['cfg2']['timeline_us']: print("  ", r)
PY
# one steady-state step's kernels, full sections (launch-skip: setup ~ 330 launches incl. bind-time sorts, then 3 warm-up steps)
ncu --set full --clock-control none --import-source on --launch-skip 420 -c 45 -o $OUT/r02a_full_T -f \
    python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0 > $OUT/r02a_ncu_full_T.log 2>&1
ncu -i $OUT/r02a_full_T.ncu-rep --page raw --csv > $OUT/r02a_ncu_full_T_raw.csv 2>/dev/null
ls -la $OUT | tail
