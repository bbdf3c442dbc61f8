cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $OUT/t24_gpu.log 2>&1; echo "gpu rc=$?"; tail -5 $OUT/t24_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > $OUT/t24_bench_T.json 2> $OUT/t24_bench_T.err; echo "bench rc=$?"; tail -3 $OUT/t24_bench_T.err
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t24_bench_T.json').read().strip().splitlines()[-1])
print("T", d['value'], d['ms_per_step'], "cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'])
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['timeline_us']))
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['extra']['cfg2']['timeline_us']))
print(d['phase_ms']); print(d['extra']['cfg2']['phase_ms'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02c_launches_T.csv \
    python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0 > $OUT/r02c_ncu_T.log 2>&1
python profiles/summarize_launches.py $OUT/r02c_launches_T.csv 50 > $OUT/r02c_launches_T_summary.txt 2>&1
cat $OUT/r02c_launches_T_summary.txt
