cd $GRAFT_REPO_ROOT
OUT=gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check > $OUT/t19_bench_T.json 2> $OUT/t19_bench_T.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t19_bench_T.json').read().strip().splitlines()[-1])
print("T", d['value'], d['ms_per_step'])
for r in d['timeline_us']: print("  ", r)
print("cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'])
for r in d['extra']['cfg2']['timeline_us']: print("  ", r)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02a_launches_T.csv \
    python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0 > $OUT/r02a_ncu_T.log 2>&1
python profiles/summarize_launches.py $OUT/r02a_launches_T.csv 50 > $OUT/r02a_launches_T_summary.txt 2>&1
cat $OUT/r02a_launches_T_summary.txt
