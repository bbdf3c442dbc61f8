cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_scale_fullsize.py -q -m gpu --timeout 300 -x > $OUT/t21_parity.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/t21_parity.log
for V in pdl nopdl; do
F=""; [ $V = nopdl ] && F="--no-pdl"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check $F > $OUT/t21_bench_T_$V.json 2> $OUT/t21_bench_T_$V.err; echo "bench $V rc=$?"
python - $V <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t21_bench_T_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],"T", d['value'], d['ms_per_step'], "cfg2", d['extra']['cfg2']['value'], d['extra']['cfg2']['ms_per_step'])
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['timeline_us']))
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['extra']['cfg2']['timeline_us']))
PY
done
