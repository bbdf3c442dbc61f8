cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 600 > $OUT/t30_dist.log 2>&1; echo "dist rc=$?"; tail -4 $OUT/t30_dist.log
for WL in T; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --workload $WL > $OUT/t30_bench_${WL}_n2.json 2> $OUT/t30_bench_${WL}_n2.err; echo "bench $WL n2 rc=$?"
tail -2 $OUT/t30_bench_${WL}_n2.err
python - $WL <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t30_bench_%s_n2.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],"n2", d['value'], d['ms_per_step'], d.get('parity_ok'), d.get('max_rel_err'), d.get('dist_phase_ms'))
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['timeline_us']))
PY
done
