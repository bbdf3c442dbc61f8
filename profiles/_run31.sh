cd $GRAFT_REPO_ROOT
OUT=gpurun_out
nvidia-smi -L | wc -l
for WL in T cfg5; do
ST=20; [ $WL = cfg5 ] && ST=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps $ST --warmup 3 --workload $WL > $OUT/t31_bench_${WL}_n8.json 2> $OUT/t31_bench_${WL}_n8.err; echo "bench $WL n8 rc=$?"
tail -2 $OUT/t31_bench_${WL}_n8.err
python - $WL <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t31_bench_%s_n8.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],"n8", d['value'], d['ms_per_step'], d.get('parity_ok'), d.get('max_rel_err'), d.get('dist_phase_ms'), d.get('setup_s'))
print(" ".join("%s:%d:%.0f"%tuple(r) for r in (d.get('timeline_us') or [])))
PY
done
