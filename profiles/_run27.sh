cd $GRAFT_REPO_ROOT
OUT=gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 600 > $OUT/t27_dist.log 2>&1; echo "dist rc=$?"; tail -8 $OUT/t27_dist.log
for WL in T cfg4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --workload $WL > $OUT/t27_bench_${WL}_n2.json 2> $OUT/t27_bench_${WL}_n2.err; echo "bench $WL n2 rc=$?"
tail -2 $OUT/t27_bench_${WL}_n2.err
python - $WL <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t27_bench_%s_n2.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],"n2", d['value'], d['ms_per_step'], d.get('parity_ok'), d.get('max_rel_err'), d.get('dist_phase_ms'), d.get('setup_s'), d.get('epoch_level'))
PY
done
