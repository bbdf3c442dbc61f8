cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 600 > $OUT/t32_dist.log 2>&1; echo "dist rc=$?"; tail -4 $OUT/t32_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --workload T --peer-dense-max-mb 0 > $OUT/t32_bench_T_n2.json 2> $OUT/t32_bench_T_n2.err; echo "bench n2 rc=$?"
tail -2 $OUT/t32_bench_T_n2.err
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/t32_bench_T_n2.json').read().strip().splitlines()[-1])
print("T n2 nccl-aside", d['value'], d['ms_per_step'], d.get('parity_ok'), d.get('max_rel_err'), d.get('dist_phase_ms'))
print(" ".join("%s:%d:%.0f"%tuple(r) for r in d['timeline_us']))
PY
