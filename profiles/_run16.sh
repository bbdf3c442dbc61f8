cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 600 -x > gpurun_out/t16_dist.log 2>&1; echo "dist rc=$?"; tail -15 gpurun_out/t16_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/t16_bench_T_n2.json 2> gpurun_out/t16_bench_T_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/t16_bench_T_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t16_bench_T_n2.json').read().strip().splitlines()[-1])
print("T n2", d['value'], d['ms_per_step'], d.get('parity_ok'), d.get('max_rel_err'), d.get('dist_phase_ms'), d.get('setup_s'), d.get('epoch_level'))
PY
