set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 180 > gpurun_out/t1_parity.log 2>&1; echo "parity rc=$?" 
tail -15 gpurun_out/t1_parity.log
timeout 600 python -m pytest tests/test_scale_fullsize.py tests/test_golden.py -x -q -m gpu --timeout 300 > gpurun_out/t1_full.log 2>&1; echo "full rc=$?"
tail -15 gpurun_out/t1_full.log
timeout 300 python bench.py --workload T --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/t1_bench_T.json 2> gpurun_out/t1_bench_T.err; echo "bench T rc=$?"
cat gpurun_out/t1_bench_T.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phase_ms'])"
timeout 300 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/t1_bench_cfg2.json 2> gpurun_out/t1_bench_cfg2.err; echo "bench cfg2 rc=$?"
cat gpurun_out/t1_bench_cfg2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phase_ms'])"
