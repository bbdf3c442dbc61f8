cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "matches_oracle or regulariser" > gpurun_out/t15_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t15_parity.log
timeout 300 python profiles/trace_tc.py T > gpurun_out/t15_trace_T.txt 2>&1; echo "trace rc=$?"
grep -A18 "== dC" gpurun_out/t15_trace_T.txt | head -40
CMD="python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0"
$CMD > gpurun_out/t15_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t15_launches_T.csv $CMD > gpurun_out/t15_ncu.log 2>&1
echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/t15_launches_T.csv 45 | grep -v "cutlass\|at::"
