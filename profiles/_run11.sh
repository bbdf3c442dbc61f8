cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "matches_oracle" > gpurun_out/t11_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t11_parity.log
TRACE_KNOCK=1 timeout 300 python profiles/trace_tc.py T > gpurun_out/t11_trace_T.txt 2>&1; echo "trace rc=$?"
grep -A8 "knock-out" gpurun_out/t11_trace_T.txt | head -12; grep -A24 "CTA 74" gpurun_out/t11_trace_T.txt | head -26
