cd $GRAFT_REPO_ROOT
TRACE_KNOCK=1 timeout 300 python profiles/trace_tc.py T > gpurun_out/t6_trace_T.txt 2>&1; echo "trace rc=$?"
grep -B2 -A9 "knock-out" gpurun_out/t6_trace_T.txt | head -30
