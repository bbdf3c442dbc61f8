cd $GRAFT_REPO_ROOT
OUT=gpurun_out
TRACE_KNOCK=1 timeout 600 python profiles/trace_tc.py T > $OUT/t33_trace_T.txt 2>&1; echo "trace rc=$?"
cat $OUT/t33_trace_T.txt | tail -90
