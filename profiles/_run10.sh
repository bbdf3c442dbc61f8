cd $GRAFT_REPO_ROOT
timeout 300 python profiles/trace_tc.py T > gpurun_out/t10_trace_T.txt 2>&1; echo "trace rc=$?"
grep -A26 "CTA 74" gpurun_out/t10_trace_T.txt | head -28
