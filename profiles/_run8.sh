cd $GRAFT_REPO_ROOT
./profiles/microbench/tmem_ld_under_mma > gpurun_out/t8_tmem_ld.txt 2>&1; cat gpurun_out/t8_tmem_ld.txt
timeout 300 python profiles/trace_tc.py T > gpurun_out/t8_trace_T.txt 2>&1; echo "trace rc=$?"
grep -A16 "CTA 74" gpurun_out/t8_trace_T.txt | head -40
