cd $GRAFT_REPO_ROOT
OUT=gpurun_out
ncu --set full --clock-control none -k "regex:k_radix_sort|k_tc_transpose|k_tc_combine|k_score|k_tc_dq|k_rows_chunk|k_dense_finalize|k_entity_long2|k_tc_bwd_finish|k_encoder_forward_v4" --launch-skip 260 -c 13 -o $OUT/r02c_full_sel -f \
    python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0 > $OUT/r02c_ncu_full_sel.log 2>&1
tail -3 $OUT/r02c_ncu_full_sel.log
ncu -i $OUT/r02c_full_sel.ncu-rep --page raw --csv > $OUT/r02c_full_sel_raw.csv 2>/dev/null
ncu -i $OUT/r02c_full_sel.ncu-rep --page details > $OUT/r02c_full_sel_details.txt 2>/dev/null
ls -la $OUT/r02c_full_sel*
S=$(stat -c %s $OUT/r02c_full_sel.ncu-rep); [ $S -gt 40000000 ] && rm -f $OUT/r02c_full_sel.ncu-rep
