for lib in librt_b200.so librt_b200_ef.so librt_b200_st1n.so; do
  RT_B200_LIBNAME=$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio --clock-control none -k regex:k_wf_traverse -s 3 -c 3 --csv python tools/prof_wf.py 1 K3 1 2 2>&1 | grep -E "k_wf_traverse" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | paste - - - - - | sed "s/^/$lib /"
done
