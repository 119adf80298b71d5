#!/bin/bash
# One ncu capture (selected sections) of every hot kernel of a short encode + decode; details page as CSV.
# Usage: tools/ncu_kernels.sh [tag] [frames]
tag=${1:-r01}; n=${2:-24}
rep=/tmp/kern_$tag.ncu-rep
timeout 900 ncu --clock-control none --section SpeedOfLight --section Occupancy --section WarpStateStats --section SchedulerStats \
  --section LaunchStats --section MemoryWorkloadAnalysis --section SourceCounters --section ComputeWorkloadAnalysis \
  -k regex:'k_rans_encode|k_replay_color|k_replay_fixed|k_mv_resolve|k_mv_search|k_mv_prematch|k_i_classify|k_i_emit|k_p_runs|k_p_emit|k_dec_fill|k_dec_chain|k_sort_place|k_frame_scan_tma|k_frame_scan32|k_mv_cands|k_i_emit_hdr|k_assemble' \
  -c 40 -f -o ${rep%.ncu-rep} python tools/ncu_decode.py cfg2_1080p_rgb32 $n > gpurun_out/ncu_kern_$tag.log 2>&1
tail -2 gpurun_out/ncu_kern_$tag.log
ncu -i $rep --page details --csv > gpurun_out/ncu_kernels_$tag.csv 2>/dev/null
ls -la $rep gpurun_out/ncu_kernels_$tag.csv
