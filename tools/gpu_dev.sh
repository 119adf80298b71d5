#!/bin/bash
# Developer loop on the GPU box: parity check with the shipped build, then a -DSCPR_PROF rebuild of the decoder
# (in-kernel cycle counters) and the stage timings of one clip.  Usage: tools/gpu_dev.sh [tag] [dev_check filters...]
# Every step runs under its own timeout so that a hung kernel cannot hold the box.
tag=${1:-x}; shift
timeout 600 python tools/dev_check.py "$@" > gpurun_out/dev_$tag.log 2>&1; echo "dev_check rc=$?"; tail -3 gpurun_out/dev_$tag.log | cut -c1-200
grep -B1 -A3 "MISMATCH" gpurun_out/dev_$tag.log | head -20
timeout 300 python tools/stage_times.py cfg2_1080p_rgb32 600 2>&1 | grep "rep 1\|_batch" | tail -3
timeout 300 python tools/stage_times.py cfg3_2160p_rgb32 60 2>&1 | grep "rep 1\|decode_batch" | tail -2
timeout 300 python tools/stage_times.py cfg5_5120x1440 120 2>&1 | grep "rep 1\|decode_batch" | tail -2
timeout 300 python tools/stage_times.py cfg4_1440p_intra 12 2>&1 | grep "rep 1\|decode_batch" | tail -2
touch screenpressor_b200/csrc/decode.cu
make -s -C screenpressor_b200/csrc EXTRA=-DSCPR_PROF 2>&1 | grep -v deprecated
timeout 300 python tools/stage_times.py cfg2_1080p_rgb32 600 2>&1 | grep "dec prof" | tail -2
