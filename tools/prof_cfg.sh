touch screenpressor_b200/csrc/decode.cu
make -s -C screenpressor_b200/csrc EXTRA=-DSCPR_PROF 2>&1 | grep -v deprecated
timeout 300 python tools/stage_times.py cfg4_1440p_intra 2 2>&1 | grep "dec prof\|rep 1" | tail -3
timeout 300 python tools/stage_times.py cfg3_2160p_rgb32 60 2>&1 | grep "dec prof\|rep 1" | tail -3
timeout 300 python tools/stage_times.py cfg5_5120x1440 120 2>&1 | grep "dec prof\|rep 1" | tail -3
