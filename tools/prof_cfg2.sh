touch screenpressor_b200/csrc/decode.cu
make -s -C screenpressor_b200/csrc EXTRA=-DSCPR_PROF 2>&1 | grep -v deprecated
timeout 300 python tools/stage_times.py cfg2_1080p_rgb32 600 2>&1 | grep "dec prof\|rec prof\|rep 1" | tail -5
