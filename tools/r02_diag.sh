#!/bin/bash
# diagnostics build on the box: per-role cycle counters (RRIN_CONV_PROF=1) of every TMA-conv launch of one forward
tag=$1
mkdir -p gpurun_out
python -m rrin_b200.build --diag --force > gpurun_out/${tag}_diag_build.log 2>&1 || { tail -5 gpurun_out/${tag}_diag_build.log; exit 1; }
BATCH=${BATCH:-4} RRIN_CONV_PROF=1 python tools/profile_step.py 1088 1920 0 > gpurun_out/${tag}_role_cycles.txt 2>&1
grep -c "conv prof" gpurun_out/${tag}_role_cycles.txt
grep "conv prof cfg 11 \|conv prof cfg 25 \|conv prof cfg 13 \|conv prof cfg 10 " gpurun_out/${tag}_role_cycles.txt | head -8
