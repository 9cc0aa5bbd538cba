#!/bin/bash
# ncu --set full captures of selected conv launches of one batch-4 1080p forward, summarised on the box
# usage: tools/r02_ncu.sh <tag> <conv launch index>...     (index among the conv3x3 launches of ONE forward, 0-based)
tag=$1; shift
mkdir -p gpurun_out
python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
nconv=89
for i in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s $((nconv + i)) -c 1 \
      -o gpurun_out/${tag}_conv${i} -f python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_ncu_conv${i}.log 2>&1
  ncu -i gpurun_out/${tag}_conv${i}.ncu-rep --page source --csv > gpurun_out/${tag}_conv${i}_source.csv 2>/dev/null
  python tools/ncu_top_stalls.py gpurun_out/${tag}_conv${i}_source.csv 45 > gpurun_out/${tag}_conv${i}_stalls.txt 2>&1
  rm -f gpurun_out/${tag}_conv${i}_source.csv
done
python tools/ncu_summary.py rep gpurun_out/${tag}_conv*.ncu-rep > gpurun_out/${tag}_ncu_full_summary.txt 2>&1
for i in "$@"; do
  ncu -i gpurun_out/${tag}_conv${i}.ncu-rep --page details > gpurun_out/${tag}_conv${i}_details.txt 2>/dev/null
done
ls -la gpurun_out/${tag}_conv*.ncu-rep
du -sh gpurun_out
[ "${KEEP_REPS:-0}" = "1" ] || rm -f gpurun_out/${tag}_conv*.ncu-rep
cat gpurun_out/${tag}_ncu_full_summary.txt | grep -E "^##|gpu__time|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes|dram_throughput" 
