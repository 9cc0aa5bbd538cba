#!/bin/bash
# round-end evidence: ncu launch list of one batch-4 1080p forward + `--set full` captures of representative launches,
# summarised on the box (the reports exceed gpurun's return limit).  usage: tools/r02_profile.sh <tag> <conv launch index>...
tag=$1; shift
mkdir -p gpurun_out
python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches_1080p_b4.csv \
    python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_ncu_list.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${tag}_launches_1080p_b4.csv 91 > gpurun_out/${tag}_launches_summary.csv 2>&1
nconv=89
for i in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s $((nconv + i)) -c 1 \
      -o gpurun_out/${tag}_conv${i} -f python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_ncu_conv${i}.log 2>&1
done
python tools/ncu_summary.py rep gpurun_out/${tag}_conv*.ncu-rep > gpurun_out/${tag}_ncu_full_summary.txt 2>&1
python tools/traffic_json.py ${tag} 4 gpurun_out/${tag}_traffic.json > /dev/null 2>&1
for i in 46 67; do
  [ -f gpurun_out/${tag}_conv${i}.ncu-rep ] && ncu -i gpurun_out/${tag}_conv${i}.ncu-rep --page details > gpurun_out/${tag}_conv${i}_details.txt 2>/dev/null
done
rm -f gpurun_out/${tag}_conv*.ncu-rep gpurun_out/${tag}_ncu_conv*.log
head -30 gpurun_out/${tag}_launches_summary.csv
cat gpurun_out/${tag}_traffic.json | head -80
