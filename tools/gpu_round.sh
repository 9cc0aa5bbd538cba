#!/bin/bash
# One GPU pass (run under gpurun from the repo root): parity tests, bench, ncu launch list and
# `--set full` captures of selected conv launches.  Usage: tools/gpu_round.sh <tag> [conv launch indices...]
# Conv launch index i counts conv3x3_umma launches of ONE forward (0-based, stream order).
tag=${1:-r01}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" 
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_ncu_list.log 2>&1
nconv=$(python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from rrin_b200 import engine
import torch
e = engine.Engine(torch.device("cuda", 0), int(os.environ.get("BATCH", "4")), 1088, 1920)
print(sum(1 for n, *_ in e.launch_table() if n.startswith("conv")))
PY
)
for i in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s $((nconv + i)) -c 1 \
      -o gpurun_out/${tag}_conv${i} -f python tools/profile_step.py 1088 1920 1 > gpurun_out/${tag}_ncu_conv${i}.log 2>&1
done
# gpurun brings back at most 64 MiB: summarise the captures here and keep only the text (KEEP_REPS=1 keeps the reports)
reps=$(ls gpurun_out/${tag}_conv*.ncu-rep 2>/dev/null)
if [ -n "$reps" ]; then
  python tools/ncu_summary.py rep $reps > gpurun_out/${tag}_ncu_full_summary.txt 2>&1
  python tools/traffic_json.py ${tag} ${BATCH:-4} gpurun_out/${tag}_traffic.json > /dev/null 2>&1
  python tools/ncu_summary.py launches gpurun_out/${tag}_launches.csv 91 > gpurun_out/${tag}_launches_summary.csv 2>&1
  [ "${KEEP_REPS:-0}" = "1" ] || rm -f gpurun_out/${tag}_conv*.ncu-rep
fi
ls -la gpurun_out | tail -20
