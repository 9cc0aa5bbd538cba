#!/bin/bash
# GPU parity suite (no -x: report every failure) + a short bench line
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=12 -s > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_tests.log | head -40
grep -E "max-abs|psnr" gpurun_out/${tag}_tests.log | head -60
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), d["clocks"])
    for k, v in d["roofline"]["classes"].items(): print("  ", k, v)
except Exception as e:
    print("bench unreadable", e)
PY
tail -5 gpurun_out/${tag}_bench.err
