#!/bin/bash
# round-2 first GPU pass: parity tests, bench (graph on / off), per-launch table
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02a_tests.log
python bench.py --steps 20 --warmup 3 --per-launch > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
RRIN_GRAPH=0 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02a_bench_nograph.json 2> gpurun_out/r02a_bench_nograph.err; echo "bench nograph rc=$?"
tail -3 gpurun_out/r02a_bench.err gpurun_out/r02a_bench_nograph.err
python - <<'PY'
import json
for f in ("r02a_bench", "r02a_bench_nograph"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), "gap", d["launch_gap"]["gap_frac"], d["launch_gap"]["cuda_graph"], d.get("batch1"), d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
