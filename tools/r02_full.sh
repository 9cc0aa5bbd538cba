#!/bin/bash
# full GPU pass: parity suite, every named bench config, reference arm, ncu launch list
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 -s > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -2
grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_tests.log | head -20
python bench.py --steps 100 --warmup 5 --per-launch > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench 1080p rc=$?"
for c in 720p_b8 1080p_t7 4k clip; do
  python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${tag}_bench_${c}.json 2> gpurun_out/${tag}_bench_${c}.err; echo "bench $c rc=$?"
done
python bench.py --precision fp16 --steps 40 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_fp16.json 2> gpurun_out/${tag}_bench_fp16.err; echo "bench fp16 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d.get("metric"), "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 3),
              "clk", (d.get("clocks") or {}).get("sm_mhz"), "roof", round(d.get("roofline", {}).get("frac", 0), 3),
              "whole", round(d.get("roofline", {}).get("whole_step", {}).get("frac_of_sustained_peak", 0), 3), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in gpurun_out/${tag}_bench_*.err; do tail -n 2 $f; done | head -40
