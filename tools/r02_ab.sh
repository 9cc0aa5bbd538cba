#!/bin/bash
# kernel tests for selected patterns, then A/B bench runs under different RRIN_* settings
# usage: tools/r02_ab.sh <tag> "<pytest -k expr>" "ENV1=.. ENV2=.." "ENV.." ...
tag=$1; kexpr=$2; shift 2
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$kexpr" > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
  grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -2
  grep -E "^(FAILED|ERROR)|Error|error:" gpurun_out/${tag}_tests.log | head -20
fi
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu --per-launch > gpurun_out/${tag}_bench$i.json 2> gpurun_out/${tag}_bench$i.err; rc=$?
  python - "$envs" gpurun_out/${tag}_bench$i.json $rc <<'PY'
import json, sys
envs, f, rc = sys.argv[1:4]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    c = d["roofline"]["classes"]
    print(f"[{envs}] value {d['value']:.2f} e2e {d['e2e']['value']:.2f} b1 {d.get('batch1', {}).get('value', 0):.1f} ms/step {d['ms_per_step']:.3f} clk {d['clocks'].get('sm_mhz')}")
    for k, v in c.items(): print("     ", k, v)
except Exception as e:
    print(f"[{envs}] rc={rc} unreadable: {e}")
PY
  tail -3 gpurun_out/${tag}_bench$i.err
done
