#!/bin/bash
# multi-GPU: bench.py under torchrun on N GPUs of one box (default line + strong-scaling clip mode)
# usage: tools/r02_scale.sh <tag> <N>
tag=$1; n=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 40 --warmup 5 \
    > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err; echo "weak n=$n rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --config clip --steps 5 --warmup 3 \
    > gpurun_out/${tag}_bench_clip_n${n}.json 2> gpurun_out/${tag}_bench_clip_n${n}.err; echo "clip n=$n rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_n${n}.json", "gpurun_out/${tag}_bench_clip_n${n}.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], d["scaling"], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in gpurun_out/${tag}_bench_n${n}.err gpurun_out/${tag}_bench_clip_n${n}.err; do tail -n 3 $f; done
