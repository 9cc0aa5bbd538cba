"""profiles/traffic.json from `ncu --set full` captures: DRAM bytes (read + write) of one representative launch per
kernel class, next to that launch's algorithmic bytes.  bench.py copies the dominant class's figure into `roofline.traffic`.

  python tools/traffic_json.py <tag> <batch> [out.json]     (reads gpurun_out/<tag>_conv<i>.ncu-rep; default output
                                                             profiles/traffic.json)

Launch index i -> (class name as bench.py prints it, layer): fixed by the engine's schedule (tools/step_table.py -v)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 1088 * 1920
# i: (class, layer, algorithmic bytes per frame pair: bf16 activations in + out, + packed weights once)
LAUNCHES = {
    5: ("conv3x3_tma#16<KCS64,KB64,NT128,MSUB3>", "Flow.down_path.2.block.2 (128->128, level 2)", (128 + 128) * 2 * P / 16, 9 * 128 * 128 * 2),
    12: ("conv3x3_tma#16<KCS64,KB64,NT128,MSUB3> (cat)", "Flow.up_path.0.conv_block.block.0 (cat 256+256->256, level 3)", (512 + 256) * 2 * P / 64, 9 * 512 * 256 * 2),
    2: ("conv3x3_tma#21<KCS32,KB32,NT64,MSUB4>", "Flow.down_path.1.block.0 (pooled 32->64, level 1)", (32 + 64) * 2 * P / 4, 9 * 32 * 64 * 2),
    1: ("conv3x3_tma#35<KCS64,KB32,NT128,MSUB1>", "Flow.down_path.0.block.2 (32->32, level 0, + pooled output; two tile streams)", (32 + 32 + 8) * 2 * P, 96 * 1024),
    3: ("conv3x3_tma#36<KCS64,KB64,NT64,MSUB1>", "Flow.down_path.1.block.2 (64->64, level 1, + pooled output; two tile streams)", (64 + 64 + 16) * 2 * P / 4, 9 * 64 * 64 * 2),
    20: ("conv3x3_tma#22<KCS64,KB64,NT64,MSUB4>", "Flow.up_path.2.conv_block.block.0 (cat 64+64->64, level 1; CTA pairs)", (128 + 64) * 2 * P / 4, 9 * 128 * 64 * 2),
    23: ("conv3x3_tma#25<KCS64,KB32,NT128,MSUB1>", "Flow.up_path.3.conv_block.block.0 (cat 32+32->32, level 0; CTA pairs, resident half-blocks)", (64 + 32) * 2 * P, 192 * 1024),
    25: ("conv3x3_tma#37<KCS64,KB32,NT16,MSUB2> +glue", "Flow.last (32->4, level 0) + fused t-scale / refine_flow head packing", (32 * 2 + 16 + 32 + 24) * P, 16 * 1024),
    46: ("conv3x3_tma#44<KCS64,KB32,NT16,MSUB2> +glue", "refine_flow.last (32->4) + fused residue add, both backward warps, Mask head packing (model.py:44-50)", (32 * 2 + 16 + 24 + 32 + 32) * P, 16 * 1024),      # frames counted once; the staged windows re-read their 8-pixel halo (mostly L2 hits)
    67: ("conv3x3_tma#37<KCS64,KB32,NT16,MSUB2> +glue (blend)", "Mask.last (32->2) + fused sigmoid, blend, final head packing (model.py:52-55,61)", (32 * 2 + 32 + 24 + 16 + 32) * P, 16 * 1024),
    88: ("conv3x3_tma#37<KCS64,KB32,NT16,MSUB2> +glue (clamp)", "final.last (32->3) + fused residue add, clamp, NCHW store (model.py:62-63)", (32 * 2 + 16 + 12) * P, 16 * 1024),
    17: ("conv3x3_tma#17<KCS64,KB64,NT128,MSUB3> fold", "Flow.up_path.2.up.1 (weight-folded bilinear x2 + 128->64, level 2 -> level 1, scatter stores)", 128 * 2 * P / 16 + 64 * 2 * P / 4, 9 * 128 * 256 * 2),
    21: ("conv3x3_tma#18<KCS64,KB64,NT128,MSUB3> fold", "Flow.up_path.3.up.1 (weight-folded bilinear x2 + 64->32, level 1 -> level 0)", 64 * 2 * P / 4 + 32 * 2 * P, 9 * 64 * 128 * 2),
    11: ("conv3x3_tma#20<KCS64,KB64,NT128,MSUB2> up", "Flow.up_path.0.up.1 (bilinear x2 + 512->256, level 3)", (512 * 2 / 4 + 256 * 2) * P / 64, 9 * 512 * 256 * 2),
}


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[2]
    d = {}
    for k, uu, x in zip(h, u, v):
        if k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[uu]
            d[k] = float(x.replace(",", "")) * mul
    return d


def main():
    tag, batch = sys.argv[1], int(sys.argv[2])
    res = {"_note": f"dram__bytes_read.sum + dram__bytes_write.sum of ONE launch per class from `ncu --set full --clock-control none` "
                    f"captures ({tag}), 1088x1920, batch {batch} frame pairs (bench.py's default step); `algorithmic` = bf16 activations "
                    f"in + out of that launch + its packed weights once"}
    for i, (cls, layer, act, wts) in LAUNCHES.items():
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_conv{i}.ncu-rep")
        if not os.path.exists(rep):
            continue
        m = metrics(rep)
        tr = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        alg = act * batch + wts
        res[cls] = {"traffic": int(tr), "algorithmic": int(alg), "ratio": round(tr / alg, 3), "launch": layer,
                    "dram_read": int(m["dram__bytes_read.sum"]), "dram_write": int(m["dram__bytes_write.sum"])}
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
