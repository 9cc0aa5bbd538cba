"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches <launches.csv> <n_tail_launches>   -> per-kernel-class time shares of the last forward
  python tools/ncu_summary.py rep <file.ncu-rep> [...]                    -> key metrics of each profiled launch (ncu --set full)
"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
]


def short(name):
    m = re.search(r"(conv3x3_\w+_kernel<[^>]*>|\w+_kernel)", name)
    return m.group(1) if m else name[:60]


def launches(path, n_tail):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"], r["Block Size"]))
    tail = rows[-n_tail:]
    tot = sum(t for _, t, *_ in tail)
    print(f"# last {len(tail)} launches (= one Net.forward at 1088x1920) of {len(rows)} captured; ncu-serialised, cold-cache: compare SHARES")
    print(f"# total {tot:.1f} us")
    cls = {}
    for n, t, *_ in tail:
        c = cls.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += t
    print("kernel,launches,total_us,share_pct,avg_us")
    for n, (k, t) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        print(f'"{n}",{k},{t:.1f},{100 * t / tot:.1f},{t / k:.1f}')
    print("\n# per launch, stream order")
    print("idx,kernel,us,grid,block")
    for i, (n, t, g, b) in enumerate(tail):
        print(f'{i},"{n}",{t:.1f},"{g}","{b}"')


def rep(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"## {p.split('/')[-1]} :: {short(d['Kernel Name'])}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for k in KEYS:
                if k in d:
                    print(f"{k} = {d[k]} {units[hdr.index(k)]}")
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        rep(sys.argv[2:])
