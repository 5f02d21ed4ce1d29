"""Turns the ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [launches.csv] [prof_a.ncu-rep prof_b.ncu-rep ...]
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic"]


def launches(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r"<.*", "", row["Kernel Name"]).split("(")[0].replace("void ", "")
        t = float(row["Metric Value"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
        n += 1
    out.write("## Launch list (ncu --metrics gpu__time_duration.sum --clock-control none), %d launches, %.1f ms of kernel time\n\n" % (n, tot / 1e6))
    out.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `kernels`, not absolutes.\n\n")
    out.write("| kernel | launches | total ms | share | avg µs |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        out.write("| `%s` | %d | %.3f | %.2f %% | %.1f |\n" % (k[:80], v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))
    out.write("\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.write("## `%s` (ncu --set full --clock-control none --import-source on)\n\n" % os.path.basename(path))
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        grid, block = r[hdr.index("Grid Size")], r[hdr.index("Block Size")]
        if (name, grid) in seen:
            continue
        seen.add((name, grid))
        out.write("### `%s` grid %s block %s\n\n| metric | value |\n|---|---|\n" % (name, grid, block))
        for k in KEYS:
            if k in hdr:
                out.write("| %s | %s %s |\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        # warp-state: the largest stall reasons (cycles a warp waits per issued instruction) and the busiest pipes
        def top(pattern, n):
            vals = []
            for i, h_ in enumerate(hdr):
                if re.search(pattern, h_):
                    try:
                        vals.append((float(r[i]), h_, units[i]))
                    except ValueError:
                        pass
            return sorted(vals, reverse=True)[:n]
        for v_, h_, u_ in top(r"issue_stalled_.*_per_warp_active|average_warps_issue_stalled_.*ratio|average_warp_latency_issue_stalled", 6):
            out.write("| stall: %s | %.3f %s |\n" % (h_, v_, u_))
        for v_, h_, u_ in top(r"inst_executed_pipe_.*pct_of_peak|pipe_.*cycles_active.avg.pct_of_peak_sustained_active", 6):
            out.write("| pipe: %s | %.2f %s |\n" % (h_, v_, u_))
        for v_, h_, u_ in top(r"^l1tex__data_pipe_lsu_wavefronts.sum$|^l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$|^l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$|^smsp__inst_executed_pipe_fma.sum$|^sm__inst_executed_pipe_fma.sum$|^smsp__inst_executed_pipe_fmaheavy.sum$|^smsp__inst_executed_pipe_fmalite.sum$", 8):
            out.write("| count: %s | %.0f %s |\n" % (h_, v_, u_))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

        def in_bytes(key):
            if key not in hdr:
                return 0.0
            return float(r[hdr.index(key)]) * scale.get(units[hdr.index(key)], 1.0)

        tr = in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")
        dur = float(r[hdr.index("gpu__time_duration.sum")]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[hdr.index("gpu__time_duration.sum")], 1e-9)
        out.write("| **traffic (read+write)** | %.4f GB (= %.0f GB/s over the kernel's duration) |\n\n" % (tr / 1e9, tr / 1e9 / max(dur, 1e-12)))


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", tag + ".md"), "w") as out:
        out.write("# ncu summary %s\n\nCommands: `scripts/ncu_round2.sh` (1 × B200; operator kernels via `scripts/prof_operator.py`, launch list and history "
                  "kernels via `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-blocks`).\n\n" % tag)
        for a in sys.argv[2:]:
            if a.endswith(".csv"):
                launches(a, out)
            else:
                full(a, out)
    print("wrote profiles/%s.md" % tag)


if __name__ == "__main__":
    main()
