#!/usr/bin/env python
"""Summarise ncu output for profiles/: either a `--set full` report (.ncu-rep, read with `ncu -i ... --page raw --csv`)
or a launch list csv from `ncu --metrics gpu__time_duration.sum --csv --log-file ...`.

    python tools/ncu_summary.py rep  gpurun_out/x.ncu-rep  > profiles/r01_x_full.txt
    python tools/ncu_summary.py list gpurun_out/launches.csv > profiles/r01_x_launches.txt
    python tools/ncu_summary.py traffic gpurun_out/ctc.ncu-rep profiles/ctc_traffic.json   # what bench.py's roofline.traffic reads
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_global_ld.sum", "sm__sass_inst_executed_op_global_st.sum"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of {path} ({len(data)} launches); units in brackets")
    for n, r in enumerate(data):
        print(f"\n## launch {n}: {r[kn]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:90s} {r[i]:>16s} [{units[i]}]")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')]); wr = float(r[hdr.index('dram__bytes_write.sum')])
            u1, u2 = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * sc[u1] + wr * sc[u2]
            t = float(r[hdr.index('gpu__time_duration.sum')]); tu = units[hdr.index('gpu__time_duration.sum')]
            t *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[tu]
            print(f"{'dram traffic (read+write)':90s} {tot/1e6:16.2f} [MB]   = {tot/t/1e9:.0f} GB/s under ncu clocks")
        except Exception:
            pass


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r[mu]]
        d = agg.setdefault(r[kn], [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list {path}: {len(data)} launches, {tot:.1f} us total (cold-cache, serialised; compare SHARES)")
    print(f"{'total us':>12s} {'share':>7s} {'n':>6s} {'avg us':>10s}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.1f} {100 * t / tot:6.1f}% {n:6d} {t / n:10.2f}  {k[:140]}")


def traffic(path, out_path):
    """DRAM bytes (read + write) of the LAST launch of the CTC scan and gradient kernels in a --set full report, with the
    hash of the kernel sources they were built from: bench.py quotes the figure only while that hash matches."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tsc = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    last = {}
    for r in data:
        m = re.search(r"(ctc_scan\w*|ctc_grad\w*)", r[kn])
        if not m:
            continue
        name = m.group(1)
        rd, wr = float(r[ir].replace(",", "")) * sc[units[ir]], float(r[iw].replace(",", "")) * sc[units[iw]]
        if rd + wr < 1e6:            # the guarded log-domain twins exit at once: not part of the traffic
            continue
        last[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "time_us": float(r[it].replace(",", "")) * tsc[units[it]]}
    head = subprocess.run(["git", "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
    d = {"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, config 2 "
                 "(B=64 T=1000 V=801 fp32), one CTC forward+backward", "report": os.path.basename(path),
         "source_hash": bench.source_hash(), "git_head": head, "kernels": last,
         "traffic_bytes": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in last.values())}
    json.dump(d, open(out_path, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
