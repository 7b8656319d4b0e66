#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py <tag> [<workload>]

Reads   gpurun_out/launches_<tag>.csv            (ncu --metrics gpu__time_duration.sum launch list)
        gpurun_out/<kernel>_<tag>.ncu-rep        (ncu --set full captures, one per hot kernel)
Writes  profiles/<tag>_launches.md               per-kernel share of one substep (cold-cache, serialised)
        profiles/<tag>_<kernel>.md               key metrics + top stall reasons + hottest SASS lines
        profiles/traffic.json                    dram bytes per launch, consumed by bench.py's roofline.traffic
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]


def kernel_key(name):
    """'void gfs::k_g2p_brick<(int)0, (bool)0>(gfs::Grid, ...)' -> 'k_g2p_brick<0>' (base name + first template integer):
    the same normalisation bench.py applies to its own kernel names when it looks the traffic up."""
    import re
    m = re.search(r"(\w+)\s*<\s*(?:\(\w+\))?\s*(\d+)", name)
    if m:
        return "%s<%s>" % (m.group(1), m.group(2))
    return re.sub(r"\(.*", "", name).replace("void ", "").split("::")[-1].strip()


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(tag):
    path = os.path.join(SRC, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        val_us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        per.setdefault(name, []).append(val_us)
    tot = sum(sum(v) for v in per.values())
    with open(os.path.join(OUT, "%s_launches.md" % tag), "w") as f:
        f.write("# ncu launch list, tag %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over a short bench.py run "
                "(all launches of the process: scene generation by torch included).  Times are cold-cache and serialised: "
                "compare SHARES, not absolutes.\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n" % tag)
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (name[:90], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
        # the kernels of ONE trilinear substep, by mean launch time: the shares to compare with bench.py's "kernels" block
        step = ["k_p2g_tile2<1>", "k_g2p_tri<0, 0, 0>", "k_finalize_assemble", "k_p2g_tile<2>", "k_g2p_brick<0, 0>", "k_build_index", "k_assemble",
                "k_p2g_finalize", "k_classify", "DeviceScanKernel", "k_resolve_collisions", "k_g2p_slow"]
        if any("k_p2g_tile2<1>" in n for n in per):       # round 2 default path: the round-1 kernels run only in variant sweeps
            step = [k for k in step if k not in ("k_p2g_tile<2>", "k_g2p_brick<0, 0>", "k_assemble", "k_p2g_finalize")]
        rows = []
        for key in step:
            hit = [(n, v) for n, v in per.items() if key in n and "at_cuda_detail" not in n.split("DeviceScanKernel")[0][:0]]
            if hit:
                n, v = hit[0]
                rows.append((key, sorted(v)[len(v) // 2]))          # median: the first launch after an upload sorts physically
        tot_step = sum(m for _, m in rows) or 1.0
        f.write("\n## one trilinear substep (median launch time of each of its kernels)\n\n| kernel | median us | share of the substep |\n|---|---:|---:|\n")
        for key, m in rows:
            f.write("| `%s` | %.1f | %.1f %% |\n" % (key, m, 100 * m / tot_step))
        f.write("| total | %.1f | |\n" % tot_step)
    print("wrote launches for", tag, "kernels:", len(per))


def kernel_report(tag, kernel, workload, traffic):
    rep = os.path.join(SRC, "%s_%s.ncu-rep" % (kernel, tag))
    if not os.path.exists(rep):
        return
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    name = m.get("Kernel Name", ("?", ""))[0]
    stalls = sorted(((float(v.replace(",", "")), h) for h, (v, u) in m.items()
                     if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    sh, sd = src[1], src[2:]
    ia, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
    tot = max(1, sum(int(r[isamp]) for r in sd))
    ops = Counter()
    for r in sd:
        t = r[ia].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[isamp])
    with open(os.path.join(OUT, "%s_%s.md" % (tag, kernel)), "w") as f:
        f.write("# ncu --set full: `%s`, tag %s, workload %s\n\n" % (name.split("(")[0], tag, workload))
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in m:
                f.write("| %s | %s | %s |\n" % (k, m[k][0], m[k][1]))
        f.write("\n## top stall reasons (warps stalled per issue-active cycle)\n\n")
        for v, h in stalls[:8]:
            f.write("- %.2f  %s\n" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        f.write("\n## stall samples by SASS opcode (%d samples)\n\n" % tot)
        for op, n in ops.most_common(12):
            f.write("- %s: %.1f %%\n" % (op, 100.0 * n / tot))
        f.write("\n## hottest SASS lines\n\n```\n")
        for r in sorted(sd, key=lambda r: -int(r[isamp]))[:15]:
            f.write("%6s samples  %9s exec  %s\n" % (r[isamp], r[iex], r[ia].strip()[:100]))
        f.write("```\n")

    def num(k):
        v, u = m[k]
        x = float(v.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic.setdefault(workload, {})[kernel_key(name)] = \
        num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    # which pipe the kernel actually sits on (bench.py attaches this to its roofline object next to the HBM fraction)
    bpath = os.path.join(OUT, "bounds.json")
    bounds = json.load(open(bpath)) if os.path.exists(bpath) else {}

    def pct(k):
        return round(float(m[k][0].replace(",", "")), 2) if k in m else None
    bounds.setdefault(workload, {})[kernel_key(name)] = {
        "tag": tag,
        "lsu_wavefronts_pct_of_peak": pct("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "shared_atomic_wavefronts": pct("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"),
        "source": "profiles/%s_%s.md (ncu --set full)" % (tag, kernel)}
    json.dump(bounds, open(bpath, "w"), indent=1, sort_keys=True)
    print("wrote", kernel, tag)


def main():
    tag = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "dambreak128"
    launches(tag)
    tpath = os.path.join(OUT, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for f in sorted(os.listdir(SRC)):
        if f.endswith("_%s.ncu-rep" % tag):
            kernel_report(tag, f[: -len("_%s.ncu-rep" % tag)], workload, traffic)
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
