"""Turn ncu exports into the tracked summaries under profiles/ (runs on the CPU box).

  python scripts/profile_summaries.py launches  gpurun_out/r01_launches.csv  profiles/r01_launches_summary.md  "<command>"
  python scripts/profile_summaries.py kernels   gpurun_out/r01_raw.csv       profiles/r01_cfg3_kernels_summary.md  "<command>"
      (raw.csv = `ncu -i report.ncu-rep --page raw --csv`; also rewrites profiles/traffic.json)
"""
import csv, json, os, re, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    name = re.sub(r"\(.*$", "", name).strip()
    return name


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr = next(r for r in rows if "Kernel Name" in r)
    iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    n = 0
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
            continue
        unit = r[hdr.index("Metric Unit")]
        v = float(r[iV].replace(",", ""))
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        a = agg.setdefault(short(r[iK]), [0, 0.0])
        a[0] += 1; a[1] += ms; n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# r02 ncu launch list of `%s` (%d launches; --metrics gpu__time_duration.sum, --clock-control none)\n\n" % (cmd, n))
        f.write("Per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.4f | %.1f%% |\n" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print(open(dst).read())


def kernels(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr, units = rows[0], rows[1]
    iK = hdr.index("Kernel Name")
    cols = []
    for r in rows[2:]:
        cols.append((short(r[iK]), {h: (r[j], units[j]) for j, h in enumerate(hdr)}))
    with open(dst, "w") as f:
        f.write("# r02 `ncu --set full --clock-control none` capture, %s\n\n" % cmd)
        f.write("| metric | " + " | ".join(c[0] for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for k in KEYS:
            if k not in hdr:
                continue
            unit = cols[0][1][k][1]
            vals = []
            for _, d in cols:
                v = d[k][0].replace(",", "")
                try:
                    x = float(v)
                    if unit in ("ns", "nsecond"):
                        x /= 1e6
                    vals.append("%.6g" % x)
                except ValueError:
                    vals.append(v)
            u = "ms" if unit in ("ns", "nsecond") else unit
            f.write("| %s [%s] | %s |\n" % (k, u, " | ".join(vals)))
    traffic = {}
    for name, d in cols:
        def val(k):
            v, u = d[k]
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        key = "prepare_sens" if "prepare" in name else ("gp_sweep" if "gp_sweep" in name else ("qp" if "qp_" in name else name))
        traffic[key] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    if "gp_sweep" in traffic or "prepare_sens" in traffic:          # the preparation phase = both of its kernels
        traffic["prepare"] = traffic.get("gp_sweep", 0.0) + traffic.get("prepare_sens", 0.0)
    traffic["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, %s (profiles/%s)" % (cmd, os.path.basename(dst))
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(open(dst).read())
    print(traffic)


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    if len(sys.argv) > 5:            # round tag for the headers (default r01)
        txt = open(sys.argv[3]).read().replace("# r01 ", "# %s " % sys.argv[5], 1)
        open(sys.argv[3], "w").write(txt)
