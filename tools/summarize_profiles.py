"""Turn ncu captures (gpurun_out/*.ncu-rep, launch-list CSV) into the small text summaries kept
under profiles/.  Runs on the CPU box (ncu -i needs no GPU).
usage: summarize_profiles.py <tag> <launches.csv> <rep> [<rep> ...]"""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def summarize_rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        name = d["Kernel Name"][0]
        lines = [f"kernel: {name}"]
        for k in KEYS:
            if k in d and d[k][0] not in ("", "n/a"):
                lines.append(f"  {k:82s} {d[k][0]:>14s} {d[k][1]}")
        t_us = float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(d["gpu__time_duration.sum"][1], 1)
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        lines.append(f"  derived: duration {t_us:.2f} us (under ncu: cold cache, serialised); DRAM traffic {1e-6 * (rd + wr):.1f} MB "
                     f"= {1e-3 * (rd + wr) / t_us:.0f} GB/s" + (f" = {100 * 1e-3 * (rd + wr) / t_us / PEAKS['hbm_gbs']:.1f} % of the measured {PEAKS['hbm_gbs']} GB/s" if PEAKS else ""))
        out.append("\n".join(lines))
    return out


def summarize_launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    seq = []
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[ui]]
        seq.append((r[ki].split("(")[0].replace("void ", ""), v))
    starts = [i for i, (n, _) in enumerate(seq) if "im2col" in n]
    step = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v in step)
    lines = [f"one forward = {len(step)} launches, sum of device durations {tot:.1f} us "
             f"(ncu: cold cache, serialised -- compare SHARES with bench.py, not absolutes)"]
    for n, (c, v) in agg.items():
        lines.append(f"  {n:62s} x{c:3d}  total {v:8.1f} us  {100 * v / tot:5.1f} %   avg {v / c:7.1f} us")
    return "\n".join(lines)


if __name__ == "__main__":
    tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.txt"), "w") as f:
        f.write(summarize_launches(launches) + "\n")
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt"), "w") as f:
        for rep in reps:
            f.write(f"==== {os.path.basename(rep)}\n" + "\n\n".join(summarize_rep(rep)) + "\n\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.txt")).read())
