"""Summarise ncu artefacts brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py <tag> <launches.csv> <prof.ncu-rep> [bench.json]

Writes profiles/<tag>_launches.md (per-kernel device time and share of the
profiled command), profiles/<tag>_top_kernel.md (key `ncu --set full` metrics and
the warp-stall breakdown of the dominant kernel)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.reader(io.StringIO("".join(lines)))
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr):
            continue
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(unit, 1.0)
        rows.append((r[ix["Kernel Name"]].split("(")[0], v * scale))
    return rows


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = defaultdict(int)
    n = 0
    top = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            k = int(r[ix["# Samples"] if "# Samples" in ix else ix["Warp Stall Sampling (All Samples)"]])
        except ValueError:
            continue
        n += k
        for s in names:
            try:
                tot[s] += int(r[ix[s]])
            except ValueError:
                pass
        top.append((k, r[ix["Source"]].strip()[:90]))
    top.sort(reverse=True)
    return n, dict(tot), top[:12]


def main():
    tag, lpath, rep = sys.argv[1:4]
    bench = sys.argv[4] if len(sys.argv) > 4 else None
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    rows = launches(lpath)
    per = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        per[k][0] += 1
        per[k][1] += us
    total = sum(v[1] for v in per.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as fh:
        fh.write(f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        fh.write(f"source: `{os.path.basename(lpath)}`; {len(rows)} launches, {total / 1e3:.3f} ms of device time "
                 "(cold-cache, serialised: compare shares, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, us) in sorted(per.items(), key=lambda x: -x[1][1]):
            fh.write(f"| `{k}` | {c} | {us:.1f} | {100 * us / total:.1f}% |\n")
    m = raw_metrics(rep)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    n, tot, top = stalls(rep)
    with open(os.path.join(ROOT, "profiles", f"{tag}_top_kernel.md"), "w") as fh:
        fh.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the dominant kernel\n\n")
        fh.write(f"source: `{os.path.basename(rep)}` (one launch = one whole tCG solve)\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in keys:
            if k in m:
                fh.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
        try:
            rd = float(m["dram__bytes_read.sum"][0].replace(",", ""))
            wr = float(m["dram__bytes_write.sum"][0].replace(",", ""))
            ur, uw = m["dram__bytes_read.sum"][1], m["dram__bytes_write.sum"][1]
            sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            traffic = rd * sc[ur] + wr * sc[uw]
            fh.write(f"\nDRAM traffic per launch: {traffic / 1e9:.3f} GB")
            if bench:
                b = json.loads(open(bench).read().strip().splitlines()[-1])
                its = b["config"]["cg_iterations_per_step"]
                alg = b["roofline"]["algorithmic_bytes_per_cg_step"] * its
                fh.write(f" = {traffic / its / 1e6:.1f} MB per CG iteration ({its:.0f} iterations); algorithmic bytes per "
                         f"launch {alg / 1e9:.3f} GB -> traffic / algorithmic = {traffic / alg:.3f}\n")
                fh.write(f"\nbench line of the same build: value {b['value']:.1f} {b['unit']}, roofline.achieved "
                         f"{b['roofline']['achieved']:.1f} GB/s = {b['roofline']['frac']:.3f} of measured peak, "
                         f"e2e {b['e2e']['value']:.1f}, kernel share of step {b['roofline']['kernel_share_of_step']:.3f}\n")
        except Exception as e:  # noqa
            fh.write(f"\n(traffic summary failed: {e})\n")
        fh.write(f"\n## warp stall sampling ({n} samples)\n\n| reason | share |\n|---|---:|\n")
        for s, v in sorted(tot.items(), key=lambda x: -x[1])[:9]:
            fh.write(f"| {s} | {100 * v / max(n, 1):.1f}% |\n")
        fh.write("\n## hottest SASS lines\n\n| samples | instruction |\n|---:|---|\n")
        for k, src in top:
            fh.write(f"| {k} | `{src}` |\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_top_kernel.md")).read())
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.md")).read())


if __name__ == "__main__":
    main()
