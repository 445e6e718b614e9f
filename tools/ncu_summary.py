"""Turns ncu output into the tracked summaries under profiles/.

  python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep  profiles/r01_ncu_full.md
  python tools/ncu_summary.py list  gpurun_out/launches.csv  profiles/r01_launches.md
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict, defaultdict
from pathlib import Path

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
           "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]


def full(rep, out):
    """rep: an .ncu-rep file, or the CSV written on the GPU box by `ncu --set full --csv --page raw`"""
    if str(rep).endswith(".csv"):
        raw = Path(rep).read_text()
        raw = raw[raw.index('"ID"'):]
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if "stall" in h and "ratio" in h and "not_issued" not in h]
    seen = OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        seen.setdefault(name.split("(")[0], r)
    lines = [f"# ncu --set full summary ({Path(rep).name})", "",
             "One launch per kernel (cold cache, serialised; compare shares, not absolutes).", ""]
    js = {}
    for name, r in seen.items():
        lines.append(f"## `{name}`")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        rec = {}
        for m in METRICS:
            if m in idx:
                lines.append(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
                rec[m] = r[idx[m]]
        vals = sorted([(float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace(
            "smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", ""))
                       for h in st if r[idx[h]] not in ("", "n/a")], reverse=True)
        lines.append(f"| top stall reasons (warps per issue) | {', '.join(f'{h} {v:.2f}' for v, h in vals[:5])} | |")
        lines.append("")
        js[name] = rec
    Path(out).write_text("\n".join(lines))
    Path(out).with_suffix(".json").write_text(json.dumps(js, indent=1))


def launch_list(csv_path, out):
    text = Path(csv_path).read_text()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = r["Kernel Name"].split("(")[0]
        tot[k] += us
        cnt[k] += 1
    total = sum(tot.values())
    lines = [f"# ncu launch list ({Path(csv_path).name}): device time per kernel", "",
             f"{sum(cnt.values())} launches, {total / 1e3:.2f} ms of kernel time (cold-cache, serialised).", "",
             "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k in sorted(tot, key=tot.get, reverse=True):
        lines.append(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / total:.1f}% | {tot[k] / cnt[k]:.2f} |")
    Path(out).write_text("\n".join(lines) + "\n")
    Path(out).with_suffix(".json").write_text(json.dumps(
        {k: {"launches": cnt[k], "total_us": tot[k], "share": tot[k] / total} for k in tot}, indent=1))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
