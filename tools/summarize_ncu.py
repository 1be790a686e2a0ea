"""tools/summarize_ncu.py -- turn ncu outputs (run under gpurun) into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/X_launches.csv profiles/r01_launches_N16384.md
    python tools/summarize_ncu.py full     gpurun_out/X.ncu-rep      profiles/r01_gemm_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        val = val / 1e6 if unit == "ns" else val / 1e3 if unit == "us" else val
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("cflx::<unnamed>::", "").replace("unnamed>::", "")
        tot[name] += val
        cnt[name] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), source `{src}`\n\n")
        f.write(f"{sum(cnt.values())} launches, {s:.2f} ms summed device time (serialised, cold caches: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / s:.2f} % |\n")


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"ncu `--set full --clock-control none --import-source on`, source `{src}`\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"### {d.get('Kernel Name', '?')[:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            for k in hdr:
                if "issue_stalled" in k and "per_issue_active" in k and d[k] not in ("", "0"):
                    f.write(f"| stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {d[k]} | warps/issue |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
