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
    """launch list with time (and, when captured, DRAM bytes): per-kernel launches, summed device time, share, and the
    achieved DRAM GB/s of each kernel against the measured HBM peak (MEASURED_PEAKS.json: 6569 GB/s)."""
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt, byt = collections.Counter(), collections.Counter(), collections.Counter()
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row in csv.DictReader(lines):
        val = float(row["Metric Value"].replace(",", "")) * scale.get(row["Metric Unit"], 1.0)
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = name.replace("cflx::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<.*", "", name)
        if row["Metric Name"].startswith("gpu__time_duration"):
            tot[name] += val
            cnt[name] += 1
        elif row["Metric Name"].startswith("dram__bytes"):
            byt[name] += val
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none`), source `{src}`\n\n")
        f.write(f"{sum(cnt.values())} launches, {s:.2f} ms summed device time (serialised, cold caches: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share | DRAM GB moved | achieved GB/s | of 6569 GB/s |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, v in tot.most_common():
            gb = byt.get(k, 0.0) / 1e9
            gbs = gb / (v * 1e-3) if v > 0 and gb > 0 else 0.0
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / s:.2f} % | {gb:.3f} | {gbs:.0f} | {gbs / 6569.3:.2f} |\n" if gb > 0 else
                    f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / s:.2f} % | - | - | - |\n")


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
        # tcgen05 kernels: int8 tensor sub-pipe, tensor-memory traffic, FP64 pipe of the epilogue, L2 / L1 throughput
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


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


def traffic(src, dst, kernel, M, N, K):
    """append {kernel, M, N, K, dram_bytes} of the FIRST captured launch of `src` to the JSON bench.py reads"""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, d = rows[0], rows[1], dict(zip(rows[0], rows[2]))
    def val(k):
        x, u = float(d[k].replace(",", "")), units[hdr.index(k)]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out = json.load(open(dst)) if os.path.exists(dst) else {"what": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the trailing-update kernel (ncu --set full)", "launches": []}
    out["launches"] = [e for e in out["launches"] if (e["kernel"], e["M"], e["N"], e["K"]) != (kernel, M, N, K)]
    out["launches"].append({"kernel": kernel, "M": M, "N": N, "K": K, "dram_bytes": tot, "duration_ms": val("gpu__time_duration.sum") / 1e6
                            if units[hdr.index("gpu__time_duration.sum")] == "ns" else val("gpu__time_duration.sum"), "source": src})
    json.dump(out, open(dst, "w"), indent=1)
    print(out["launches"][-1])


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]))
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
