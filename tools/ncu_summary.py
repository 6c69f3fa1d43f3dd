"""Summarise ncu outputs into small text files for profiles/:
   python tools/ncu_summary.py launches <launches.csv> <out.md> [N]  (from --metrics gpu__time_duration.sum --csv; N = keep only
                                                                      the last N launches of our kernels; -1 = one step)
   python tools/ncu_summary.py full <report.ncu-rep> <out.md>        (from --set full)"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]


OURS = ("igemm_kernel", "attention_kernel", "gn_", "layernorm", "conv_in_kernel", "conv_out_kernel", "cfg_sched_kernel",
        "transpose_tokens_kernel", "linear_small", "timestep_sinusoid", "f32_to_bf16")


def launches(path, out, last=0):
    """last > 0: keep only the last `last` launches of this library's kernels (one step out of a longer capture)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if last == -1:      # exactly one denoise step: the launches between two consecutive cfg_sched kernels (the step's last kernel)
        ends = [i for i, r in enumerate(rows) if "cfg_sched_kernel" in r["Kernel Name"]]
        rows = [r for r in rows[ends[-2] + 1: ends[-1] + 1] if any(k in r["Kernel Name"] for k in OURS)]
    elif last:
        rows = [r for r in rows if any(k in r["Kernel Name"] for k in OURS)][-last:]
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:80]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n{sum(v[0] for v in agg.values())} launches, {tot:.0f} us summed "
                "(per-launch times are cold-cache and serialised: compare SHARES)\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n")
            for c in cols:
                f.write(f"- {hdr[c]} = {r[c]} {units[c]}\n")
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3])
