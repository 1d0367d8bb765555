"""Turns the raw ncu output of scripts/run_profiles.sh into the summaries kept under profiles/.
usage: python scripts/summarize_profiles.py <tag> [gpurun_out] [profiles]"""
import csv
import collections
import os
import re
import sys

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
dst = sys.argv[3] if len(sys.argv) > 3 else "profiles"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def read_raw(path):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def launch_summary(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    seq = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3) for r in rows
           if r["Metric Name"] == "gpu__time_duration.sum"]  # us
    return seq


def table(seq, title):
    agg = collections.OrderedDict()
    for k, us in seq:
        k = re.sub(r"\(.*", "", k)[:80]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = [f"### {title}: {len(seq)} launches, {tot/1e3:.2f} ms (sum of per-launch gpu__time_duration, cold cache, serialised)", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {us/1e3:.3f} | {100*us/tot:.1f}% | {us/n:.1f} |")
    return "\n".join(out) + "\n"


ll = os.path.join(src, f"{tag}_launches_profile_step.csv")
if os.path.exists(ll):
    seq = launch_summary(ll)
    # the prefill batch ends at the first decode-regime kernel after the argmax of the extend
    names = [k for k, _ in seq]
    first_argmax = next(i for i, k in enumerate(names) if "argmax" in k)
    pre, dec = seq[: first_argmax + 1], seq[first_argmax + 1:]
    # drop torch's batch-merge kernels between the phases from the decode table start
    per_step = [i for i, (k, _) in enumerate(dec) if "argmax" in k]
    one = dec[per_step[0] + 1: per_step[1] + 1] if len(per_step) > 1 else dec
    with open(os.path.join(dst, f"{tag}_launch_list_summary.md"), "w") as f:
        f.write(table(pre, "prefill batch, 8 x 1024 tokens, Llama-3-8B, full GPU (148 SMs), eager") + "\n")
        f.write(table(one, "one decode step, B = 64, context ~1090, eager (the engine replays this as one CUDA graph)"))
    print(open(os.path.join(dst, f"{tag}_launch_list_summary.md")).read())

with open(os.path.join(dst, f"{tag}_ncu_full_metrics.txt"), "w") as f:
    for fn in sorted(os.listdir(src)):
        if fn.startswith(tag + "_") and fn.endswith(".raw.csv"):
            m = read_raw(os.path.join(src, fn))
            f.write(f"== {fn[:-8]}  [{m.get('Kernel Name', ('?', ''))[0][:70]}]\n")
            for k in KEYS:
                if k in m:
                    f.write(f"   {k} = {m[k][0]} {m[k][1]}\n")
print(open(os.path.join(dst, f"{tag}_ncu_full_metrics.txt")).read())
