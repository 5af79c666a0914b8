"""Turn the ncu CSVs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py r1      # reads gpurun_out/r1_*.csv, writes profiles/r1_*.{md,csv,json}
"""
import collections
import csv
import gzip
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__cycles_elapsed\.max|sm__cycles_elapsed\.avg\.per_second|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
    r"launch__shared_mem_per_block_dynamic|launch__shared_mem_per_block_static|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__inst_executed\.sum|sm__inst_executed_pipe_(xu|fma|alu|lsu|tensor_subpipe_hmma|tmem|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active|sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"lts__t_sectors_srcunit_tex_op_read\.sum|lts__t_bytes\.sum)$")
TRIAGE = re.compile(r"TriageCompute\.(sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed|dram__throughput.*|lts__throughput.*)$")


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def launches(tag):
    src = os.path.join(GO, f"{tag}_final_launches.csv")
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out = os.path.join(PR, f"{tag}_launches_by_kernel.csv")
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,avg_us,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"\"{k}\",{n},{t / 1e6:.4f},{t / n / 1e3:.2f},{100 * t / tot:.2f}\n")
    with open(src, "rb") as fi, gzip.open(os.path.join(PR, f"{tag}_launches_raw.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return agg, tot


def raw(tag, name):
    src = os.path.join(GO, f"{tag}_{name}_raw.csv")
    rows = list(csv.reader(open(src)))
    h, units = rows[0], rows[1]
    kn = h.index("Kernel Name")
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[kn])}
        for i, c in enumerate(h):
            if KEEP.match(c) or TRIAGE.search(c):
                d[c] = (r[i], units[i])
        out.append(d)
    return out


def main(tag):
    os.makedirs(PR, exist_ok=True)
    agg, tot = launches(tag)
    md = [f"# ncu summaries, round {tag[1:]} (B200, `--clock-control none`)\n",
          "Source: `gpurun_out/` captures of `python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-e2e` (launch list: "
          "setup + 5 evaluation steps + the roofline probe; the same command exited 0 without ncu first) and "
          "`ncu --set full` of single launches; regenerate with `python tools/summarize_profiles.py " + tag + "`.\n",
          "Per-launch times of the launch list are cold-cache and serialised: read SHARES, not absolutes.\n",
          "## Launch list — top kernels (setup + evaluation steps + roofline probe)\n",
          "| kernel | launches | total ms | avg µs | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
        md.append(f"| `{k}` | {n} | {t / 1e6:.3f} | {t / n / 1e3:.1f} | {100 * t / tot:.1f} % |")
    traffic = {}
    for name in ("rank_mma", "recheck", "eval_small", "train", "k1", "k1t"):
        if not os.path.exists(os.path.join(GO, f"{tag}_{name}_raw.csv")):
            continue
        md.append(f"\n## `ncu --set full` — {name}\n")
        note = {"rank_mma": "Final build of the round (the dominant kernel; `bench.py` workload, one launch).",
                "recheck": "Final build of the round (whole-row staging, 4 lanes per pair).",
                "eval_small": "Earlier build of the round: the small kernels around the contraction (K1 at 500 queries, target scores, "
                              "filter pass); the target / filter kernels have since moved to the re-check kernel's staging.",
                "train": "Earlier build of the round (before the multi-table scatter / Adagrad launches and the lane-group K3 mapping): "
                         "kept for the per-kernel picture of the training chain.",
                "k1": "K1 lane-group kernel, forward and adjoint, at 2^20 queries (`tools/k1_profile_case.py`).",
                "k1t": "K1 thread-per-query variant at 2^20 queries, ungrouped input (`tools/k1_profile_case.py`)."}.get(name)
        if note:
            md.append(note + "\n")
        for d in raw(tag, name):
            if name == "eval_small" and "recheck_kernel" in d["kernel"]:
                continue                              # superseded by the dedicated capture of the current re-check kernel
            md.append(f"### `{d['kernel']}`\n")
            md.append("| metric | value | unit |")
            md.append("|---|---:|---|")
            for c, v in d.items():
                if c != "kernel":
                    md.append(f"| {c} | {v[0]} | {v[1]} |")
            md.append("")
            if "rank_mma_kernel" in d["kernel"]:
                def num(key):
                    v, u = d[key]
                    x = float(v.replace(",", ""))
                    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
                traffic["rank_mma_kernel_big4m_dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                traffic["rank_mma_kernel_big4m_duration_ms_under_ncu"] = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        shutil.copyfile(os.path.join(GO, f"{tag}_{name}_raw.csv"), os.path.join(PR, f"{tag}_{name}_raw.csv"))
    for name in ("rank_mma", "recheck", "k1", "k1t"):
        src = os.path.join(GO, f"{tag}_{name}_source.csv.gz")
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(PR, f"{tag}_{name}_source.csv.gz"))
    open(os.path.join(PR, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(PR, "traffic.json"), "w"), indent=1)
    print("\n".join(md[:30]))
    print(traffic)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
