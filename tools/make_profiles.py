"""Turns the raw files a GPU run left in gpurun_out/ into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""      # only the gpurun_out files whose name contains this (e.g. "r02")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else dict(hbm_gbs=6650.0, bf16_tflops_sustained=1400.0)
HBM, TF = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"]


def per_op(src, dst, title):
    d = json.load(open(src))
    ops = d["ops"]
    tot = sum(o["ms"] for o in ops)
    lines = [f"# {title}", "",
             f"CUDA-event time per launch (`yx_engine_profile`, mean of 3, each op timed back to back on the stream), "
             f"batch {d['batch']}, {d['size']}x{d['size']}.  Peaks: {HBM:.0f} GB/s HBM, {TF:.0f} TFLOP/s (measured, sustained).",
             f"Total {tot:.2f} ms over {len(ops)} launches.  `bound` = roofline side of the op (AI vs ridge {TF * 1e3 / HBM:.0f} FLOP/B); "
             "`frac` = achieved / peak on that side.", "",
             "| # | op | ms | % | GFLOP | MB | AI | TFLOP/s | GB/s | bound | frac | launch shape |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    ridge = TF * 1e3 / HBM
    agg = collections.defaultdict(float)
    for i, o in enumerate(ops):
        ms = max(o["ms"], 1e-6)
        tf, gb = o["flops"] / ms / 1e9, o["bytes"] / ms / 1e6
        ai = o["flops"] / max(o["bytes"], 1)
        bound = "tensor" if ai >= ridge else "hbm"
        frac = tf / TF if bound == "tensor" else gb / HBM
        agg[bound + "_ms"] += o["ms"]; agg[bound + "_w"] += (o["flops"] if bound == "tensor" else o["bytes"])
        lines.append(f"| {i} | {o['name']} | {o['ms']:.3f} | {100 * o['ms'] / tot:.1f} | {o['flops'] / 1e9:.1f} | {o['bytes'] / 1e6:.1f} | "
                     f"{ai:.0f} | {tf:.0f} | {gb:.0f} | {bound} | {frac:.2f} | {o.get('shape', '')} |")
    lines += ["", f"Tensor-bound ops: {agg['tensor_ms']:.2f} ms, aggregate {agg['tensor_w'] / max(agg['tensor_ms'], 1e-9) / 1e9:.0f} TFLOP/s "
              f"({agg['tensor_w'] / max(agg['tensor_ms'], 1e-9) / 1e9 / TF:.2f} of peak).",
              f"HBM-bound ops: {agg['hbm_ms']:.2f} ms, aggregate {agg['hbm_w'] / max(agg['hbm_ms'], 1e-9) / 1e6:.0f} GB/s "
              f"({agg['hbm_w'] / max(agg['hbm_ms'], 1e-9) / 1e6 / HBM:.2f} of peak)."]
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


def launches(src, dst):
    """ncu launch list (one row per launch and metric).  Writes the share table and, when the list carries
    dram__bytes_*, profiles/<tag>_traffic.json = measured DRAM bytes per launch of each kernel family (bench.py
    reports the conv family's figure as roofline.traffic)."""
    by_id = collections.OrderedDict()
    for r in csv.reader(open(src)):
        if len(r) < 11 or r[0] == "ID":
            continue
        v, u, metric = float(r[-1].replace(",", "")), r[-2], r[-3]
        k = r[4].split("(")[0].replace("void ", "")
        e = by_id.setdefault(r[0], dict(kernel=k))
        if metric.startswith("gpu__time_duration"):
            e["us"] = v / 1000 if u in ("nsecond", "ns") else v if u in ("usecond", "us") else v * 1000
        elif metric.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            e["dram"] = e.get("dram", 0.0) + v * scale
    tot = collections.OrderedDict()
    for e in by_id.values():
        t = tot.setdefault(e["kernel"], [0, 0.0, 0.0])
        t[0] += 1; t[1] += e.get("us", 0.0); t[2] += e.get("dram", 0.0)
    s = sum(v[1] for v in tot.values())
    has_dram = any(v[2] > 0 for v in tot.values())
    lines = ["# ncu launch list of bench steps (gpu__time_duration.sum" + (", dram__bytes_read/write.sum" if has_dram else "") +
             ", --clock-control none)", "",
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
             f"{len(by_id)} launches, {s / 1000:.2f} ms summed.", "",
             "| kernel | launches | sum us | share | DRAM MB / launch |", "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / s:.1f} % | {v[2] / v[0] / 1e6:.1f} |")
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for k, v in tot.items():
        f = "conv_gemm_kernel" if "conv_gemm_kernel" in k else k.split("<")[0]
        fam[f][0] += v[0]; fam[f][1] += v[1]; fam[f][2] += v[2]
    lines += ["", "| kernel family | launches | share | DRAM MB / launch |", "|---|---|---|---|"]
    for f, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{f}` | {v[0]} | {100 * v[1] / s:.1f} % | {v[2] / v[0] / 1e6:.1f} |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)
    if has_dram:
        j = {f: dict(launches=v[0], share=v[1] / s, dram_bytes_per_launch=v[2] / v[0]) for f, v in fam.items()}
        jp = os.path.join(OUT, f"{TAG}_traffic.json")
        json.dump(dict(source=os.path.basename(src), families=j), open(jp, "w"), indent=1)
        print("wrote", jp)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct"]


WANT += ["sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__cluster_size", "sm__inst_executed_pipe_uniform.sum"]


def ncu(rep, dst, note):
    """rep: an .ncu-rep, or the `--page raw --csv` export of one (made on the GPU box: reports exceed its return limit)."""
    if not os.path.exists(rep):
        return
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full: {note}", ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        lines.append(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or any(h == w + s for w in WANT for s in (".per_second", ".pct_of_peak_sustained_elapsed")):
                lines.append(f"{h} [{u}] = {v}")
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    for name in sorted(os.listdir(GO)):
        p = os.path.join(GO, name)
        if ONLY and ONLY not in name:
            continue
        if name.startswith("profile_") and name.endswith(".json"):
            per_op(p, os.path.join(OUT, f"{TAG}_per_op_{name[8:-5].replace('_' + ONLY, '') if ONLY else name[8:-5]}.md"), f"Per-op profile ({name})")
        elif name.startswith("launches") and name.endswith(".csv"):
            launches(p, os.path.join(OUT, f"{TAG}_ncu_{name[:-4].replace('_' + ONLY, '') if ONLY else name[:-4]}.md"))
        elif name.endswith(".ncu-rep"):
            ncu(p, os.path.join(OUT, f"{TAG}_ncu_{name[:-8]}.txt"), name)
        elif name.endswith("_raw.csv"):
            ncu(p, os.path.join(OUT, f"{TAG}_ncu_{name[:-8].replace(ONLY + '_', '') if ONLY else name[:-8]}.txt"), name)
