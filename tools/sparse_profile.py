"""Per-op comparison of the 2:4 sparse tensor-core variant against dense-with-zeros on BASELINE config 3's second mask
set (YOLOX-M-P6 1280x1280, 2:4 masks over the non-head convs): three engines on the same weights --
YX_SPARSE=0 (dense-with-zeros, tuned), YX_SPARSE=force (every eligible layer sparse, heuristic shape) and the default
(the tuner picks per layer by measurement) -- each profiled op by op with CUDA events.  Writes a markdown table with the
dense-equivalent and the 50 %-FLOP TFLOP/s of every layer that has a sparse form.
    python tools/sparse_profile.py [batch] [size] [out.md]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/sparse_profile.md"
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, device=dev)
os.environ["YX_TUNE_CACHE"] = "0"
res = {}
for mode, env in (("dense", "0"), ("forced", "force"), ("tuned", "1")):
    os.environ["YX_SPARSE"] = env
    model = bench.build_model(dev, masks="two_four")
    eng = model.engine_for(x)
    for _ in range(3):
        eng.run(x, 0.9, 11.4)
    torch.cuda.synchronize()
    res[mode] = eng.profile(x, iters=5)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        eng.run(x, 0.9, 11.4)
    b.record()
    torch.cuda.synchronize()
    res[mode + "_step_ms"] = a.elapsed_time(b) / 10
    del model, eng
    torch.cuda.empty_cache()
peaks = bench.measured_peaks()
lines = [f"# 2:4 sparse tensor-core path vs dense-with-zeros (batch {B}, {S}x{S}, YOLOX-M-P6, 2:4 masks on the non-head convs)", "",
         f"Network time per step (CUDA events, 10 steps): dense-with-zeros {res['dense_step_ms']:.2f} ms, every eligible layer forced sparse "
         f"{res['forced_step_ms']:.2f} ms, tuner's per-layer choice {res['tuned_step_ms']:.2f} ms.", "",
         "`dense` = tuned dense-with-zeros launch; `sparse` = tcgen05.mma.sp variant (heuristic shape); `TF/s eq` counts the masked zeros as "
         "FLOPs (dense-equivalent), `TF/s 50%` counts the multiplications the sparse MMA really performs; peak "
         f"{peaks['tflops']:.0f} TFLOP/s dense ({peaks['source']}).", "",
         "| # | op | GFLOP | dense ms | sparse ms | speed-up | sparse TF/s eq | sparse TF/s 50% | tuner picked | sparse launch shape |", "|---|---|---|---|---|---|---|---|---|---|"]
n_sp = n_pick = 0
sd = ss = 0.0
for i, (d, f, t) in enumerate(zip(res["dense"], res["forced"], res["tuned"])):
    if "sparse24" not in f["shape"]:
        continue
    n_sp += 1
    picked = "sparse24" in t["shape"]
    n_pick += picked
    sd += d["ms"]; ss += f["ms"]
    tf = f["flops"] / (f["ms"] * 1e-3) / 1e12
    lines.append(f"| {i} | {f['name']} | {f['flops'] / 1e9:.1f} | {d['ms']:.3f} | {f['ms']:.3f} | {d['ms'] / f['ms']:.2f} | {tf:.0f} | {tf / 2:.0f} | "
                 f"{'sparse' if picked else 'dense'} | {f['shape'].split(': ')[-1]} |")
lines += ["", f"{n_sp} layers have a sparse form; summed: dense {sd:.2f} ms, sparse {ss:.2f} ms; the tuner picked sparse for {n_pick} of them."]
open(out, "w").write("\n".join(lines) + "\n")
json.dump(res, open(out.replace(".md", ".json"), "w"))
print("\n".join(lines[:4]))
print(lines[-1])
