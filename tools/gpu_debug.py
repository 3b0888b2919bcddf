"""GPU bring-up driver (diagnostics, not a test): each stage prints numeric evidence and exits 0/1.
Run one conv case per process so a faulting kernel cannot poison later cases:
    python tools/gpu_debug.py conv <case-index>
    python tools/gpu_debug.py post
    python tools/gpu_debug.py model <config> <H> <W> <B>
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

torch.set_grad_enabled(False)


def stage_conv(i):
    from tests.conv_util import CASES, run_conv_case, tolerance
    case = CASES[i]
    t = time.time()
    r = run_conv_case(**case)
    ok = r["max_err"] <= tolerance(case) and not r["clobbered"] and not r["pad_nonzero"]
    print(f"conv[{i}] {case} -> max_err {r['max_err']:.4g} mean_err {r['mean_err']:.4g} ref_scale {r['ref_scale']:.3g} "
          f"clobbered {r['clobbered']} pad_nonzero {r['pad_nonzero']} {'OK' if ok else 'FAIL'} ({time.time()-t:.1f}s)")
    if not ok:
        err = (r["out"] - r["ref"]).abs()  # [B,H,W,C]
        B, H, W, C = err.shape
        print("  err by image:", [round(float(err[b].max()), 3) for b in range(B)])
        print("  err by row y (first 24):", [round(float(err[:, y].max()), 2) for y in range(min(H, 24))])
        print("  err by col x (first 24):", [round(float(err[:, :, x].max()), 2) for x in range(min(W, 24))])
        print("  err by channel/8:", [round(float(err[..., c:c + 8].max()), 2) for c in range(0, C, 8)][:40])
        print("  out sample:", r["out"][0, 0, 0, :8].tolist())
        print("  ref sample:", r["ref"][0, 0, 0, :8].tolist())
        print("  out sample (1,1):", r["out"][0, 1, 1, :8].tolist())
        print("  ref sample (1,1):", r["ref"][0, 1, 1, :8].tolist())
        print("  frac wrong:", float((err > 0.02).float().mean()), "nan:", int(torch.isnan(r['out']).sum()))
    return ok


def stage_post():
    import glob
    import yolox_b200 as yb
    from oracle import post_ref as pr
    ok_all = True
    for path in sorted(glob.glob("tests/golden/post_*.npz")):
        g = np.load(path)
        img, strides = int(g["img"]), [int(s) for s in g["strides"]]
        hw = [(img // s, img // s) for s in strides]
        conf, thr = float(g["conf"]), float(g["nms_thr"])
        dev = "cuda"
        reg, obj, cls = (torch.from_numpy(g[k]).to(dev) for k in ("reg", "obj", "cls"))
        grids, scales = yb.postprocess.yolox_generate_grid(img, strides, torch.float16)
        boxes, oc, cc = yb.postprocess.yolox_postprocess_output_torch_batch(reg, obj, cls, grids.to(dev), scales.to(dev))
        e1 = float((boxes.cpu() - torch.from_numpy(g["boxes"])).abs().max())
        e2 = float((cc.cpu() - torch.from_numpy(g["cls_conf"])).abs().max())
        # NMS on the reference's decoded tensors (identical inputs)
        rb, ro, rc = (torch.from_numpy(g[k]).to(dev) for k in ("boxes", "obj_conf", "cls_conf"))
        dets = yb.postprocess.yolox_nms_torch_batch(rb, ro, rc, nms_threshold=thr, conf_threshold=conf)
        same = True
        for i, d in enumerate(dets):
            want = g[f"main_det_{i}"]
            got = d.cpu().numpy() if d is not None else np.zeros((0, 7), np.float32)
            if got.shape != want.shape or not np.array_equal(got, want):
                same = False
                print(f"  {os.path.basename(path)} img {i}: got {got.shape} want {want.shape}")
                n = min(len(got), len(want))
                bad = np.nonzero((got[:n] != want[:n]).any(1))[0]
                print("   first differing rows:", bad[:5], got[bad[:2]] if len(bad) else "", want[bad[:2]] if len(bad) else "")
        # fused path from logits
        det, cnt, anc = yb.postprocess.detect_main(reg, obj, cls, hw, strides, conf, thr)
        det2, cnt2, _ = yb.postprocess.nms_main_raw(boxes, oc, cc, thr, conf, 5000, 300)
        fused_same = bool(torch.equal(det, det2) and torch.equal(cnt, cnt2))
        # yolox flavour
        pred = torch.from_numpy(g["yolox_pred"]).to(dev)
        res = yb.postprocess.postprocess(pred, pred.shape[2] - 5, conf, thr)
        ysame = True
        for i, d in enumerate(res):
            want = g[f"yolox_det_{i}"]
            got = d.cpu().numpy() if d is not None else np.zeros((0, 7), np.float32)
            if got.shape != want.shape or not np.array_equal(got, want):
                ysame = False
                print(f"  yolox {os.path.basename(path)} img {i}: got {got.shape} want {want.shape}")
        print(f"post {os.path.basename(path)}: decode |d| boxes {e1:.3g} cls_conf {e2:.3g}; main NMS bit-exact {same}; "
              f"fused==unfused {fused_same}; yolox postprocess bit-exact {ysame}; counts {cnt.tolist()}")
        ok_all = ok_all and same and fused_same and ysame and e1 < 1e-2 and e2 < 1e-5
    return ok_all


def stage_model(name, H, W, B):
    import yolox_b200 as yb
    from oracle import model_ref as mr
    cfg = mr.CONFIGS[name]
    fused = mr.fold_bn(mr.synth_train_state(cfg, 3, calib_hw=(H, W)))
    cls_ = yb.infer.YOLOXP6 if cfg.kind == "p6" else yb.infer.YOLOX
    model = cls_(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    model.load_state_dict(fused, strict=True)
    model = model.cuda().half()
    x = mr.synth_images(11, B, H, W)
    t = time.time()
    reg, obj, cls = model(x.cuda().half())
    torch.cuda.synchronize()
    print(f"model {name} {H}x{W} b{B}: first forward {time.time()-t:.2f}s")
    q = {k: (v.half().float() if k.endswith("weight") else v) for k, v in fused.items()}
    rr, ro, rc = mr.forward_raw(q, cfg, x.half().float())
    ok = True
    for nm, a, b in (("reg", reg, rr), ("obj", obj, ro), ("cls", cls, rc)):
        a = a.float().cpu()
        err = (a - b).abs()
        scale = float(b.abs().mean())
        print(f"  {nm}: max|d| {float(err.max()):.4g} mean|d| {float(err.mean()):.4g} ref mean|x| {scale:.3g} "
              f"nan {int(torch.isnan(a).sum())}")
        ok = ok and float(err.max()) < 0.25 and float(err.mean()) < 0.02
    # per-layer drill-down when wrong: compare every arena buffer with the CPU interpretation of the plan
    if not ok:
        from tests.plan_interp import run_graph_cpu
        eng = model.engine_for(x.cuda().half())
        print("  (drill-down omitted: run tools/gpu_debug.py conv cases)")
    return ok


def stage_drill(name, H, W, B):
    """Every op of a real network, one at a time, vs torch fp32 on the same device inputs."""
    import yolox_b200 as yb
    from oracle import model_ref as mr
    from tests.plan_interp import teacher_forced_errors
    cfg = mr.CONFIGS[name]
    fused = mr.fold_bn(mr.synth_train_state(cfg, 3, calib_hw=(min(H, 640), min(W, 640))))
    cls_ = yb.infer.YOLOXP6 if cfg.kind == "p6" else yb.infer.YOLOX
    model = cls_(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    model.load_state_dict(fused, strict=True)
    model = model.cuda().half()
    x = mr.synth_images(11, B, H, W).cuda().half()
    errs = teacher_forced_errors(model, x)
    worst = sorted(errs, key=lambda e: -e[2])[:12]
    for e in worst:
        print(f"  op {e[0]:3d} {e[1]:45s} ratio {e[2]:8.2f}  abs {e[3]:.4g}")
    bad = [e for e in errs if e[2] > 1.5]
    print(f"drill {name} {H}x{W} b{B}: {len(errs)} ops, {len(bad)} above 1.5x the one-rounding-step slack")
    return not bad


if __name__ == "__main__":
    st = sys.argv[1]
    if st == "conv":
        ok = stage_conv(int(sys.argv[2]))
    elif st == "post":
        ok = stage_post()
    elif st == "drill":
        ok = stage_drill(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    elif st == "model":
        ok = stage_model(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    else:
        raise SystemExit("unknown stage")
    sys.exit(0 if ok else 1)
