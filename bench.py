#!/usr/bin/env python3
"""Benchmark of the hot path named by BASELINE.json: YOLOX-M-P6 1280x1280 inference
(forward + decode + class-aware NMS [+ all-gather of detections at N>1]), images/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); rank 0 prints ONE JSON line.
  value     device-resident images/s (inputs already in HBM), max-over-ranks device time
  e2e       same metric through the public API with pinned HOST buffers: H2D of the batch and D2H of the
            detections inside the timed region (copies double-buffered against compute)
  roofline  all launches of the dominant kernel (conv_gemm_kernel): algorithmic conv FLOPs / their
            summed CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = dict(name="YOLOX-M-P6", depth=0.67, width=0.75, act="hard_swish", num_classes=80, strides=(8, 16, 32, 64))
CONF_THR, NMS_THR, MAX_NMS, MAX_DET = 0.001, 0.65, 5000, 300


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=1280)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_reference / configs / strong_bs64 (quick runs)")
    ap.add_argument("--profile-out", default="")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        """Start early (nvidia-smi needs ~1 s to produce its first row); mark_begin/mark_end bracket the timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.06]
        rows = inside if inside else [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm), in_timed_region=bool(inside))


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_run(size, n_images, steps, warmup):
    """Times oracle forward (fp32, torch CPU ops, all host threads) + C decode + C NMS per image."""
    import numpy as np
    import torch
    from oracle import model_ref as mr
    from oracle import post_ref as pr
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = mr.ModelCfg("p6", MODEL["depth"], MODEL["width"], MODEL["act"], MODEL["num_classes"])
    sd = mr.fold_bn(mr.synth_train_state(cfg, 0, calibrate=False))
    sd = mr.apply_masks(sd, mr.magnitude_masks(mr.synth_train_state(cfg, 0, calibrate=False), 49.0))
    x = torch.rand(1, 3, size, size) * 255
    hw = mr.level_hw(cfg, size, size)

    def one_image():
        reg, obj, cls = mr.forward_raw(sd, cfg, x)
        boxes, oc, cc = pr.decode_infer(reg[0].numpy(), obj[0].numpy(), cls[0].numpy(), hw, cfg.strides)
        n = int((cc.max(-1) >= CONF_THR).sum())
        pr.nms_image_main(boxes, oc, cc, CONF_THR, NMS_THR, MAX_NMS, MAX_DET, pr.torchvision_mode(min(n, MAX_NMS), "cpu"))

    for _ in range(warmup):
        one_image()
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(n_images):
            one_image()
    dt = time.perf_counter() - t0
    return dict(images_per_s=steps * n_images / dt, seconds=dt, cores=torch.get_num_threads(),
                sample=f"{steps} steps x {n_images} image(s) of {size}x{size}, fp32, forward+decode+NMS, "
                       f"49% masked random-init weights")


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_run(args.size, 1, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = dict(impl="reference", metric="images/sec YOLOX-M-P6 1280x1280 inference (forward+decode+NMS)",
                value=r["images_per_s"], unit="images/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * r["seconds"] / max(1, args.steps), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="fp32", data="synthetic",
                config=dict(workload=f"{MODEL['name']} {args.size}x{args.size}, 1 image per step (bounded sample of the bs64 workload)",
                            model=MODEL["name"]),
                cpu_baseline=dict(value=r["images_per_s"], unit="images/s", cores=r["cores"], kind="port", sample=r["sample"]),
                e2e=dict(value=r["images_per_s"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_model(device, kind="p6", depth=None, width=None, act=None, masks="magnitude49"):
    """Random-init model of BASELINE.json's architecture.  config 3: 49 % global-magnitude masks over the non-head 4-D
    tensors (01_mask_generator.py rule), run dense-with-zeros; masks="two_four": the synthetic 2:4-compliant set (keep the
    two largest |w| of every four input channels), which the engine may run on the sparse tensor-core path."""
    import torch
    import yolox_b200 as yb
    torch.manual_seed(0)
    depth, width, act = depth or MODEL["depth"], width or MODEL["width"], act or MODEL["act"]
    cls_ = yb.infer.YOLOXP6 if kind == "p6" else yb.infer.YOLOX
    model = cls_(depth, width, act=act, num_classes=MODEL["num_classes"])
    ws = [p for n, p in model.named_parameters() if "head" not in n and p.dim() == 4]
    with torch.no_grad():
        if masks == "magnitude49":
            allw = torch.cat([p.detach().abs().clamp_max(1.0).flatten() for p in ws])
            thr = allw.kthvalue(int(len(allw) * 0.49) + 1).values
            for p in ws:
                p.mul_((p.abs() > thr).to(p.dtype))
        elif masks == "two_four":
            for p in ws:
                co, ci, kh, kw = p.shape
                if ci % 4:
                    continue
                a = p.abs().permute(0, 2, 3, 1).reshape(-1, 4)
                m = torch.zeros_like(a, dtype=torch.bool)
                m.scatter_(1, a.argsort(dim=1, descending=True, stable=True)[:, :2], True)
                p.mul_(m.reshape(co, kh, kw, ci).permute(0, 3, 1, 2).to(p.dtype))
    return model.to(device).half().eval()


# ------------------------------------------------------------------------------------------
# the reference's own GPU path (SURVEY 8d "reference GPU baseline"): torch + cuDNN fp16, cudnn.benchmark (tools/eval.py:122),
# the op sequence of the reference's modules (oracle/model_ref.py run on CUDA), decode + per-image batched_nms loop
# (postprocess_utils.py:27-129), timed like speed_evaluation.py:33-44 -- in this process, on this box
# ------------------------------------------------------------------------------------------
def gpu_reference_run(model, size, batch, steps):
    import torch
    import torchvision
    from oracle import model_ref as mr
    cfg = mr.CONFIGS["yolox_m_p6"]
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    sd = {k: v.detach() for k, v in model.state_dict().items()}      # the SAME masked fp16 weights the engine runs
    hw = mr.level_hw(cfg, size, size)
    grids, strides = (t.cuda() for t in mr.grids_and_strides(hw, cfg.strides, torch.float16))

    def post(reg, obj, cls):
        reg, obj, cls = reg.float(), obj.float(), cls.float()
        reg[..., :2].add_(grids).mul_(strides)
        reg[..., 2:].exp_().mul_(strides / 2)
        boxes = torch.stack([reg[..., 0] - reg[..., 2], reg[..., 1] - reg[..., 3], reg[..., 0] + reg[..., 2],
                             reg[..., 1] + reg[..., 3]], -1)
        oc = obj.sigmoid_()
        cc = cls.sigmoid_() * oc
        res = []
        for i in range(reg.shape[0]):
            sc, lab = torch.max(cc[i], -1, keepdim=True)
            m = sc.squeeze(-1) >= CONF_THR
            det = torch.cat((boxes[i], oc[i], sc, lab.float()), 1)[m]
            if det.size(0) > MAX_NMS:
                det = det[torch.argsort(det[:, 5], descending=True)[:MAX_NMS]]
            keep = torchvision.ops.batched_nms(det[:, :4], det[:, 5], det[:, 6], NMS_THR)[:MAX_DET]
            res.append(det[keep])
        return res

    def one(x, sdi):
        x = x.clone().mul_(0.9).add_(11.4)                          # main.py:164
        return post(*mr.forward_raw(sdi, cfg, x))

    out = {}
    x64 = (torch.rand(batch, 3, size, size, device="cuda") * 255).half()
    for fmt in ("nchw", "channels_last"):
        cl = fmt == "channels_last"
        sdi = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 and cl else v) for k, v in sd.items()}
        xi = x64.contiguous(memory_format=torch.channels_last) if cl else x64
        for _ in range(3):
            one(xi, sdi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one(xi, sdi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        e0.record()
        for _ in range(steps):                                       # the network alone (no decode / NMS)
            mr.forward_raw(sdi, cfg, xi)
        e1.record()
        torch.cuda.synchronize()
        ms_fwd = e0.elapsed_time(e1) / steps
        x1 = xi[:1].contiguous(memory_format=torch.channels_last) if cl else xi[:1].contiguous()
        for _ in range(8):
            one(x1, sdi)
        lat = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            one(x1, sdi)
            b.record()
            torch.cuda.synchronize()
            lat.append(a.elapsed_time(b))
        out[fmt] = dict(images_per_s=batch / ms * 1e3, ms_per_step=ms, forward_only_ms_per_step=ms_fwd,
                        forward_only_images_per_s=batch / ms_fwd * 1e3, latency_bs1_ms_p50=statistics.median(lat))
        del sdi, xi
    torch.backends.cudnn.benchmark = prev
    torch.cuda.empty_cache()
    best = max(out.values(), key=lambda d: d["images_per_s"])
    return dict(what="torch + cuDNN fp16, cudnn.benchmark=True, the reference modules' op sequence (oracle/model_ref.py on CUDA) + "
                     "decode + per-image torchvision.batched_nms, same masked weights, same box, same process",
                batch=batch, size=size, steps=steps, nchw=out["nchw"], channels_last=out["channels_last"],
                best_images_per_s=best["images_per_s"], best_latency_bs1_ms_p50=min(d["latency_bs1_ms_p50"] for d in out.values()),
                best_forward_only_images_per_s=max(d["forward_only_images_per_s"] for d in out.values()),
                note="the synthetic workload keeps ~all 34 000 anchors per image above conf 0.001, the worst case for the reference's "
                     "per-image Python loop (argsort + batched_nms + host syncs); forward_only_* isolates the network",
                torch=torch.__version__)


def time_steps(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def network_config_line(name, kind, depth, width, act, batch, size, steps, dev, peaks):
    """One of BASELINE.json's other network configurations on this GPU: images/s of forward + decode + NMS (device-resident
    uint8-free fp16 input larger than L2), with the conv family's roofline fraction from the per-op profile."""
    import torch
    import yolox_b200 as yb
    from yolox_b200 import postprocess as pp
    model = build_model(dev, kind, depth, width, act, masks="none")
    strides = (8, 16, 32, 64) if kind == "p6" else (8, 16, 32)
    x = (torch.rand(batch, 3, size, size, device=dev) * 255).half()
    net = []

    def step():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng, reg8, cls = model.run_engine(x)
        b.record()
        net.append((a, b))
        pp.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :MODEL["num_classes"]], model.head.hw, strides, CONF_THR, NMS_THR,
                       MAX_NMS, MAX_DET)

    ms = time_steps(step, steps)
    net_ms = sum(a.elapsed_time(b) for a, b in net[-steps:]) / steps
    eng = model.engine_for(x)
    prof = eng.profile(x, iters=3)
    conv = [p for p in prof if p["kind"] == 0]
    flops, conv_share = sum(p["flops"] for p in conv), sum(p["ms"] for p in conv) / sum(p["ms"] for p in prof)
    tf = flops / (net_ms * conv_share * 1e-3) / 1e12
    floor_ms = sum(max(p["flops"] / (peaks["tflops"] * 1e12), p["bytes"] / (peaks["hbm_gbs"] * 1e9)) for p in prof) * 1e3
    del model, eng
    torch.cuda.empty_cache()
    return dict(workload=f"{name} {size}x{size}, {batch} images/step, fp16, random-init dense weights, forward+decode+NMS",
                images_per_s=batch / ms * 1e3, ms_per_step=ms, network_ms_in_step=net_ms,
                roofline=dict(bound="tensor", kernel="conv_gemm_kernel", achieved=tf, peak=peaks["tflops"], unit="TFLOP/s",
                              frac=tf / peaks["tflops"], layerwise_floor_ms=floor_ms, step_over_floor=ms / floor_ms))


def post_stress_lines(batch, steps, dev, peaks):
    """BASELINE config 4: no network -- synthetic head tensors (B, 34000, 80) through fused decode + threshold + top-5000 +
    class-aware NMS (main.py semantics), two distributions: max-candidate and clustered."""
    import torch
    from yolox_b200 import postprocess as pp
    strides, S = (8, 16, 32, 64), 1280
    hw = [(S // s, S // s) for s in strides]
    A, C = sum(h * w for h, w in hw), 80
    g = torch.Generator(device=dev).manual_seed(4)
    out = {}
    for dist_name in ("max_candidate", "clustered"):
        if dist_name == "max_candidate":
            reg = torch.randn(batch, A, 4, device=dev, generator=g).half()
            obj = (torch.randn(batch, A, 1, device=dev, generator=g) * 2 - 2).half()
            cls = (torch.randn(batch, A, C, device=dev, generator=g) * 2 - 2).half()
        else:
            # 200 boxes per image; every anchor inside a box regresses to it (N(0, 0.05) jitter) with a +6 one-hot class logit
            grids, scales = pp.yolox_generate_grid(S, strides)
            gx = ((grids[0, :, 0] + 0.5) * scales[0, :, 0]).to(dev)
            gy = ((grids[0, :, 1] + 0.5) * scales[0, :, 0]).to(dev)
            sc = scales[0, :, 0].to(dev)
            reg = torch.randn(batch, A, 4, device=dev, generator=g) * 0.05
            obj = torch.full((batch, A, 1), -6.0, device=dev)
            cls = torch.full((batch, A, C), -6.0, device=dev)
            for b in range(batch):
                ctr = torch.rand(200, 2, device=dev, generator=g) * S
                wh = torch.rand(200, 2, device=dev, generator=g) * 200 + 30
                lab = torch.randint(0, C, (200,), device=dev, generator=g)
                inside = ((gx[None] - ctr[:, :1]).abs() < wh[:, :1] / 2) & ((gy[None] - ctr[:, 1:]).abs() < wh[:, 1:] / 2)   # [200, A]
                owner = torch.where(inside.any(0), inside.float().argmax(0), torch.full((A,), -1, device=dev, dtype=torch.long))
                m = owner >= 0
                o = owner[m]
                reg[b, m, 0] += (ctr[o, 0] - (gx[m] - 0.5 * sc[m])) / sc[m]
                reg[b, m, 1] += (ctr[o, 1] - (gy[m] - 0.5 * sc[m])) / sc[m]
                reg[b, m, 2] += torch.log(wh[o, 0] / sc[m])
                reg[b, m, 3] += torch.log(wh[o, 1] / sc[m])
                obj[b, m, 0] = 3.0
                cls[b, m, lab[o]] = 6.0
            reg, obj, cls = reg.half(), obj.half(), cls.half()
        res = {}

        def step():
            res["o"] = pp.detect_main(reg, obj, cls, hw, strides, CONF_THR, NMS_THR, MAX_NMS, MAX_DET)

        ms = time_steps(step, steps)
        det, cnt, _ = res["o"]
        n_cand = float(((torch.sigmoid(cls.float()) * torch.sigmoid(obj.float())).max(-1).values >= CONF_THR).sum()) / batch
        alg_bytes = batch * (A * (5 + C) * 2 + n_cand * 28)
        out[dist_name] = dict(images_per_s=batch / ms * 1e3, ms_per_step=ms, candidates_per_image=n_cand,
                              detections_per_image=float(cnt.float().mean()),
                              roofline=dict(bound="hbm", kernel="select_infer + sort_keys + nms (whole decode+NMS step)",
                                            achieved=alg_bytes / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                                            frac=alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]))
    return dict(workload=f"decode + NMS only, B={batch}, A={A}, C={C}, conf {CONF_THR} nms {NMS_THR} top-{MAX_NMS}/{MAX_DET}", **out)


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, so the pinned host batches (and the staging the
    H2D DMA reads) live on that GPU's NUMA node -- with 8 ranks streaming 24 GB/s each, remote-node buffers throttle the
    end-to-end number.  Best effort: returns a note for the JSON line."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        try:      # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by its PCI bus id
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"cpu affinity = GPU {local_rank}'s NUMA node ({len(os.sched_getaffinity(0))} cpus)"
    except Exception as e:  # noqa: BLE001 (no NVML / not permitted: run unbound)
        return f"unbound ({type(e).__name__})"


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import yolox_b200 as yb
    from yolox_b200 import postprocess as pp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    B, S = args.batch, args.size
    numa = bind_to_gpu_numa_node(local_rank)   # before any pinned allocation: first touch decides the NUMA node
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    model = build_model(dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    # the batch as a camera / decoder delivers it: uint8 pixel values 0..255, NCHW.  The engine reads uint8 directly (YX_U8:
    # evaluated exactly as the same values in fp16, x*0.9+11.4 fused into the read), so a step moves 315 MB over PCIe
    # instead of the 629 MB of an fp16 batch.
    host_imgs = [torch.empty(B, 3, S, S, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for h in host_imgs:
        h.copy_(torch.randint(0, 256, (B, 3, S, S), generator=gen, dtype=torch.uint8))
    dev_img = host_imgs[0].to(dev, non_blocking=True)
    strides = MODEL["strides"]

    # N > 1: every rank ends the step holding the detections of the whole global batch.  Default: the NMS kernel stores
    # its rows straight into every rank's window over NVLink (yx_detect_main_gather); YX_PEER_GATHER=0 selects the
    # separate NCCL all-gather it replaces.
    gathered, peer = None, None
    if world > 1:
        if os.environ.get("YX_PEER_GATHER", "1") != "0":
            try:
                peer = yb.dist.PeerGather(B, MAX_DET, dev)
            except RuntimeError as e:     # raised on every rank together (CUDA IPC not permitted on this box)
                if rank == 0:
                    print(f"bench: {e}; using the NCCL all-gather", file=sys.stderr, flush=True)
        if peer is None:
            gathered = torch.empty(world, B, MAX_DET * 7 + 1, dtype=torch.float32, device=dev)

    net_events = []   # (start, end) CUDA events around the network's launches of every timed step -> roofline.achieved

    def step(img, mark=False):
        if mark:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
        eng, reg8, cls = model.run_engine(img, in_scale=0.9, in_shift=11.4)
        if mark:
            eb.record()
            net_events.append((ea, eb))
        det, cnt, _ = pp.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :MODEL["num_classes"]], model.head.hw, strides,
                                     CONF_THR, NMS_THR, MAX_NMS, MAX_DET, gather=peer)
        if gathered is not None:  # one all-gather of fixed-shape detections (+count packed as a trailing column)
            packed = torch.cat([det.view(B, -1), cnt.view(B, 1).float()], dim=1)
            dist.all_gather_into_tensor(gathered.view(world * B, -1), packed)
        return det, cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        det, cnt = step(dev_img)
    barrier()
    n_launch_step = model.engine_for(dev_img).n_launches + 3 + (1 if peer is not None else 0)  # + select, sort, nms (+ gather wait)

    # ---- device-resident timing -------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        det, cnt = step(dev_img, mark=True)
    if peer is not None:
        peer.wait()          # the last step's rows of every rank have arrived in this rank's window (steps before it were
                             # completed by the following step's own pre-wait): the gather is inside the timed region
    ev1.record()
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    net_ms_step = sum(a.elapsed_time(b) for a, b in net_events) / max(len(net_events), 1)   # network launches, in the timed region
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())

    # ---- end-to-end: pinned host input -> H2D -> forward+detect -> D2H detections --------------
    copy_stream, comp_stream = torch.cuda.Stream(), torch.cuda.current_stream()
    dbuf = [torch.empty_like(dev_img), torch.empty_like(dev_img)]
    h_det = torch.empty(B, MAX_DET, 7, dtype=torch.float32).pin_memory()
    h_cnt = torch.empty(B, dtype=torch.int32).pin_memory()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(k):
        with torch.cuda.stream(copy_stream):
            dbuf[0].copy_(host_imgs[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(k):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < k:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done[nxt]) if i >= 1 else None
                    dbuf[nxt].copy_(host_imgs[nxt], non_blocking=True)
                    ready[nxt].record(copy_stream)
            comp_stream.wait_event(ready[cur])
            d, c = step(dbuf[cur])
            done[cur].record(comp_stream)
            h_det.copy_(d, non_blocking=True)
            h_cnt.copy_(c, non_blocking=True)
        if peer is not None:
            peer.wait()
        torch.cuda.synchronize()

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())

    # ---- strong scaling of the stated configuration: GLOBAL batch 64, 64 / N images per GPU ----------------
    strong = None
    if not args.no_extras and B * 1 == 64 and 64 % world == 0:
        Bs = 64 // world
        if world == 1:
            strong = dict(global_batch=64, per_gpu_batch=64, images_per_s=B * world * args.steps / (ms_dev * 1e-3),
                          ms_per_step=ms_dev / args.steps, note="N = 1: this is the headline measurement itself")
        else:
            xs = dev_img[:Bs].contiguous()
            peer_s = None
            if peer is not None:
                peer_s = yb.dist.PeerGather(Bs, MAX_DET, dev)

            def sstep():
                eng, reg8, cls = model.run_engine(xs, in_scale=0.9, in_shift=11.4, use_graph=True)   # the network as one CUDA graph
                d, c, _ = pp.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :MODEL["num_classes"]], model.head.hw, strides,
                                         CONF_THR, NMS_THR, MAX_NMS, MAX_DET, gather=peer_s)
                if peer_s is None:
                    dist.all_gather_into_tensor(torch.empty(world * Bs, MAX_DET * 7 + 1, device=dev),
                                                torch.cat([d.view(Bs, -1), c.view(Bs, 1).float()], dim=1))

            for _ in range(max(args.warmup, 3)):
                sstep()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_strong = args.steps * 4
            s0.record()
            for _ in range(n_strong):
                sstep()
            if peer_s is not None:
                peer_s.wait()
            s1.record()
            barrier()
            t = torch.tensor([s0.elapsed_time(s1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t.item()) / n_strong
            strong = dict(global_batch=64, per_gpu_batch=Bs, images_per_s=64 / ms_s * 1e3, ms_per_step=ms_s, steps=n_strong,
                          mode="network replayed as one CUDA graph per step + fused decode/NMS with the peer-store gather",
                          gather_status=peer_s.status() if peer_s is not None else 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- bs1 latency, p50: the whole step (network + decode + NMS) replayed as ONE CUDA graph -------
    x1 = dev_img[:1].contiguous()
    lat_pred = yb.predict.Predictor(model, CONF_THR, NMS_THR, MAX_NMS, MAX_DET, in_scale=0.9, in_shift=11.4, whole_graph=True)
    lat = []
    for i in range(60):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lat_pred(x1)
        b.record()
        torch.cuda.synchronize()
        if i >= 10:
            lat.append(a.elapsed_time(b))
    lat_p50 = statistics.median(lat)

    # ---- roofline of the dominant kernel: per-op CUDA-event times of the same engine -------------
    peaks = measured_peaks()
    eng = model.engine_for(dev_img)
    prof = eng.profile(dev_img, iters=3)
    conv = [p for p in prof if p["kind"] == 0]
    conv_ms, conv_flops = sum(p["ms"] for p in conv), sum(p["flops"] for p in conv)
    all_ms = sum(p["ms"] for p in prof)
    # achieved = algorithmic FLOPs of the conv launches of one step / the time those launches took INSIDE the timed region:
    # CUDA events bracket the network's launches of every timed step (net_ms_step); the per-op profile only supplies the
    # conv kernels' share of the network's launches (the remaining launches are s2d / SPP).
    conv_share_net = conv_ms / all_ms
    conv_ms_in_step = net_ms_step * conv_share_net
    achieved_tf = conv_flops / (conv_ms_in_step * 1e-3) / 1e12
    per_op_tf = conv_flops / (conv_ms * 1e-3) / 1e12
    hbm_bound = [p for p in conv if p["flops"] / max(p["bytes"], 1) < peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)]
    hbm_ms, hbm_bytes = sum(p["ms"] for p in hbm_bound), sum(p["bytes"] for p in hbm_bound)
    traffic, traffic_src = None, None   # measured DRAM bytes per conv launch, from the committed ncu launch list
    import glob
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))   # newest round's ncu launch list
    if tps:
        tj = json.load(open(tps[-1]))
        fam = tj["families"].get("conv_gemm_kernel")
        if fam:
            traffic, traffic_src = fam["dram_bytes_per_launch"], f"profiles/{os.path.basename(tps[-1])} ({tj['source']})"
    roof = dict(bound="tensor", kernel="conv_gemm_kernel (all instantiations)", achieved=achieved_tf, peak=peaks["tflops"],
                unit="TFLOP/s", frac=achieved_tf / peaks["tflops"], traffic=traffic, traffic_source=traffic_src,
                algorithmic_bytes_per_launch=sum(p["bytes"] for p in conv) / max(len(conv), 1),
                algorithmic_flops_per_launch=conv_flops / max(len(conv), 1), peak_source=peaks["source"],
                launches_per_step=len(conv), avg_launch_ms_in_step=conv_ms_in_step / max(len(conv), 1),
                network_ms_in_step=net_ms_step, share_of_step=conv_ms_in_step / (ms_dev / args.steps),
                per_op_back_to_back=dict(achieved=per_op_tf, frac=per_op_tf / peaks["tflops"],
                                         note="each launch timed on its own after the step loop (yx_engine_profile): "
                                              "less power-throttled than inside the step"),
                hbm_bound_layers=dict(n=len(hbm_bound), achieved_gbs=hbm_bytes / max(hbm_ms, 1e-9) / 1e6,
                                      peak_gbs=peaks["hbm_gbs"], note="per-op timing"))
    if args.profile_out:
        with open(args.profile_out, "w") as f:
            json.dump(dict(batch=B, size=S, ops=prof, peaks=peaks), f, indent=1)

    gpu_ref, configs = None, None
    if not args.no_extras and world == 1:
        try:
            gpu_ref = gpu_reference_run(model, S, B, steps=5)
            ours_ips = B * world * args.steps / (ms_dev * 1e-3)
            gpu_ref["ours_over_best_reference_throughput"] = ours_ips / gpu_ref["best_images_per_s"]
            gpu_ref["ours_network_over_best_reference_forward_only"] = (B / net_ms_step * 1e3) / gpu_ref["best_forward_only_images_per_s"]
            gpu_ref["ours_latency_bs1_ms_p50"] = lat_p50
        except Exception as e:  # noqa: BLE001  (the baseline must never cost the headline line)
            gpu_ref = dict(error=f"{type(e).__name__}: {e}")
        torch.cuda.empty_cache()
        try:
            model._engines = {}     # free the headline engines' arenas before building the other configurations
            torch.cuda.empty_cache()
            configs = dict(
                m_640_bs64=network_config_line("YOLOX-M", "yolox", 0.67, 0.75, "silu", 64, 640, args.steps, dev, peaks),
                l_640_bs32_per_gpu=network_config_line("YOLOX-L", "yolox", 1.0, 1.0, "silu", 32, 640, args.steps, dev, peaks),
                post_stress_b64=post_stress_lines(64, args.steps, dev, peaks))
        except Exception as e:  # noqa: BLE001
            configs = dict(error=f"{type(e).__name__}: {e}")

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(S, 1, 8, 1)
        cpu = dict(value=r["images_per_s"], unit="images/s", cores=r["cores"], kind="port", sample=r["sample"])

    imgs = B * world * args.steps
    h2d = dev_img.numel() * dev_img.element_size()
    d2h = h_det.numel() * 4 + h_cnt.numel() * 4
    line = dict(metric="images/sec YOLOX-M-P6 1280x1280 inference (forward+decode+NMS)", value=imgs / (ms_dev * 1e-3),
                unit="images/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp16",
                data="synthetic",
                config=dict(workload=f"pruned {MODEL['name']} {S}x{S}, {B} images/GPU/step (global {B * world}), 49% synthetic "
                                     f"magnitude masks dense-with-zeros, uint8 input batch, conf {CONF_THR} nms {NMS_THR} top-{MAX_NMS}/{MAX_DET}, "
                                     f"random-init preds => ~all {sum((S // s) ** 2 for s in strides)} anchors/img are candidates",
                            model=MODEL["name"], global_batch=B * world, parallelism=(f"dp{world} batch shard, detections gathered " +
                                         ("by the NMS kernel into peer windows (NVLink stores)" if peer is not None
                                          else "with one NCCL all-gather")) if world > 1 else "single GPU",
                            l2="inputs larger than L2 (activations per layer >> 126 MB at bs64)"),
                e2e=dict(value=imgs / (ms_e2e * 1e-3), unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         note="pinned host uint8 NCHW batch (pixel values 0..255, as decoded images arrive) -> H2D (double-buffered) "
                              "-> x*0.9+11.4 fused into the image read -> forward+decode+NMS -> D2H detections",
                         host_numa=numa),
                gpu_launches=n_launch_step * args.steps, clocks=clocks, roofline=roof, cpu_baseline=cpu,
                latency_bs1_ms_p50=lat_p50, detections_last_step=int(cnt.sum().item()),
                launch_shapes=eng.shape_source)
    if strong is not None:
        line["strong_bs64"] = strong
    if gpu_ref is not None:
        line["gpu_reference"] = gpu_ref
    if configs is not None:
        line["configs"] = configs
    if world > 1:
        line["gather"] = dict(kind="nms_fused_peer_store (3 windows, wait deferred by one step)" if peer is not None else "nccl_all_gather",
                              status=peer.status() if peer is not None else 0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
