"""Graph builder: turns the module tree into the flat op list the native engine executes.

Host logic only (pure Python + torch CPU tensors for weight packing) — unit-tested without a GPU.
The builder implements the fusions that need no kernel support:
  * concat elimination: producers write straight into channel slices of the consumer's buffer
    (torch.cat at network_blocks.py:318, yolo_pafpn_p6.py:154-176 never materialises);
  * CSP conv1+conv2 merged into one GEMM over the shared input (network_blocks.py:314-316);
  * head cls_convs[0] + reg_convs[0] merged (same input, yolo_head.py:140-147);
  * reg_pred + obj_pred merged into one 1x1 conv writing a packed [B,A,8] tensor.
Buffers get arena offsets from a liveness-based first-fit allocator.
"""
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import _capi

ALIGN = 1024


def _rup(a: int, b: int) -> int:
    return (a + b - 1) // b * b


@dataclass
class Buf:
    name: str
    n: int
    h: int
    w: int
    c: int  # channel pitch
    offset: int = -1
    first: int = 10 ** 9
    last: int = -1
    pinned: bool = False  # outputs: kept alive to the end
    nbytes_override: int = 0

    @property
    def nbytes(self) -> int:
        return self.nbytes_override or _rup(self.n * self.h * self.w * self.c * 2, ALIGN)

    def view(self, c_off: int = 0, c: Optional[int] = None) -> "V":
        return V(self, c_off, self.c - c_off if c is None else c)


@dataclass
class V:
    """Channel slice [c_off, c_off+c) of a buffer; optionally a pyramid-level window of a [B,A,C] output."""
    buf: Buf
    c_off: int
    c: int
    # level window (head outputs): element offset of the level's first anchor and the level dims
    lvl_off: int = 0
    h: int = 0
    w: int = 0
    nstride: int = 0

    @property
    def H(self):
        return self.h or self.buf.h

    @property
    def W(self):
        return self.w or self.buf.w

    def to_c(self) -> _capi.View:
        b = self.buf
        assert b.offset >= 0, f"buffer {b.name} not placed"
        v = _capi.View()
        v.offset = b.offset + (self.lvl_off * b.c + self.c_off) * 2
        v.nstride = self.nstride or b.h * b.w * b.c
        v.n, v.h, v.w, v.c, v.pitch = b.n, self.H, self.W, self.c, b.c
        return v


@dataclass
class PlannedOp:
    kind: int
    name: str
    src: Optional[V]
    dst: V
    res: Optional[V] = None
    ksize: int = 1
    stride: int = 1
    act: int = 0
    weight: Optional[torch.Tensor] = None  # fp32 [cout, cin/groups, k, k]
    bias: Optional[torch.Tensor] = None    # fp32 [cout]
    aux: int = 0
    cin_pad: int = 0
    cout_pad: int = 0
    w_offset: int = 0
    b_offset: int = 0
    up: Optional[V] = None      # low-resolution tensor whose x2 nearest upsampling is concatenated in front of src


class Graph:
    def __init__(self, batch: int, in_h: int, in_w: int):
        self.batch, self.in_h, self.in_w = batch, in_h, in_w
        self.bufs: List[Buf] = []
        self.ops: List[PlannedOp] = []
        self.arena_bytes = 0
        self.weight_blob: Optional[torch.Tensor] = None
        self.bias_blob: Optional[torch.Tensor] = None

    # ---- buffers -------------------------------------------------------------------------
    def new_buf(self, name: str, h: int, w: int, c: int, pinned: bool = False) -> Buf:
        assert c % 8 == 0, f"{name}: channel pitch {c} must be a multiple of 8"
        b = Buf(name, self.batch, h, w, c, pinned=pinned)
        self.bufs.append(b)
        return b

    def new_output(self, name: str, anchors: int, c: int) -> Buf:
        """[B, A, c] head output (h=A, w=1 as far as the allocator cares)."""
        return self.new_buf(name, anchors, 1, c, pinned=True)

    def _touch(self, v: Optional[V], idx: int):
        if v is not None:
            v.buf.first = min(v.buf.first, idx)
            v.buf.last = max(v.buf.last, idx)

    def _add(self, op: PlannedOp) -> V:
        idx = len(self.ops)
        self._touch(op.src, idx); self._touch(op.dst, idx); self._touch(op.res, idx); self._touch(op.up, idx)
        self.ops.append(op)
        return op.dst

    # ---- ops -----------------------------------------------------------------------------
    def s2d(self, dst: V, order: str, padded: bool = False) -> V:
        """padded: rows are [0 | W/2 pixels | 0 0 0] for the row-packed stem conv (aux bit 1)."""
        assert dst.c == 16 and dst.buf.c == 16
        return self._add(PlannedOp(_capi.OP_S2D, "s2d", None, dst,
                                   aux=(1 if order == "unshuffle" else 0) | (2 if padded else 0)))

    def conv_rowpack(self, name: str, src: V, dst: V, weight: torch.Tensor, bias: torch.Tensor, act: str) -> V:
        """3x3/s1 conv over the padded 16-channel s2d tensor, executed as 3 vertical taps with K = 48:
        the engine reads the source through an overlapping 64-channel view, so one 128-byte TMA row holds the
        three horizontal neighbours of a pixel.  weight: the ordinary [cout, 12, 3, 3] tensor."""
        cout, cin, k, k2 = weight.shape
        assert (cin, k, k2) == (12, 3, 3) and src.c == 16 and src.buf.c == 16 and dst.c == cout
        assert (dst.H, dst.W) == (src.H, src.W - 4)
        w = torch.zeros(cout, 3, 16, 3, dtype=torch.float32)           # [cout, dx, c(16), dy]
        w[:, :, :12, :] = weight.detach().float().cpu().permute(0, 3, 1, 2)
        w48 = w.reshape(cout, 48, 3, 1)                                  # cin index = dx*16 + c ; kernel (3, 1)
        return self._add(PlannedOp(_capi.OP_CONV, name, src, dst, None, 3, 1, _capi.act_code(act), w48,
                                   bias.detach().float().cpu(), aux=1, cin_pad=48, cout_pad=_rup(cout, 16)))

    def conv_stem_image(self, name: str, dst: V, weight: torch.Tensor, bias: torch.Tensor, act: str, order: str) -> V:
        """The 3x3 stem conv FED FROM THE IMAGE: the kernel's producer warps do the space-to-depth (and the input affine)
        themselves while building the halo operand tiles, so neither the s2d op nor its 16-channel tensor exists.
        Must be the graph's first op.  aux = 4 (image-fed) | 8 (pixel_unshuffle order); ordinary KRSC weights, K = 16."""
        cout, cin, k, k2 = weight.shape
        assert (cin, k, k2) == (12, 3, 3) and dst.c == cout and not self.ops, "the image-fed stem is the first op of a graph"
        assert (dst.H, dst.W) == (self.in_h // 2, self.in_w // 2)
        src = V(dst.buf, dst.c_off, 16)                                  # placeholder: the op reads the image, not the arena
        return self._add(PlannedOp(_capi.OP_CONV, name, src, dst, None, 3, 1, _capi.act_code(act), weight.detach().float().cpu(),
                                   bias.detach().float().cpu(), aux=4 | (8 if order == "unshuffle" else 0), cin_pad=16,
                                   cout_pad=_rup(cout, 16)))

    @staticmethod
    def can_fuse_upsample(up: V) -> bool:
        """The engine folds nearest x2 upsample + concat into a 1x1 conv's loads when the upsampled part is whole
        64-channel chunks of a plain (not level-windowed) buffer."""
        return up.c % 64 == 0 and up.lvl_off == 0 and up.nstride == 0

    def conv(self, name: str, src: V, dst: V, weight: torch.Tensor, bias: torch.Tensor, stride: int, act: str,
             res: Optional[V] = None, up: Optional[V] = None) -> V:
        """up: conv over torch.cat([upsample2x(up), src], channel) without materialising either (1x1 only)."""
        cout, cin, k, k2 = weight.shape
        assert k == k2 and (k in (1, 3) or (k == 4 and stride == 2)), f"{name}: unsupported kernel size {k}"
        if up is not None:
            assert k == 1 and stride == 1 and self.can_fuse_upsample(up), f"{name}: cannot fuse this upsample"
            assert (2 * up.H, 2 * up.W) == (src.H, src.W) and cin == up.c + src.c, f"{name}: upsample/concat shape mismatch"
            assert cout <= dst.c <= _rup(cout, 16) and (dst.H, dst.W) == (src.H, src.W)
            return self._add(PlannedOp(_capi.OP_CONV, name, src, dst, res, k, stride, _capi.act_code(act),
                                       weight.detach().float().cpu(), bias.detach().float().cpu(), up=up,
                                       cin_pad=up.c + _rup(src.c, 16), cout_pad=_rup(cout, 16)))
        assert cin <= src.c <= _rup(cin, 16), f"{name}: src has {src.c} channels, weight expects {cin}"
        assert cout <= dst.c <= _rup(cout, 16), f"{name}: dst has {dst.c} channels, weight gives {cout}"
        pad = (k - 1) // 2
        ho, wo = (src.H + 2 * pad - k) // stride + 1, (src.W + 2 * pad - k) // stride + 1
        assert (dst.H, dst.W) == (ho, wo), f"{name}: dst is {dst.H}x{dst.W}, conv gives {ho}x{wo}"
        if res is not None:
            assert (res.H, res.W, res.c) == (dst.H, dst.W, dst.c), f"{name}: residual shape mismatch"
        return self._add(PlannedOp(_capi.OP_CONV, name, src, dst, res, k, stride, _capi.act_code(act),
                                   weight.detach().float().cpu(), bias.detach().float().cpu(),
                                   cin_pad=_rup(cin, 16), cout_pad=_rup(cout, 16)))

    def dwconv(self, name: str, src: V, dst: V, weight: torch.Tensor, bias: torch.Tensor, stride: int, act: str) -> V:
        c, one, k, _ = weight.shape
        assert one == 1 and c == src.c == dst.c and k in (3, 5)
        return self._add(PlannedOp(_capi.OP_DWCONV, name, src, dst, None, k, stride, _capi.act_code(act),
                                   weight.detach().float().cpu(), bias.detach().float().cpu(), cin_pad=c, cout_pad=c))

    def spp(self, src: V, dst: V) -> V:
        assert dst.c == 3 * src.c
        return self._add(PlannedOp(_capi.OP_SPP, "spp", src, dst))

    def upsample(self, src: V, dst: V) -> V:
        assert (dst.H, dst.W, dst.c) == (2 * src.H, 2 * src.W, src.c)
        return self._add(PlannedOp(_capi.OP_UPSAMPLE, "upsample", src, dst))

    # ---- finalisation --------------------------------------------------------------------
    def place_buffers(self):
        """Greedy-by-size first-fit over [first,last] live ranges."""
        n_ops = len(self.ops)
        # Small batches run the op list over several streams (csrc/yx_engine.cu "lanes": the head's pyramid levels, its cls /
        # reg branches and the bottom-up path side by side), ordered by the arena ranges every op reads and writes.  Reusing a
        # dead buffer's memory would add write-after-read / write-after-write edges between otherwise independent ops, so every
        # buffer keeps its own range there (all of YOLOX-M-P6 1280x1280 is 0.55 GB per image).
        lanes = os.environ.get("YX_LANES")
        no_reuse = self.batch <= int(os.environ.get("YX_LANES_MAX_BATCH", "8")) and (lanes is None or lanes != "0")
        for b in self.bufs:
            assert b.last >= 0, f"buffer {b.name} is never used"
            if b.pinned:
                b.last = n_ops
            if no_reuse:
                b.first, b.last = 0, n_ops
        placed: List[Buf] = []
        for b in sorted(self.bufs, key=lambda x: -x.nbytes):
            busy = sorted((p.offset, p.offset + p.nbytes) for p in placed
                          if not (p.last < b.first or b.last < p.first))
            off = 0
            for lo, hi in busy:
                if off + b.nbytes <= lo:
                    break
                off = max(off, hi)
            b.offset = off
            placed.append(b)
        self.arena_bytes = max(b.offset + b.nbytes for b in self.bufs)
        return self.arena_bytes

    def pack_weights(self):
        """fp16 KRSC blob [cout_pad][k*k][cin_pad] per conv (depthwise: [k*k][c]); fp32 bias blob."""
        w_parts, b_parts, w_off, b_off = [], [], 0, 0
        for op in self.ops:
            if op.kind == _capi.OP_CONV:
                cout, cin, kh, kw = op.weight.shape
                w = torch.zeros(op.cout_pad, kh * kw, op.cin_pad, dtype=torch.float16)
                w[:cout, :, :cin] = op.weight.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin).to(torch.float16)
            elif op.kind == _capi.OP_DWCONV:
                c, _, k, _ = op.weight.shape
                cout = c
                w = op.weight.reshape(c, k * k).t().contiguous().to(torch.float16)
            else:
                continue
            b = torch.zeros(op.cout_pad, dtype=torch.float32)
            b[:cout] = op.bias
            op.w_offset, op.b_offset = w_off, b_off
            w_parts.append(w.reshape(-1)); b_parts.append(b)
            nw = _rup(w.numel() * 2, 256) // 2
            if nw > w.numel():
                w_parts.append(torch.zeros(nw - w.numel(), dtype=torch.float16))
            nb = _rup(b.numel() * 4, 256) // 4
            if nb > b.numel():
                b_parts.append(torch.zeros(nb - b.numel(), dtype=torch.float32))
            w_off += nw * 2; b_off += nb * 4
        self.weight_blob = torch.cat(w_parts)
        self.bias_blob = torch.cat(b_parts)

    def c_ops(self):
        arr = (_capi.Op * len(self.ops))()
        for i, op in enumerate(self.ops):
            o = arr[i]
            o.kind, o.ksize, o.stride, o.act, o.aux = op.kind, op.ksize, op.stride, op.act, op.aux
            if op.src is not None:
                o.src = op.src.to_c()
            o.dst = op.dst.to_c()
            if op.res is not None:
                o.res = op.res.to_c()
            if op.up is not None:
                o.up = op.up.to_c()
            o.w_offset, o.b_offset, o.cin_pad, o.cout_pad = op.w_offset, op.b_offset, op.cin_pad, op.cout_pad
        return arr

    def finalize(self):
        self.place_buffers()
        self.pack_weights()
        return self

    # algorithmic work: conv flops / fp16 activation+weight bytes
    def conv_flops(self) -> float:
        t = 0.0
        for op in self.ops:
            if op.kind == _capi.OP_CONV and (op.aux & 1):
                t += 2.0 * self.batch * op.dst.H * op.dst.W * op.weight.shape[0] * 12 * 9
            elif op.kind == _capi.OP_CONV:
                cout, cin, k, _ = op.weight.shape
                t += 2.0 * self.batch * op.dst.H * op.dst.W * cout * cin * k * k
            elif op.kind == _capi.OP_DWCONV:
                c, _, k, _ = op.weight.shape
                t += 2.0 * self.batch * op.dst.H * op.dst.W * c * k * k
        return t


class TuneCache:
    """Persists the launch shapes yx_engine_tune picked, so that they -- and with them the fp32 summation order and the
    low bits of every result -- are the same from run to run (the counterpart of cuDNN's benchmark cache, which the
    reference re-fills on every start, tools/eval.py:122).

    One JSON file: {"<device name>|<kernel revision>": {"<layer geometry incl. batch>": [yx_conv_tune fields]}}.
    Location: $YX_TUNE_CACHE, else tune_cache.json next to this file (shipped with the shapes measured on this pool's
    B200s); YX_TUNE_CACHE=0 disables it, YX_TUNE=force re-tunes and overwrites, YX_TUNE=0 keeps the heuristic shapes.
    An engine uses the cache only when EVERY conv of it has an entry that still plans; otherwise it tunes and the file
    is rewritten atomically."""

    def __init__(self):
        env = os.environ.get("YX_TUNE_CACHE", "")
        self.enabled = env != "0"
        self.path = env if env not in ("", "0", "1") else os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_cache.json")

    @staticmethod
    def op_key(op: "PlannedOp", batch: int, sparse_ok: bool = False) -> str:
        def v(x):
            return "-" if x is None else f"{x.H}x{x.W}x{x.c}p{x.buf.c}" + (f"@{x.lvl_off}" if x.lvl_off or x.nstride else "")
        inplace = op.res is not None and op.res.buf is op.dst.buf and op.res.c_off == op.dst.c_off
        # a layer with 2:4-compliant weights has the sparse tensor-core shapes among its candidates (and YX_SPARSE changes which)
        sp = f" sp24:{os.environ.get('YX_SPARSE', 'auto')}" if sparse_ok else ""
        return (f"b{batch} k{op.ksize} s{op.stride} act{op.act} aux{op.aux} cin{op.cin_pad} cout{op.cout_pad} "
                f"src{v(op.src)} dst{v(op.dst)} res{'inplace' if inplace else v(op.res)} up{v(op.up)}{sp}")

    def section(self, device) -> str:
        import torch
        from . import _build
        return f"{torch.cuda.get_device_name(device)}|{_build.kernel_rev()}"

    def load(self, device) -> Dict[str, list]:
        import json
        if not self.enabled or not os.path.exists(self.path):
            return {}
        try:
            with open(self.path) as f:
                return json.load(f).get(self.section(device), {})
        except (OSError, ValueError):
            return {}

    def store(self, device, entries: Dict[str, list]):
        import fcntl
        import json
        if not self.enabled or os.environ.get("YX_TUNE_CACHE_WRITE", "1") == "0":
            return
        try:
            with open(self.path + ".lock", "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                data = {}
                if os.path.exists(self.path):
                    try:
                        with open(self.path) as f:
                            data = json.load(f)
                    except ValueError:
                        data = {}
                sec = self.section(device)
                data = {sec: data.get(sec, {})}          # shapes of other kernel revisions are dead weight
                data[sec].update(entries)
                tmp = f"{self.path}.{os.getpid()}.tmp"
                with open(tmp, "w") as f:
                    json.dump(data, f, indent=0, sort_keys=True)
                os.replace(tmp, self.path)
        except OSError:
            pass                                          # a read-only tree: the shapes simply are not persisted


class Engine:
    """Owns the device arena / weight blobs and the native engine handle for one (B, H, W)."""

    def __init__(self, graph: Graph, device):
        import torch
        self.lib = _capi.load()
        self.graph = graph
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("yolox_b200: the engine runs on CUDA (sm_100a) only; there is no CPU path")
        with torch.cuda.device(self.device):
            self.arena = torch.empty(graph.arena_bytes + ALIGN, dtype=torch.uint8, device=self.device)
            pad = (-self.arena.data_ptr()) % ALIGN
            self.arena_base = self.arena.data_ptr() + pad
            self._arena_pad = pad
            self.weights = graph.weight_blob.to(self.device)
            self.biases = graph.bias_blob.to(self.device)
            ops = graph.c_ops()
            handle = _capi.c_vp()
            _capi.check(self.lib.yx_engine_create(ops, len(graph.ops), self.arena_base, graph.arena_bytes,
                                                  self.weights.data_ptr(), self.weights.numel() * 2,
                                                  self.biases.data_ptr(), self.biases.numel() * 4,
                                                  graph.in_h, graph.in_w, graph.batch, handle), "yx_engine_create")
            self.handle = handle
        self.n_launches = self.lib.yx_engine_num_launches(self.handle)
        # per-layer launch shapes are picked by measurement on the first real input (like cudnn.benchmark);
        # YX_TUNE=0 keeps the heuristic shapes
        self.tuned = os.environ.get("YX_TUNE", "1") == "0"
        self.shape_source = "heuristic"     # -> "cache" (persisted shapes loaded) or "tuned" (measured in this process)

    def tensor_of(self, buf: Buf):
        """fp16 torch view [n, h, w, c] of an arena buffer (no copy)."""
        import torch
        n = buf.n * buf.h * buf.w * buf.c
        start = self._arena_pad + buf.offset
        return self.arena[start:start + 2 * n].view(torch.float16).view(buf.n, buf.h, buf.w, buf.c)

    def run(self, image, in_scale: float = 1.0, in_shift: float = 0.0, use_graph: bool = False):
        import torch
        g = self.graph
        _capi.require_cuda(image, "image")
        if tuple(image.shape) != (g.batch, 3, g.in_h, g.in_w):
            raise RuntimeError(f"engine built for {(g.batch, 3, g.in_h, g.in_w)}, got {tuple(image.shape)}")
        dt = _capi.image_dtype(image)
        image = image.contiguous()
        if not self.tuned:
            self.tuned = True
            if not self._shapes_from_cache():
                _capi.check(self.lib.yx_engine_tune(self.handle, image.data_ptr(), dt, float(in_scale), float(in_shift),
                                                    int(os.environ.get("YX_TUNE_ITERS", "3")), _capi.current_stream_ptr()),
                            "yx_engine_tune")
                self.shape_source = "tuned"
                self._shapes_to_cache()
        _capi.check(self.lib.yx_engine_run(self.handle, image.data_ptr(), dt, float(in_scale), float(in_shift),
                                           int(use_graph), _capi.current_stream_ptr()), "yx_engine_run")

    # ---- persisted launch shapes (TuneCache) -----------------------------------------------------------------
    def _conv_ops(self):
        return [(i, op) for i, op in enumerate(self.graph.ops) if op.kind == _capi.OP_CONV]

    def _op_key(self, i, op) -> str:
        return TuneCache.op_key(op, self.graph.batch, bool(self.lib.yx_engine_op_sparse_ok(self.handle, i)))

    def _shapes_from_cache(self) -> bool:
        import ctypes
        if os.environ.get("YX_TUNE", "1") == "force" or os.environ.get("YX_TUNE_CHECK"):
            return False
        cache = TuneCache()
        entries = cache.load(self.device)
        convs = self._conv_ops()
        keys = [self._op_key(i, op) for i, op in convs]
        if not entries or any(k not in entries for k in keys):
            return False
        previous = []
        for (i, _), k in zip(convs, keys):
            t = _capi.ConvTune()
            self.lib.yx_engine_get_tune(self.handle, i, ctypes.byref(t))
            previous.append(t)
            new = _capi.ConvTune.from_list(entries[k])
            if self.lib.yx_engine_set_tune(self.handle, i, ctypes.byref(new)) != 0:
                for (j, _), old in zip(convs, previous):      # a stale entry: back to the heuristic shapes, then tune
                    self.lib.yx_engine_set_tune(self.handle, j, ctypes.byref(old))
                return False
        self.shape_source = "cache"
        return True

    def _shapes_to_cache(self):
        import ctypes
        entries = {}
        for i, op in self._conv_ops():
            t = _capi.ConvTune()
            _capi.check(self.lib.yx_engine_get_tune(self.handle, i, ctypes.byref(t)), "yx_engine_get_tune")
            entries[self._op_key(i, op)] = t.as_list()
        TuneCache().store(self.device, entries)

    def run_ops(self, image, first: int, count: int, in_scale: float = 1.0, in_shift: float = 0.0):
        """Diagnostic: run ops [first, first+count) only."""
        import torch
        dt = _capi.image_dtype(image)
        _capi.check(self.lib.yx_engine_run_ops(self.handle, image.contiguous().data_ptr(), dt, float(in_scale),
                                               float(in_shift), first, count, _capi.current_stream_ptr()),
                    "yx_engine_run_ops")

    def view_tensor(self, cview):
        """Strided fp16 torch view [n,h,w,c] of a C view (no copy)."""
        import torch
        a16 = self.arena[self._arena_pad:self._arena_pad + (self.graph.arena_bytes // 2) * 2].view(torch.float16)
        # as_strided's storage_offset is absolute in the storage: add the slice's own offset (the alignment pad)
        return torch.as_strided(a16, (cview.n, cview.h, cview.w, cview.c),
                                (cview.nstride, cview.w * cview.pitch, cview.pitch, 1),
                                a16.storage_offset() + cview.offset // 2)

    def profile(self, image, iters: int = 5):
        import ctypes
        import torch
        n = len(self.graph.ops)
        ms = (ctypes.c_float * n)(); fl = (ctypes.c_double * n)(); by = (ctypes.c_double * n)()
        dt = _capi.image_dtype(image)
        _capi.check(self.lib.yx_engine_profile(self.handle, image.contiguous().data_ptr(), dt, iters,
                                               _capi.current_stream_ptr(), ms, fl, by, n), "yx_engine_profile")
        return [dict(name=op.name, kind=op.kind, ms=ms[i], flops=fl[i], bytes=by[i], shape=self.op_desc(i))
                for i, op in enumerate(self.graph.ops)]

    def tune_mismatches(self):
        """(YX_TUNE_CHECK=1) descriptions of tuning candidates whose output differed from the default launch shape."""
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        n = self.lib.yx_engine_tune_mismatches(self.handle, buf, 1 << 16)
        return [m for m in buf.value.decode().split("\n") if m] if n > 0 else []

    def op_desc(self, i: int) -> str:
        import ctypes
        buf = ctypes.create_string_buffer(320)
        _capi.check(self.lib.yx_engine_op_desc(self.handle, i, buf, 320), "yx_engine_op_desc")
        return buf.value.decode()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.yx_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
