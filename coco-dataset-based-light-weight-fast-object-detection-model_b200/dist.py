"""Multi-GPU plumbing for the hot path: images are sharded by batch across ranks (one process per GPU,
weights replicated, NO collective on the forward pass) and the fixed-shape detections are collected with
ONE all-gather.  Replaces the reference's pickled Gloo object gather
(yolox/evaluators/coco_evaluator.py:127, yolox/utils/dist.py:224-265)."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of the global batch owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_detections(det: torch.Tensor, cnt: torch.Tensor) -> torch.Tensor:
    """det [B,R,7] fp32 + cnt [B] int32 -> one [B, R*7+1] fp32 wire tensor (count in the last column;
    exact for counts < 2^24)."""
    B = det.shape[0]
    return torch.cat([det.reshape(B, -1), cnt.reshape(B, 1).to(det.dtype)], dim=1).contiguous()


def unpack_detections(packed: torch.Tensor, rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n = packed.shape[0]
    return packed[:, :rows * 7].reshape(n, rows, 7), packed[:, rows * 7].round().to(torch.int32)


def all_gather_detections(det: torch.Tensor, cnt: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One collective: every rank receives the detections of the whole global batch, rank-major
    ([world*B, R, 7], [world*B]).  All ranks must pass the same B (pad the last shard)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det, cnt
    world = dist.get_world_size(group)
    packed = pack_detections(det, cnt)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    return unpack_detections(torch.cat(out, 0), det.shape[1])


def wrap_i32(v: int) -> int:
    """v as a wrapping signed 32-bit counter value (what the device's int arrival counters hold after v increments)."""
    return ((int(v) + 2 ** 31) % 2 ** 32) - 2 ** 31


class PeerGather:
    """Receive windows for the all-gather fused into the NMS kernel (include/yolox_b200.h, yx_detect_main_gather).

    Every rank owns ONE device buffer holding THREE windows (consecutive steps rotate) of det [world,B,rows,7] fp32 +
    cnt [world,B] int32, the arrival counters int32[world] and a status word.  The buffers are exported with CUDA IPC and
    the 64-byte handles exchanged once over the process group; from then on a step moves no data through
    torch.distributed: each rank's NMS kernel stores its rows into all windows over NVLink.

    Waiting is off the critical path (pipelined=True, the default): step s only waits -- before its selection kernels,
    i.e. after its own network has run -- for the rows of step s-1, which arrived long ago, so no rank is lock-stepped to
    the slowest one.  result() completes the LATEST step on demand (one tiny wait kernel on the current stream);
    result(lag=1) returns the previous step, already complete, with no wait at all.  Three windows make this safe: a
    window is overwritten three steps later, and a writer's step s+3 is ordered after its wait for step s+2, i.e. after
    every reader issued its step s+2, which follows that reader's use of step s (below).

    Validity: the views returned by result() may be read (on the stream the steps are issued on, or ordered after it)
    until the NEXT step is issued on this rank -- not longer."""

    WINDOWS = 3

    def __init__(self, batch: int, rows: int, device, group=None, timeout_ms: int = 60000, pipelined: bool = True):
        from . import _capi
        self._capi, self.lib = _capi, _capi.load()
        self.group, self.B, self.rows, self.timeout_ms = group, int(batch), int(rows), int(timeout_ms)
        self.pipelined = bool(pipelined)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _capi.MAX_PEERS:
            raise ValueError(f"PeerGather supports up to {_capi.MAX_PEERS} ranks on one node")
        self.device = torch.device(device)
        W, B = self.world, self.B
        self.det_bytes = W * B * self.rows * 7 * 4
        self.cnt_bytes = (W * B * 4 + 255) // 256 * 256
        self.win_bytes = self.det_bytes + self.cnt_bytes
        self.arrive_off = self.WINDOWS * self.win_bytes
        self.status_off = self.arrive_off + 256
        self.buf = torch.zeros(self.status_off + 256, dtype=torch.uint8, device=self.device)
        self.step = 0
        self._waited = 0          # steps [1.._waited] are known complete on the issuing stream
        # ---- exchange IPC handles, map the peers' buffers ---------------------------------------------------
        import ctypes
        handle = (ctypes.c_ubyte * _capi.IPC_HANDLE_BYTES)()
        off = ctypes.c_int64(0)
        self.ptrs = [0] * W
        self.ptrs[self.rank] = self.buf.data_ptr()
        self._opened = []
        if W > 1:
            # Every rank takes the same decision: a failure anywhere (export or open) is exchanged before anyone raises,
            # so a caller may fall back to the NCCL gather on all ranks together.
            mine, err = None, None
            with torch.cuda.device(self.device):
                try:
                    _capi.check(self.lib.yx_ipc_export(self.buf.data_ptr(), handle, ctypes.byref(off)), "yx_ipc_export")
                    mine = (bytes(handle), int(off.value))
                except RuntimeError as e:
                    err = f"rank {self.rank}: {e}"
                everyone = [None] * W
                dist.all_gather_object(everyone, mine, group=group)
                if err is None and all(x is not None for x in everyone):
                    try:
                        for r, (h, o) in enumerate(everyone):
                            if r == self.rank:
                                continue
                            base = ctypes.c_void_p()
                            hb = (ctypes.c_ubyte * _capi.IPC_HANDLE_BYTES).from_buffer_copy(h)
                            _capi.check(self.lib.yx_ipc_open(hb, ctypes.byref(base)), "yx_ipc_open")
                            self._opened.append(base.value)
                            self.ptrs[r] = base.value + o
                    except RuntimeError as e:
                        err = f"rank {self.rank}: {e}"
                elif err is None:
                    err = f"rank {self.rank}: a peer could not export its window"
                errors = [None] * W
                dist.all_gather_object(errors, err, group=group)
            if any(errors):
                self.close()
                raise RuntimeError("peer gather unavailable: " + "; ".join(e for e in errors if e))
            torch.cuda.synchronize(self.device)     # zero fill done before any peer may store into this buffer
            dist.barrier(group=group)

    def _slot(self, step: int) -> int:
        return (step - 1) % self.WINDOWS

    def _window(self, slot: int):
        w = self.buf[slot * self.win_bytes:(slot + 1) * self.win_bytes]
        det = w[:self.det_bytes].view(torch.float32).view(self.world * self.B, self.rows, 7)
        cnt = w[self.det_bytes:self.det_bytes + self.world * self.B * 4].view(torch.int32)
        return det, cnt

    def _cnt_ptr(self, slot: int) -> int:
        return self.buf.data_ptr() + slot * self.win_bytes + self.det_bytes

    def next_step(self, batch: int, rows: int):
        """yx_peer_out for the next step (pointers of this rank's block inside every rank's window)."""
        if batch != self.B or rows != self.rows:
            raise ValueError("PeerGather was sized for a different batch / row count")
        self.step += 1
        slot = self._slot(self.step)
        po = self._capi.PeerOut()
        po.world, po.timeout_ms = self.world, self.timeout_ms
        if self.pipelined:
            prev = self.step - 1        # completed before this step's selection kernels (free: it arrived long ago)
            if prev >= 1 and self._waited < prev:
                po.wait_mode, po.wait_target = self._capi.PEER_WAIT_BEFORE, wrap_i32(prev * self.B)
                po.wait_cnt = self._cnt_ptr(self._slot(prev))
                self._waited = prev
            else:
                po.wait_mode, po.wait_target = self._capi.PEER_WAIT_NONE, 0
        else:
            po.wait_mode, po.wait_target = self._capi.PEER_WAIT_AFTER, wrap_i32(self.step * self.B)
            po.wait_cnt = self._cnt_ptr(slot)
            self._waited = self.step
        blk_det, blk_cnt = self.B * self.rows * 7 * 4, self.B * 4
        for w in range(self.world):
            win = self.ptrs[w] + slot * self.win_bytes
            po.det[w] = win + self.rank * blk_det
            po.cnt[w] = win + self.det_bytes + self.rank * blk_cnt
            po.arrive[w] = self.ptrs[w] + self.arrive_off + 4 * self.rank
        po.local_arrive = self.buf.data_ptr() + self.arrive_off
        po.status = self.buf.data_ptr() + self.status_off
        return po

    def wait(self, step: Optional[int] = None):
        """Stream-ordered completion of `step` (default: the latest issued) on the current stream; no host sync."""
        step = self.step if step is None else int(step)
        if step < 1 or step <= self._waited:
            return
        with torch.cuda.device(self.device):
            self._capi.check(self.lib.yx_peer_wait(self.buf.data_ptr() + self.arrive_off, self.world, wrap_i32(step * self.B),
                                                   self.buf.data_ptr() + self.status_off, self.timeout_ms,
                                                   self._cnt_ptr(self._slot(step)), self.B,
                                                   self._capi.current_stream_ptr()), "yx_peer_wait")
        self._waited = step

    def result(self, lag: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """(det [world*B, rows, 7], cnt [world*B]) of step (latest - lag), rank-major: views of the window, valid until
        the next step is issued on this rank.  lag=0 issues the wait for the latest step; lag=1 (pipelined consumers)
        needs none: that step was completed by the latest step's own pre-wait."""
        step = self.step - int(lag)
        if step < 1:
            raise RuntimeError("PeerGather.result: no such step yet")
        if lag >= self.WINDOWS - 1:
            raise RuntimeError("PeerGather.result: that window has been handed back to the writers")
        self.wait(step)
        return self._window(self._slot(step))

    def status(self) -> int:
        """0 while healthy; 1 + rank of a peer whose rows did not arrive within the timeout (host sync)."""
        return int(self.buf[self.status_off:self.status_off + 4].view(torch.int32).item())

    def check(self):
        """Raises if a peer's rows ever missed the timeout (host sync: call it every N steps, not every step)."""
        st = self.status()
        if st != 0:
            raise RuntimeError(f"peer gather: rank {st - 1}'s detections did not arrive within {self.timeout_ms} ms "
                               "(its counts were zeroed in the affected window); the peer is slow or dead")

    def close(self):
        for base in self._opened:
            self.lib.yx_ipc_close(base)
        self._opened = []


def gathered_for_images(det_all, cnt_all, n_images: int, world: int, per_rank: int):
    """Drop the padding rows added so every rank ran the same batch size."""
    keep = []
    for r in range(world):
        s, e = shard_range(n_images, r, world)
        keep.extend(range(r * per_rank, r * per_rank + (e - s)))
    idx = torch.tensor(keep, device=det_all.device, dtype=torch.long)
    return det_all[idx], cnt_all[idx]
