"""Phase timeline (cycles, thread 0 of image 0) of sort_keys_kernel / nms_kernel; library built with YX_NVCC_DEFS=-DYX_POST_DBG."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolox_b200 import _capi
from yolox_b200 import postprocess as pp
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _capi.load()
names = {0: "sort start", 1: "pass0", 2: "pass1", 3: "pass2", 4: "pass3", 5: "pass4", 6: "pass5", 7: "pass6", 8: "pass7", 9: "compact", 10: "bitonic", 11: "writeback",
         16: "nms start", 17: "gather", 18: "trick", 19: "greedy", 20: "write", 23: "blk0 end", 24: "blk1 loaded", 25: "blk1 tested", 26: "blk1 sync", 27: "blk1 resolved", 28: "blk1 end"}
def show(tag):
    buf = (ctypes.c_longlong * 64)()
    lib.yx_post_dbg_read(buf)
    v = list(buf)
    for a in (0, 16):
        t0 = v[a]
        print(tag, " | ".join(f"{names[i]} {v[i] - t0}" for i in sorted(names) if a <= i < a + 16 and v[i]), flush=True)
strides, S = (8, 16, 32, 64), 1280
hw = [(S // s, S // s) for s in strides]
A, C = sum(h * w for h, w in hw), 80
g = torch.Generator(device=dev).manual_seed(4)
reg = torch.randn(B, A, 4, device=dev, generator=g).half()
obj = (torch.randn(B, A, 1, device=dev, generator=g) * 2 - 2).half()
cls = (torch.randn(B, A, C, device=dev, generator=g) * 2 - 2).half()
for _ in range(3):
    pp.detect_main(reg, obj, cls, hw, strides, 0.001, 0.65, 5000, 300)
show("max_candidate")
bench.post_stress_lines(B, 3, dev, bench.measured_peaks())   # its last distribution = clustered
show("clustered")
