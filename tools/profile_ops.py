"""Per-op CUDA-event profile of the bench model at a given batch/size -> JSON (see tools/make_profiles.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

torch.set_grad_enabled(False)
B, S, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
x = (torch.rand(B, 3, S, S, device=dev) * 255).half()
eng = model.engine_for(x)
for _ in range(3):
    eng.run(x)
torch.cuda.synchronize()
prof = eng.profile(x, iters=iters)
# whole-engine time with and without graph
def timed(use_graph, n=30):
    for _ in range(5):
        eng.run(x, use_graph=use_graph)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        eng.run(x, use_graph=use_graph)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
res = dict(batch=B, size=S, ops=prof, engine_ms_stream=timed(False), engine_ms_graph=timed(True),
           sum_ops_ms=sum(p["ms"] for p in prof))
json.dump(res, open(out, "w"), indent=1)
print(B, S, "sum ops", res["sum_ops_ms"], "stream", res["engine_ms_stream"], "graph", res["engine_ms_graph"])
