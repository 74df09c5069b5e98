import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from phones_las_b200 import _lib, synth, weights
from phones_las_b200.model import LASModel
cfg = bench.workload("c2"); hp, fa = cfg["hp"], cfg["fa"]
params = weights.init_params(hp, cfg["C"], seed=4321)
model = LASModel(params, hp, fa, precision="bf16")
hw = [torch.from_numpy(synth.synth_audio(64, 15.0, seed=i)[0]).pin_memory() for i in range(3)]
dw = [h.cuda() for h in hw]
for i in range(3): model.transcribe(dw[i % 3])
torch.cuda.synchronize()
def timeit(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / n
print("device loop      ", timeit(lambda n: [model.transcribe(dw[i % 3]) for i in range(n)]))
print("transcribe_host  ", timeit(lambda n: [model.transcribe_host(hw[i % 3]) for i in range(n)]))
print("transcribe_stream", timeit(lambda n: list(model.transcribe_stream(hw[i % 3] for i in range(n)))))
# copy alone
s = torch.cuda.Stream(); buf = torch.empty_like(dw[0])
def cp(n):
    with torch.cuda.stream(s):
        for i in range(n): buf.copy_(hw[i % 3], non_blocking=True)
    s.synchronize()
print("H2D copy alone   ", timeit(cp))
# kernels with a concurrent copy loop
def both(n):
    with torch.cuda.stream(s):
        for i in range(n): buf.copy_(hw[i % 3], non_blocking=True)
    for i in range(n): model.transcribe(dw[i % 3])
print("device loop + concurrent copies", timeit(both))
_lib.timeline_start()
with torch.cuda.stream(s):
    for i in range(3): buf.copy_(hw[i % 3], non_blocking=True)
model.transcribe(dw[0])
print({k: round(sum(v), 3) for k, v in _lib.timeline_stop().items()})
# --- elimination: untrimmed async loop with per-step deferred read, no H2D
def async_loop(n, h2d=None):
    pend = None
    pin = [None, None]
    for i in range(n):
        if h2d == "main":
            x = hw[i % 3].to("cuda", non_blocking=True)
        else:
            x = dw[i % 3]
        pred = model.transcribe(x, want_alignment=False, trim=False, want_probs=False)
        sl = i & 1
        if pin[sl] is None:
            pin[sl] = (torch.empty(pred["sample_ids"].shape, dtype=torch.int32).pin_memory(), torch.empty((1,), dtype=torch.int32).pin_memory())
        pin[sl][0].copy_(pred["sample_ids"], non_blocking=True); pin[sl][1].copy_(pred["n_steps"], non_blocking=True)
        ev = torch.cuda.Event(); ev.record()
        if pend is not None:
            pend.synchronize()
        pend = ev
    pend.synchronize()
print("async deferred-read, no H2D     ", timeit(async_loop))
print("async deferred-read, H2D on main", timeit(lambda n: async_loop(n, "main")))
print("untrimmed no-read loop          ", timeit(lambda n: [model.transcribe(dw[i % 3], want_alignment=False, trim=False, want_probs=False) for i in range(n)]))
print("transcribe_stream again         ", timeit(lambda n: list(model.transcribe_stream(hw[i % 3] for i in range(n)))))
