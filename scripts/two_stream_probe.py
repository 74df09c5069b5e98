"""Probe: consecutive batches on two alternating CUDA streams (kernels of batch i+1 may fill the SMs batch i's latency-bound
recurrence / decoder leave idle) vs one stream.  c2 shape, device-resident inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phones_las_b200 import synth, weights
from phones_las_b200.hparams import baseline_config, num_feature_channels
from phones_las_b200.model import LASModel

cfg = baseline_config("c2")
hp, fa = cfg["hp"], cfg["fa"]
C = num_feature_channels(fa)
model = LASModel(weights.init_params(hp, C, seed=4321), hp, fa, precision="bf16")
waves = [torch.from_numpy(synth.synth_audio(64, 15.0, seed=1234 + i)[0]).cuda() for i in range(4)]
K = 12


def run(n_streams):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for i in range(4):  # warm-up on every stream
        with torch.cuda.stream(streams[i % n_streams]):
            model.transcribe(waves[i % 4], want_alignment=False, trim=False, want_probs=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        with torch.cuda.stream(streams[i % n_streams]):
            model.transcribe(waves[i % 4], want_alignment=False, trim=False, want_probs=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3


for ng, rows in ((0, 0), (2, 16), (4, 16), (4, 11)):
    for k, v in (("PLAS_REC_NG", ng), ("PLAS_REC_ROWS", rows)):
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    for n in (1, 2, 3, 4):
        print(f"rec NG={ng or 'auto'} rows={rows or 'auto'}  {n} stream(s): {run(n):.3f} ms per batch", flush=True)
