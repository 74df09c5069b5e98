"""Probe: rounds of R consecutive batches -- the R encoders (front-end, projections, recurrences) run concurrently on R
streams with the recurrence confined to 2 clusters per batch, then the R decoders (which need the whole GPU) run back to back.
Compared with the plain alternating-stream pipeline.  c2 shape, device-resident inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phones_las_b200 import synth, weights, _lib
from phones_las_b200.hparams import baseline_config, num_feature_channels
from phones_las_b200.model import LASModel
from phones_las_b200.listener import listener
from phones_las_b200.speller import speller

cfg = baseline_config("c2")
hp, fa = cfg["hp"], cfg["fa"]
C = num_feature_channels(fa)
model = LASModel(weights.init_params(hp, C, seed=4321), hp, fa, precision="bf16")
W = model.weights
waves = [torch.from_numpy(synth.synth_audio(64, 15.0, seed=1234 + i)[0]).cuda() for i in range(4)]


def encode(i):
    feats, nf = model.plan(waves[i % 4], None)
    return listener(feats, nf, "infer", hp, W.listener)


def decode(enc):
    (enc_out, enc_len), enc_state = enc
    return speller(enc_out, enc_state, None, enc_len, None, "infer", hp, W.speller, memory_is_masked=True, want_alignment=False, trim=False)


def run_rounds(R, n_rounds, prio):
    lo, hi = -1, 0
    enc_streams = [torch.cuda.Stream(priority=0) for _ in range(R)]
    dec_stream = torch.cuda.Stream(priority=-1 if prio else 0)
    def go(n):
        prev_dec_done = None
        for r in range(n):
            encs, evs = [], []
            for k in range(R):
                with torch.cuda.stream(enc_streams[k]):
                    if prev_dec_done is not None and os.environ.get("ROUND_STRICT"):
                        enc_streams[k].wait_event(prev_dec_done)
                    encs.append(encode(r * R + k))
                    ev = torch.cuda.Event(); ev.record(); evs.append(ev)
            with torch.cuda.stream(dec_stream):
                for k in range(R):
                    dec_stream.wait_event(evs[k])
                    decode(encs[k])
                prev_dec_done = torch.cuda.Event(); prev_dec_done.record()
    go(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    go(n_rounds)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (n_rounds * R) * 1e3


os.environ["PLAS_REC_MAX_CLUSTERS"] = "2"
for ng, rows in ((4, 16),):
    for k, v in (("PLAS_REC_NG", ng), ("PLAS_REC_ROWS", rows)):
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    for R in (2, 3, 4):
        for prio in (0, 1):
            for strict in (0, 1):
                if strict: os.environ["ROUND_STRICT"] = "1"
                else: os.environ.pop("ROUND_STRICT", None)
                print(f"rec NG={ng or 'auto'} rows={rows or 'auto'} round of {R} prio={prio} strict={strict}: {run_rounds(R, 6, prio):.3f} ms per batch", flush=True)
