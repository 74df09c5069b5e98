"""Listener-only timing of the training recurrences at c3 shapes: cluster (default) vs L2 exchange, B = 32 and 64."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    sys.path.insert(0, ".")
    from phones_las_b200 import _lib, synth, weights, train as tr
    from phones_las_b200.hparams import create_hparams
    B, T, C = int(os.environ.get("PB", 32)), 297, 39
    hp = create_hparams(target_vocab_size=64, encoder_layers=3, encoder_units=256, decoder_units=256, decoder_layers=1, num_channels=C,
                        dropout=0.0, sampling_probability=0.0)
    params = {k: v for k, v in weights.init_params(hp, seed=1).items() if k.startswith("listener/")}
    st = tr.TrainState(params)
    x, lens = synth.synth_features(B, T, C)
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda()
    for it in range(3):
        if it == 2:
            _lib.timeline_start()
        out, ol, tape = tr.listener_train_fwd(xd, ld, st, hp)
        tr.listener_train_bwd(torch.ones_like(out), tape, st, hp)
    tl = _lib.timeline_stop()
    print("B", B, "exchange", os.environ.get("PLAS_RT_EXCHANGE", "cluster"), {k: round(sum(v), 3) for k, v in tl.items() if "rec" in k}, flush=True)
else:
    for pb in ("32", "64"):
        for mode in (None, "l2"):
            env = dict(os.environ, PB=pb)
            if mode:
                env["PLAS_RT_EXCHANGE"] = mode
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
