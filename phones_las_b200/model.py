"""End-to-end hot path: waveform -> features -> listener -> speller -> predictions dict.

``las_predict`` mirrors ``las_model_fn(mode=PREDICT)`` (reference model_helper.py:165-297): same
``features`` dict in (``encoder_inputs``, ``source_sequence_length``), same ``predictions`` keys
out (``encoder_out, source_length, embedding, sample_ids, alignment, probs``).  ``LASModel``
bundles a front-end plan with the device weights and follows the call order of
``transcribe_audio_file.py:97-101``.
"""
import numpy as np
import torch

from . import _lib, weights as wts
from .frontend import FrontendPlan
from .hparams import num_feature_channels
from .listener import ListenerWeights, listener
from .speller import SpellerWeights, speller


class DeviceWeights:
    def __init__(self, params, hp, num_channels=None, precision="fp32", device="cuda"):
        C = num_channels or hp["num_channels"]
        self.precision = precision
        self.listener = ListenerWeights(params, hp, C, precision, device)
        self.speller = SpellerWeights(params, hp, wts.encoder_output_depth(hp), precision, device)


def las_predict(features, hp, weights, want_alignment=True):
    """model_helper.py:165-297 in PREDICT mode with beam_width == 0."""
    x = features["encoder_inputs"]
    lens = features["source_sequence_length"]
    (enc_out, enc_len), enc_state = listener(x, lens, "infer", hp, weights.listener)
    out, state, final_len = speller(enc_out, enc_state, None, enc_len, None, "infer", hp, weights.speller)
    logits = out.rnn_output
    pred = {"encoder_out": enc_out, "source_length": enc_len, "sample_ids": out.sample_id,
            "logits": logits, "final_sequence_length": final_len, "alignment": state.alignment_history,
            "probs": torch.softmax(logits, dim=-1)}
    if isinstance(enc_state[0], tuple):
        emb_c = torch.cat([s[0] for s in enc_state], dim=1)
        emb_h = torch.cat([s[1] for s in enc_state], dim=1)
    else:
        emb_c, emb_h = enc_state
    pred["embedding"] = torch.stack([emb_c, emb_h], dim=1)
    return pred


class LASModel:
    """Front-end + listener + speller with device-resident weights."""

    def __init__(self, params, hp, feature_flags, precision="fp32", means=None, stds=None, device="cuda"):
        _lib.require_cuda()
        self.hp, self.fa, self.precision = hp, feature_flags, precision
        self.plan = FrontendPlan(feature_flags, means, stds, device)
        self.weights = DeviceWeights(params, hp, num_feature_channels(feature_flags), precision, device)

    def features(self, wave, n_samples=None):
        return self.plan(wave, n_samples)

    def predict_from_features(self, feats, n_frames):
        return las_predict({"encoder_inputs": feats, "source_sequence_length": n_frames}, self.hp, self.weights)

    def transcribe(self, wave, n_samples=None):
        """wave [B,N] float32 on the device -> predictions dict (transcribe_audio_file.py:97-101)."""
        feats, n_frames = self.plan(wave, n_samples)
        return self.predict_from_features(feats, n_frames)

    def transcribe_host(self, wave_host, n_samples_host=None):
        """Host entry point: pinned host waveform in, host numpy ids out (H2D + D2H inside)."""
        wave = wave_host.to("cuda", non_blocking=True)
        ns = n_samples_host.to("cuda", non_blocking=True) if n_samples_host is not None else None
        pred = self.transcribe(wave, ns)
        return pred["sample_ids"].cpu(), pred["final_sequence_length"].cpu()
