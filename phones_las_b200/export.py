"""Export for serving: the reference's ``export.py`` on this path's own runtime.

``export.py:57-79`` builds an Estimator around ``export_las_model_fn`` (listener + phone speller only, PREDICT) and calls
``export_saved_model(export_dir, serving_input_receiver_fn)``: a time-stamped directory holding the graph (``saved_model.pb``),
``variables/variables.{index,data-*}`` and a serving signature with inputs ``encoder_inputs float32 [None, None, C]``,
``source_sequence_length int32 [None]`` and outputs ``sample_ids``, ``alignment``, ``probs`` (``export.py:39-65``).

Here the graph is the CUDA library, so the exported directory holds what is left: the variables as a TF V2 bundle under the
same ``variables/variables`` prefix (tf_checkpoint.py), ``hparams.json`` in the reference's double-JSON format, and
``signature.json`` stating the serving signature.  ``ServingModel`` is the loader: ``predict`` is ``export_las_model_fn``.
"""
import json
import os
import time

import numpy as np

SIGNATURE_NAME = "serving_default"
METHOD_NAME = "tensorflow/serving/predict"


def serving_signature(num_channels, hp):
    """The signature export.py:57-65 registers, as data."""
    outputs = {"sample_ids": {"dtype": "int32", "shape": [None, None] + ([None] if int(hp.get("beam_width", 0) or 0) > 0 else [])},
               "alignment": {"dtype": "float32", "shape": [None, None, None]}}
    if int(hp.get("beam_width", 0) or 0) == 0:  # export.py:48-52
        outputs["probs"] = {"dtype": "float32", "shape": [None, None, hp["target_vocab_size"]],
                            "activation": "sigmoid" if hp.get("binary_outputs") else "softmax"}
    return {SIGNATURE_NAME: {"method_name": METHOD_NAME,
                             "inputs": {"encoder_inputs": {"dtype": "float32", "shape": [None, None, int(num_channels)]},
                                        "source_sequence_length": {"dtype": "int32", "shape": [None]}},
                             "outputs": outputs}}


def export_saved_model(model_dir, export_dir, num_channels, timestamp=None):
    """``python export.py --model_dir M --num_channels C --export_dir E``: reads ``M/hparams.json`` and the latest checkpoint
    of ``M``, writes ``E/<timestamp>/`` and returns that path.  Only the variables export_las_model_fn touches are kept
    (scopes ``listener/`` and ``speller/``; optimizer slots, counters and the ``speller_binf`` / ``ctc_logits`` heads are not
    part of the exported graph)."""
    from . import hparams as hps, tf_checkpoint
    hp = hps.load_hparams(model_dir)
    variables = tf_checkpoint.load_model_variables(model_dir)
    kept = {k: v for k, v in variables.items() if k.startswith(("listener/", "speller/"))}
    if not kept:
        raise ValueError(f"{model_dir}: the checkpoint holds no listener/ or speller/ variables")
    stamp = str(int(time.time()) if timestamp is None else timestamp)
    out = os.path.join(export_dir, stamp)
    os.makedirs(os.path.join(out, "variables"), exist_ok=True)
    tf_checkpoint.write_checkpoint(os.path.join(out, "variables", "variables"), kept)
    hp = dict(hp, num_channels=int(num_channels))
    hps.save_hparams(hp, out)
    with open(os.path.join(out, "signature.json"), "w") as f:
        json.dump(serving_signature(num_channels, hp), f, indent=1)
    return out


class ServingModel:
    """Loader of an exported directory; ``predict`` = export_las_model_fn (export.py:8-54)."""

    def __init__(self, export_path, precision="fp32", device="cuda"):
        from . import hparams as hps, tf_checkpoint
        from .model import DeviceWeights
        self.hp = hps.load_hparams(export_path)
        with open(os.path.join(export_path, "signature.json")) as f:
            self.signature = json.load(f)[SIGNATURE_NAME]
        self.num_channels = self.signature["inputs"]["encoder_inputs"]["shape"][-1]
        variables = tf_checkpoint.read_checkpoint(os.path.join(export_path, "variables", "variables"), skip_unsupported=True)
        self.variables = {k: np.asarray(v, np.float32) for k, v in variables.items()}
        hp = dict(self.hp, binary_outputs=False, multitask=False, ctc_weight=-1.0)  # the exported graph is listener + phone speller
        self.weights = DeviceWeights(self.variables, hp, self.num_channels, precision, device)
        self._hp_run, self.device = hp, device

    def predict(self, features):
        """features: {'encoder_inputs' [B, T, C] float32, 'source_sequence_length' [B] int32} (host or device tensors / arrays)
        -> {'sample_ids', 'alignment', 'probs'} device tensors, the keys of the serving signature."""
        import torch
        from .model import las_predict
        x = torch.as_tensor(features["encoder_inputs"], dtype=torch.float32).to(self.device)
        n = torch.as_tensor(features["source_sequence_length"], dtype=torch.int32).to(self.device)
        if x.dim() != 3 or x.shape[-1] != self.num_channels:
            raise ValueError(f"encoder_inputs must be [batch, time, {self.num_channels}], got {tuple(x.shape)}")
        pred = las_predict({"encoder_inputs": x, "source_sequence_length": n}, self._hp_run, self.weights)
        out = {"sample_ids": pred["sample_ids"]}
        if pred.get("alignment") is not None:
            out["alignment"] = pred["alignment"]
        if "probs" in pred:
            out["probs"] = pred["probs"]
        return out
