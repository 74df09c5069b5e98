"""Checkpoint variable layout of the reference (TF names -> shapes) and synthetic initialisers.

Variable names and shapes follow the scopes the reference opens: ``listener/`` model_helper.py:206,
``bilstm_{l}`` las/ops.py:76, ``speller/`` model_helper.py:213, ``ctc_logits`` model_helper.py:349-350,
``projection_layer`` las/model.py:251-257, the LSTMCell ``kernel``/``bias`` pair created by
tf.nn.rnn_cell.LSTMCell at las/ops.py:11-12.  Weights are exchanged as a flat
``{tf_variable_name: float32 ndarray}`` dict (an ``.npz`` with the TF names), so reference weights
load unchanged; device-side re-layout happens in listener.py / speller.py.
"""
import numpy as np


def encoder_output_depth(hp):
    U = hp["encoder_units"]
    nd = 1 if hp["unidirectional"] else 2
    if hp["use_pyramidal"]:
        return nd * U * (2 if hp["encoder_layers"] > 1 else 1)
    return nd * U


def variable_shapes(hp, num_channels=None):
    """Ordered {name: shape} for the default wiring (bottom_only=False, attention_layer_size=None)."""
    C = num_channels or hp["num_channels"]
    U, L = hp["encoder_units"], hp["encoder_layers"]
    Ud, Ld, V = hp["decoder_units"], hp["decoder_layers"], hp["target_vocab_size"]
    nd = 1 if hp["unidirectional"] else 2
    shapes = {}
    if hp["use_pyramidal"]:
        din = C
        for l in range(L):
            dirs = ["rnn"] if hp["unidirectional"] else ["bidirectional_rnn/fw", "bidirectional_rnn/bw"]
            for d in dirs:
                shapes[f"listener/bilstm_{l}/{d}/lstm_cell/kernel"] = (din + U, 4 * U)
                shapes[f"listener/bilstm_{l}/{d}/lstm_cell/bias"] = (4 * U,)
            din = nd * U * (1 if l == 0 else 2)
    else:
        dirs = ["rnn"] if hp["unidirectional"] else ["bidirectional_rnn/fw", "bidirectional_rnn/bw"]
        for d in dirs:
            din = C
            for l in range(L):
                shapes[f"listener/{d}/multi_rnn_cell/cell_{l}/lstm_cell/kernel"] = (din + U, 4 * U)
                shapes[f"listener/{d}/multi_rnn_cell/cell_{l}/lstm_cell/bias"] = (4 * U,)
                din = U
    D = encoder_output_depth(hp)
    A = hp.get("attention_layer_size") or D  # attention_layer_size=None -> attention = context (depth D)
    shapes["speller/memory_layer/kernel"] = (D, Ud)
    Ein = int(hp.get("embedding_size") or 0) or V  # width of the decoder inputs: embedding lookup or one-hot (las/model.py:228-246)
    if hp.get("embedding_size"):
        shapes["speller/target_embedding"] = (V, Ein)
    bottom = bool(hp.get("bottom_only"))
    pre = "speller/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper" if bottom else "speller/decoder/attention_wrapper"
    for k in range(Ld):
        if bottom:  # AttentionMultiCell (las/model.py:20-69): cell 0 under the attention wrapper, upper cells read [prev; old attention]
            name = f"{pre}/lstm_cell" if k == 0 else f"speller/decoder/multi_rnn_cell/cell_{k}/lstm_cell"
            din = (Ein + A) if k == 0 else ((A if k == 1 else Ud) + A)
        else:
            name = f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell"
            din = (Ein + A) if k == 0 else Ud
        shapes[name + "/kernel"] = (din + Ud, 4 * Ud)
        shapes[name + "/bias"] = (4 * Ud,)
    at = hp["attention_type"]
    if at == "bahdanau":
        shapes[f"{pre}/bahdanau_attention/query_layer/kernel"] = (Ud, Ud)
        shapes[f"{pre}/bahdanau_attention/attention_v"] = (Ud,)
    elif at == "luong_monotonic":
        shapes[f"{pre}/luong_monotonic_attention/attention_score_bias"] = ()
    elif at == "bahdanau_monotonic":  # scopes [3P-recalled] like the others
        shapes[f"{pre}/bahdanau_monotonic_attention/query_layer/kernel"] = (Ud, Ud)
        shapes[f"{pre}/bahdanau_monotonic_attention/attention_v"] = (Ud,)
        shapes[f"{pre}/bahdanau_monotonic_attention/attention_score_bias"] = ()
    elif at == "custom":              # CustomAttention's own Dense 'query_layer' (las/model.py:92-93)
        shapes[f"{pre}/query_layer/kernel"] = (Ud, Ud)
    if hp.get("attention_layer_size"):  # AttentionWrapper's Dense over [cell output; context], no bias (las/model.py:180-200)
        shapes[f"{pre}/attention_layer/kernel"] = (Ud + D, A)
    if bottom and Ld > 1:
        A = Ud  # the projection reads the top cell's output
    shapes["speller/decoder/projection_layer/kernel"] = (A, V)
    shapes["speller/decoder/projection_layer/bias"] = (V,)
    if hp.get("ctc_weight", -1) > 0:
        shapes["ctc_logits/kernel"] = (D, V + 1)
        shapes["ctc_logits/bias"] = (V + 1,)
    return shapes


def init_params(hp, num_channels=None, seed=4321, projection_scale=1.0, bias_scale=0.0, shapes=None):
    """Synthetic weights: LSTM kernels and the projection U(-0.075, 0.075)
    (las/ops.py:12, las/model.py:257), LSTM biases 0 (or U(-bias_scale, bias_scale) to exercise the
    bias path), Dense kernels / attention_v Glorot-uniform, attention_score_bias 0, a trainable binf2phone U(0, 1)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape in (shapes or variable_shapes(hp, num_channels)).items():
        if name.endswith("lstm_cell/kernel") or name.endswith("projection_layer/kernel"):
            w = rng.uniform(-0.075, 0.075, size=shape)
            if name.endswith("projection_layer/kernel"):
                w = w * projection_scale
        elif name.endswith("/bias"):
            w = rng.uniform(-bias_scale, bias_scale, size=shape) if bias_scale else np.zeros(shape)
        elif name.endswith("attention_score_bias"):
            w = np.zeros(shape)
        elif name == "binf2phone":  # --binf_trainable: tf.random_uniform_initializer(0, 1), model_helper.py:181-184
            w = rng.uniform(0.0, 1.0, size=shape)
        elif name.endswith("attention_v"):
            lim = np.sqrt(6.0 / (shape[0] + 1))
            w = rng.uniform(-lim, lim, size=shape)
        else:  # Dense kernels: glorot uniform
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            w = rng.uniform(-lim, lim, size=shape)
        params[name] = np.asarray(w, dtype=np.float32)
    return params


def save_npz(path, params):
    np.savez(path, **{k.replace("/", "|"): v for k, v in params.items()})


def load_npz(path):
    with np.load(path) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


def count_params(params):
    return int(sum(int(np.prod(v.shape)) for v in params.values()))
