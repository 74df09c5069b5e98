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
    def __init__(self, params, hp, num_channels=None, precision="fp32", device="cuda", binf=None):
        """``binf`` (binf2phone [n, V]): with hp['binary_outputs'] and hp['binf_projection'] the 'speller_binf' scope is built
        too (model_helper.py:219-227); the phone speller then exists only under --multitask (model_helper.py:212-217)."""
        C = num_channels or hp["num_channels"]
        self.precision = precision
        self.listener = ListenerWeights(params, hp, C, precision, device)
        D = wts.encoder_output_depth(hp)
        self.speller = self.speller_binf = None
        if not hp.get("binary_outputs") or hp.get("multitask"):
            self.speller = SpellerWeights(params, hp, D, precision, device)
        if hp.get("binary_outputs") and hp.get("binf_projection"):
            if binf is None:
                raise ValueError("--binf_projection needs the binf2phone matrix (binf=)")
            self.speller_binf = SpellerWeights(params, hp, D, precision, device, scope="speller_binf", binf=binf)
            self.binf = torch.as_tensor(binf, dtype=torch.float32, device=device)
        elif hp.get("binary_outputs") and not hp.get("multitask"):
            # without --binf_projection the reference's own non-TRAIN graph of the binary-feature speller cannot be built
            # (transform_binf_to_phones slices [n:2n] out of an n-wide Dense output, utils/training_helper.py:19-21); under
            # --multitask the phone speller is served and the binary-feature one (a TRAIN-time auxiliary task) is skipped
            raise NotImplementedError("binary_outputs without --binf_projection has no inference graph (DESIGN.md section 8)")
        self.ctc = None
        if hp.get("ctc_weight", -1.0) > 0 and "ctc_logits/kernel" in params:  # the CTC head of EVAL mode (model_helper.py:347-363)
            up = lambda a: torch.as_tensor(np.asarray(a, np.float32), device=device).contiguous()
            self.ctc = (up(params["ctc_logits/kernel"]), up(params["ctc_logits/bias"]))


def las_predict(features, hp, weights, want_alignment=True, trim=True, want_probs=True):
    """model_helper.py:165-297 in PREDICT mode with beam_width == 0.  ``trim=False`` keeps the step axis at its
    static capacity and leaves the executed step count on the device (``n_steps``), so the call enqueues without a
    host synchronisation (serving loop)."""
    x = features["encoder_inputs"]
    lens = features["source_sequence_length"]
    (enc_out, enc_len), enc_state = listener(x, lens, "infer", hp, weights.listener)
    # the listener already zeroes its outputs past each (reduced) length: no separate masking pass
    pred = {"encoder_out": enc_out, "source_length": enc_len}
    logits = None
    if int(hp.get("beam_width", 0) or 0) > 0:
        # PREDICT with beam search (model_helper.py:231-237): sample_ids = predicted_ids [B, T, W] (beam 0 is the best hypothesis);
        # there are no logits / probs / alignment in this mode
        sp = weights.speller if weights.speller is not None else weights.speller_binf
        out, state, final_len = speller(enc_out, enc_state, None, enc_len, None, "infer", hp, sp,
                                        binf_embedding=getattr(weights, "binf", None) if sp is weights.speller_binf else None, memory_is_masked=True)
        key = "sample_ids" if weights.speller is not None else "sample_ids_phones_binf"
        pred.update({key: out.predicted_ids, "final_sequence_length": final_len, "n_steps": state.n_steps})
        want_probs = False
    elif weights.speller is not None:
        out, state, final_len = speller(enc_out, enc_state, None, enc_len, None, "infer", hp, weights.speller,
                                        memory_is_masked=True, want_alignment=want_alignment, trim=trim)
        logits = out.rnn_output
        pred.update({"sample_ids": out.sample_id, "logits": logits, "final_sequence_length": final_len,
                     "alignment": state.alignment_history, "n_steps": state.n_steps})
    if getattr(weights, "speller_binf", None) is not None and not int(hp.get("beam_width", 0) or 0):  # model_helper.py:219-227,241-251,272-275
        out_b, state_b, final_len_b = speller(enc_out, enc_state, None, enc_len, None, "infer", hp, weights.speller_binf,
                                              binf_embedding=weights.binf, memory_is_masked=True, want_alignment=want_alignment, trim=trim)
        pred.update({"logits_binf": out_b.rnn_output, "sample_ids_phones_binf": out_b.sample_id,
                     "final_sequence_length_binf": final_len_b, "alignment_binf": state_b.alignment_history})
        if logits is None:
            logits = out_b.rnn_output  # probs = softmax(logits_binf) when there is no phone speller (model_helper.py:295-296)
            pred["n_steps"] = state_b.n_steps
    if want_probs:
        pred["probs"] = torch.softmax(logits, dim=-1)
    emb = encoder_embedding(enc_state)
    if emb is not None:
        pred["embedding"] = emb
    return pred


def encoder_embedding(enc_state):
    """predictions['embedding'] (model_helper.py:258-268): stack(concat of the c's, concat of the h's) when ``encoder_state`` is a
    sequence of (c, h) pairs -- the pyramidal bidirectional (fw, bw) pair or the unidirectional MultiRNNCell stack --, the pair
    itself for the pyramidal unidirectional listener, and NO embedding for the stacked bidirectional listener, whose state is
    (tuple of fw layer states, tuple of bw layer states): there ``x.c`` raises and the reference falls through to None."""
    is_pair = lambda s: isinstance(s, (tuple, list)) and len(s) == 2 and all(torch.is_tensor(t) for t in s)
    if is_pair(enc_state):
        emb_c, emb_h = enc_state
    elif isinstance(enc_state, (tuple, list)) and len(enc_state) and all(is_pair(s) for s in enc_state):
        emb_c = torch.cat([s[0] for s in enc_state], dim=1)
        emb_h = torch.cat([s[1] for s in enc_state], dim=1)
    else:
        return None
    return torch.stack([emb_c, emb_h], dim=1)


def las_eval(features, labels, hp, weights):
    """model_helper.py:165-345 in EVAL mode (beam_width == 0; phone speller and / or the --binf_projection speller): greedy decode, the EVAL branch of
    ``compute_loss`` (logits cut at max(final_sequence_length), targets padded with eos, mask = elementwise max of the
    lengths; model_helper.py:54-76) and the ``edit_distance`` metric with repeat merging and EOS trimming
    (utils/metrics_utils.py:8-41, host side).  Returns {'loss', 'edit_distance' [B] (numpy), 'sample_ids', ...}."""
    from . import losses, metrics
    pred = las_predict(features, hp, weights, want_alignment=False, want_probs=False)
    dev = pred["encoder_out"].device
    targets = labels["targets_outputs"].to(dev)
    tlen = labels["target_sequence_length"].to(dev)
    out, loss = {}, None
    if "logits" in pred:
        loss = losses.compute_loss(pred["logits"], targets, pred["final_sequence_length"], tlen, "eval", hp["eos_id"])
        out.update(edit_distance=metrics.edit_distance(pred["sample_ids"].cpu().numpy(), targets.cpu().numpy(), hp["eos_id"], hp.get("mapping")),
                   sample_ids=pred["sample_ids"], logits=pred["logits"], final_sequence_length=pred["final_sequence_length"])
    if "logits_binf" in pred:
        # --binf_projection: the same softmax loss on the transformed logits (model_helper.py:326-329; the log-probability
        # regulariser exists in TRAIN only, :243-245,330-331) and 'edit_distance_binf' (:304-306)
        loss_b = losses.compute_loss(pred["logits_binf"], targets, pred["final_sequence_length_binf"], tlen, "eval", hp["eos_id"])
        loss = loss_b if loss is None else loss + loss_b
        out.update(loss_binf=loss_b, logits_binf=pred["logits_binf"], sample_ids_phones_binf=pred["sample_ids_phones_binf"],
                   edit_distance_binf=metrics.edit_distance(pred["sample_ids_phones_binf"].cpu().numpy(), targets.cpu().numpy(),
                                                            hp["eos_id"], hp.get("mapping")))
        out.setdefault("edit_distance", out["edit_distance_binf"])  # model_helper.py:308
    if getattr(weights, "ctc", None) is not None:
        # CTC head (model_helper.py:347-363): loss += mean(ctc_loss) * ctc_weight; 'ctc_edit_distance' of the greedy CTC path
        # (its blank is the last class, while the loss uses blank 0 -- the reference's own mismatch, kept)
        ctc, ctc_logits = losses.ctc_head(pred["encoder_out"], pred["source_length"], targets, tlen, *weights.ctc)
        loss = loss + ctc * float(hp["ctc_weight"])
        best = ctc_logits.argmax(-1).cpu().numpy()
        decoded = metrics.ctc_greedy_decode(best, pred["source_length"].cpu().numpy(), ctc_logits.shape[-1] - 1)
        out.update(ctc_loss=ctc, ctc_edit_distance=metrics.edit_distance(decoded, targets.cpu().numpy(), hp["eos_id"], hp.get("mapping")))
    out["loss"] = loss
    return out


class LASModel:
    """Front-end + listener + speller with device-resident weights."""

    def __init__(self, params, hp, feature_flags, precision="fp32", means=None, stds=None, device="cuda", binf=None):
        """``binf``: binf2phone [n, V] (utils/ipa_utils.py:313-328) for --binary_outputs --binf_projection models."""
        _lib.require_cuda()
        self.hp, self.fa, self.precision = hp, feature_flags, precision
        self.plan = FrontendPlan(feature_flags, means, stds, device)
        self.weights = DeviceWeights(params, hp, num_feature_channels(feature_flags), precision, device, binf=binf)

    @classmethod
    def from_model_dir(cls, model_dir, feature_flags, precision="fp32", means=None, stds=None, device="cuda", binf=None):
        """Build the model from a reference ``model_dir`` as train.py leaves it: ``hparams.json`` (utils/params_utils.py:28-30)
        and the latest TF checkpoint (tf_checkpoint.py), variables under their TF names."""
        from . import hparams as hps, tf_checkpoint
        hp = hps.create_hparams(model_dir=model_dir)
        params = tf_checkpoint.load_model_variables(model_dir)
        if binf is None and "binf2phone" in params:  # --binf_trainable keeps the matrix as a variable (model_helper.py:183)
            binf = params["binf2phone"]
        return cls(params, hp, feature_flags, precision=precision, means=means, stds=stds, device=device, binf=binf)

    def features(self, wave, n_samples=None):
        return self.plan(wave, n_samples)

    def predict_from_features(self, feats, n_frames, **kw):
        return las_predict({"encoder_inputs": feats, "source_sequence_length": n_frames}, self.hp, self.weights, **kw)

    def transcribe(self, wave, n_samples=None, **kw):
        """wave [B,N] float32 on the device -> predictions dict (transcribe_audio_file.py:97-101)."""
        feats, n_frames = self.plan(wave, n_samples)
        return self.predict_from_features(feats, n_frames, **kw)

    def transcribe_host(self, wave_host, n_samples_host=None):
        """Host entry point: pinned host waveform in, host numpy ids out (H2D + D2H inside)."""
        wave = wave_host.to("cuda", non_blocking=True)
        ns = n_samples_host.to("cuda", non_blocking=True) if n_samples_host is not None else None
        pred = self.transcribe(wave, ns)
        key, klen = ("sample_ids", "final_sequence_length") if "sample_ids" in pred else ("sample_ids_phones_binf", "final_sequence_length_binf")
        return pred[key].cpu(), pred[klen].cpu()

    # SMs the recurrence may hold while batches overlap on several streams (the other batch's front-end / GEMMs need the rest)
    PIPELINED_REC_SMS = 64

    def default_streams(self):
        """Compute streams of the serving loop: 2 on the fused bf16 tensor-core path (its recurrence occupies 64 of the 148 SMs
        and its decoder is latency-bound, so the next batch's front-end / GEMMs / recurrence fill the idle SMs: measured
        14.5 -> 11.7 ms per c2 batch), 1 otherwise."""
        w = self.weights
        fused = (self.precision == "bf16" and getattr(w.speller, "tc", False)
                 and all(lw.get("whh_tc") is not None for lw in w.listener.layers))
        return 2 if fused else 1

    def transcribe_stream(self, host_batches, n_streams=None, ahead=None, want_alignment=False):
        """Serving loop over pinned host waveform batches ([B,N] float32 each), fully pipelined: the host->device copy of a batch
        runs on a copy stream into one of ``n_streams + 1`` preallocated device buffers, consecutive batches run on ``n_streams``
        compute streams (default: ``default_streams()``), nothing in a step synchronises the host (the decode step count stays
        on the device), and the ids of a batch are read back (pinned, asynchronous) while the following batches are already
        enqueued.  Kernels that need the whole GPU co-resident never overlap each other (_lib.grid_sync_kernel).  Yields
        (sample_ids [B,steps], final_sequence_length [B]) host tensors per batch, in order.  Batches that are already device
        tensors skip the staging copy (same streams, pacing and read-back)."""
        from collections import deque
        dev = self.plan.device
        ns = int(n_streams or self.default_streams())
        ahead = int(ahead or ns)  # batches the host may have in flight before it waits for the oldest one's ids
        nb = ahead + 1
        if int(self.hp.get("beam_width", 0) or 0) > 0:
            raise NotImplementedError("transcribe_stream serves greedy decoding (beam search returns [B, T, W] ids: use transcribe)")
        # staging state lives on the model: cudaMalloc / cudaHostAlloc are slow and synchronise the device, so the streams,
        # the device input buffers and the pinned result buffers are created once and reused
        st = self.__dict__.setdefault("_stream_state", {"copy_stream": torch.cuda.Stream(device=dev), "compute": [], "bufs": [], "pinned": []})
        while len(st["compute"]) < ns:
            st["compute"].append(torch.cuda.Stream(device=dev))
        while len(st["bufs"]) < nb:
            st["bufs"].append(None)
            st["pinned"].append(None)
        copy_stream, compute, bufs, pinned = st["copy_stream"], st["compute"], st["bufs"], st["pinned"]
        consumed = [None] * nb  # "kernels are done with the staging buffer" events of this call
        main_stream = torch.cuda.current_stream()
        main_stream.synchronize()  # buffers may still be in use by an earlier call
        for cs in compute[:ns]:
            cs.wait_stream(main_stream)

        def launch(hw, i):
            slot, cs = i % nb, compute[i % ns]
            if hw.is_cuda:  # already resident (bench.py's device-resident arm): no staging copy, same pacing and read-back
                with torch.cuda.stream(cs), _lib.rec_sms(self.PIPELINED_REC_SMS if ns > 1 else 0):
                    return run(hw, slot, cs)
            with torch.cuda.stream(copy_stream):
                if bufs[slot] is None or bufs[slot].shape != hw.shape:
                    # allocated under the copy stream: the block is then ordered after its previous owner's work on this stream
                    bufs[slot] = torch.empty(hw.shape, dtype=torch.float32, device=dev)
                if consumed[slot] is not None:
                    copy_stream.wait_event(consumed[slot])  # do not overwrite a batch that is still being read
                bufs[slot].copy_(hw, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            with torch.cuda.stream(cs), _lib.rec_sms(self.PIPELINED_REC_SMS if ns > 1 else 0):
                cs.wait_event(ev)
                return run(bufs[slot], slot, cs)

        def run(wave, slot, cs):  # under stream cs: transcribe + asynchronous read-back of ids / lengths / step count
            # want_alignment: the decoder also writes predictions['alignment'] (it stays on the device; only ids / lengths are read back)
            pred = self.transcribe(wave, want_alignment=want_alignment, trim=False, want_probs=False)
            done = torch.cuda.Event()
            done.record(cs)
            consumed[slot] = done
            key, klen = (("sample_ids", "final_sequence_length") if "sample_ids" in pred
                         else ("sample_ids_phones_binf", "final_sequence_length_binf"))  # --binf_projection without --multitask
            ids_d, len_d, n_d = pred[key], pred[klen], pred["n_steps"]
            if pinned[slot] is None or pinned[slot][0].shape != ids_d.shape:
                pinned[slot] = (torch.empty(ids_d.shape, dtype=ids_d.dtype).pin_memory(),
                                torch.empty(len_d.shape, dtype=len_d.dtype).pin_memory(),
                                torch.empty(n_d.shape, dtype=n_d.dtype).pin_memory())
            for h_t, d_t in zip(pinned[slot], (ids_d, len_d, n_d)):
                h_t.copy_(d_t, non_blocking=True)
            rd = torch.cuda.Event()
            rd.record(cs)
            return pinned[slot], rd

        def finish(job):
            (ids_h, len_h, n_h), ev = job
            ev.synchronize()
            n = int(n_h[0])
            return ids_h[:, :n].clone(), len_h.clone()

        pending = deque()
        for i, hw in enumerate(host_batches):
            if len(pending) == ahead:  # the slot batch i is about to reuse belongs to batch i - ahead - 1: already yielded
                yield finish(pending.popleft())
            pending.append(launch(hw, i))
        while pending:
            yield finish(pending.popleft())
        for cs in compute[:ns]:
            main_stream.wait_stream(cs)
