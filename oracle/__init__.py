"""CPU oracle for the phones-las hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in plain numpy, the arithmetic of the reference's hot path
(acoustic front-end -> pyramidal BiLSTM listener -> attention decoder -> losses).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; nothing under ``phones_las_b200/`` does.

PARITY UNPINNED: the reference (sciforce/phones-las) ships no tests, golden vectors
or known-answer fixtures and its arithmetic lives in
un-vendored third-party packages that cannot be installed in this image
(tensorflow==1.15.2, librosa==0.7.1, speechpy==2.4; reference requirements.txt:12,16,19).
The oracle therefore restates those packages' published algorithms at the
reference's call sites and is cross-checked against independent implementations
that ARE available (torch.nn.LSTM, torch ctc_loss, scipy savgol/dct, numpy rfft,
torchaudio mel filterbanks) in tests/test_oracle_*.py.  The librosa branch of the
front-end is additionally checked END TO END (mel spectrogram, dB with top_db, the
amplitude_to_db-of-power quirk, MFCC) against torchaudio.transforms and
transformers.audio_utils, two third-party implementations written to reproduce
librosa (tests/test_oracle_frontend.py::test_librosa_*_vs_*): agreement to 4e-6 dB /
5e-5 on MFCCs.  That anchors the librosa rows to something other than this restatement;
the speechpy rows and the TF seq2seq decoder have no such second implementation here
(the monotonic attention's closed forms are checked against the recurrence of Raffel et al. 2017
they stand for, tests/test_oracle_las.py::test_monotonic_attention_closed_form_matches_the_recurrence;
the Levenshtein core of the edit-distance metric against torchaudio.functional.edit_distance).
"""
