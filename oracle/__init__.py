"""CPU oracle for the SALSA hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy / torch-fp32 restatement of the reference's algorithm
(thomeou/SALSA, `dataset/salsa_feature_extraction.py`,
`dataset/salsa_lite_feature_extraction.py`, `models/*`).  It exists to check the
CUDA path and to be timed as the CPU baseline.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py` (`cpu_baseline` / `--impl reference`)
may import it.  Nothing under `salsa_b200/` imports it, and the product path
raises if the CUDA library is missing -- there is no CPU fallback.

Parity pinning
--------------
The reference ships no tests, golden vectors or fixtures, and
`librosa` (un-vendored dependency, pinned 0.8.0 in `requirements.yml:101`) is
absent from this image.  The oracle is therefore pinned like this:

* `extract_normalized_eigenvector` and `MagStftExtractor` were executed
  VERBATIM from `/root/reference` (behind stub modules, `oracle/ref_import.py`)
  and their outputs are frozen in `tests/golden/*.npz` by
  `oracle/make_golden.py`; `tests/test_oracle_golden.py` checks the restatement
  against those files on every run, and against the live reference when
  `/root/reference` is present.
* `librosa.stft` / `librosa.power_to_db` are restated from librosa 0.8.0's
  published algorithm (`oracle/stft.py`); that restatement is cross-checked
  against `torch.stft` and `scipy.signal.stft`, but NOT against librosa
  itself: for the STFT stage alone, parity is unpinned.
* The model oracle is the reference's own torch modules executed verbatim (for
  the golden logits) plus a plain functional fp32 restatement in
  `oracle/crnn.py`.
"""
