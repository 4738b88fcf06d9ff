"""CPU oracle for the DC-VIC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package.  Nothing under ``dc_vic_b200/`` imports it (a test enforces that).

Contents
--------
``vq_oracle``       restatement of taming's ``VectorQuantizer`` / ``VectorQuantizer2``
                    (reference: ``taming/modules/vqvae/quantize.py``).  PINNED: checked
                    against goldens produced by importing that vendored file
                    (``tests/golden/make_golden.py``).
``entropy_oracle``  restatement of CompressAI 1.2.4 ``EntropyModel`` / ``GaussianConditional``
                    / ``EntropyBottleneck`` / ``LowerBound`` (the un-vendored dependency of
                    ``src/models/subnet/entropy_model/*.py``) plus the DC-VIC wrappers.
                    PARITY UNPINNED at the CompressAI boundary: compressai==1.2.4
                    (pyproject.toml:13, poetry.lock:312-313) is not installed, has no wheel
                    in the offline wheelhouse, and the reference ships no tests or golden
                    vectors for it.  The restatement follows the published 1.2.4 algorithm
                     and is cross-checked in FP64 and against
                    closed-form identities only.
"""
