"""CPU oracle for the audio-sheet retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`audio_sheet_retrieval_b200/`) may import this package.  The only permitted
callers are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`.

Each function restates the reference algorithm (CPJKU/audio_sheet_retrieval)
and cites the reference file:line it follows (paths relative to the reference
root, `asr/` = `audio_sheet_retrieval/`).

Pinning status
--------------
* NumPy/SciPy parts of the reference (`eval_retrieval`, `CCA.fit('svd')`,
  `batch_compute1/2`, `_retrieve_*`, `detect_*`) ARE pinned: the reference's own
  source text is executed under Python 3 by `tests/golden/make_golden.py`
  (py2 shims only: xrange, print statement) and its outputs are committed as
  fixtures which `tests/test_oracle_golden.py` checks this oracle against.
* Theano/Lasagne parts (the two encoders, `CCALayer`) are PARITY UNPINNED:
  neither Theano, Lasagne nor Python 2 exist in this image and the reference
  has no tests or stored activations.  The restatement follows Lasagne's
  documented semantics; the only pins are the shipped
  parameter pickle's layout and self-consistent BN/CCA statistics.
"""
