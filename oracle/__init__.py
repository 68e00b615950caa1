"""CPU oracle: a restatement of the reference's algorithms for the hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; nothing under ``pfst_b200/`` does (a test enforces
this). It is plain PyTorch/numpy on whatever device its inputs live on (CPU in
practice), written to issue the SAME ATen operator sequence as the reference so
that (a) results are bit-identical to the reference on the same torch build and
(b) its wall time is a fair stand-in for the reference's CPU path.

Pinning status:
  * metrics  — pinned by the reference's own golden test tests/test_metrics.py
               (re-targeted in tests/test_oracle_metrics.py) AND against the
               reference module loaded by path.
  * ema / pseudo-label / ClassMix / PFGSTLoss — the reference ships no tests or
               fixtures for these; pinned against outputs of the reference code
               itself, loaded by path from /root/reference with import stubs
               (tests/golden/make_golden.py -> tests/golden/*.npz, committed).
  * prototypes (P1–P3) — north_star extension with no reference code:
               PARITY UNPINNED; anchored on PFGST.masked_feat_dist only.

Each function cites the reference file:line it follows.
"""
