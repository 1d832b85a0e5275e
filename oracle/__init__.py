"""CPU oracle for the HAConvDR exact inner-product retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline - never as the thing shipped.  The product
(``haconvdr_b200``) never imports this package and has no CPU fallback.

PARITY STATUS: *unpinned for the faiss arithmetic*.  The reference
(`/root/reference/src/test_HAConvDR_topiocqa.py:20,52,98,102,122`) delegates all
scoring / selection to the third-party ``faiss-gpu 1.7.2`` (pinned only in prose,
`/root/reference/README.md:12`), which is neither vendored under
`/root/reference` nor installed in this image, and the reference ships no tests,
golden vectors or fixtures.  The oracle therefore restates the published
``IndexFlatIP`` algorithm (see ``flat_ip.py``) and is pinned by
  * analytic known-answer sets (integer-valued scores, exact ties, planted tops),
  * an fp64 brute-force arbiter,
  * the reference's *own* orchestration function run in the build container
    (``ref_harness.py``; outputs committed under ``tests/golden/``).
"""
