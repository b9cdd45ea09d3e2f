"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU (NumPy/SciPy) restatement of the per-trial spectral hot path of
esi-neuroscience/syncopy v2023.09 (freqanalysis / connectivityanalysis compute
bodies).  It exists to *check* the CUDA engine in ``syncopy_b200``; it is never
shipped or measured as the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline legs of ``bench.py`` may import it.  The product package
``syncopy_b200`` must never import from here (tests/test_no_oracle_in_product.py
enforces that).

Parity status: PINNED.  ``oracle/make_golden.py`` imports the *real* reference
backends by file path from /root/reference (see ``oracle/ref_loader.py``) and
(1) asserts the restatement agrees with them on seeded inputs and
(2) writes the reference's own outputs to ``tests/golden/*.npz``; the CPU test
suite re-checks the oracle against those committed vectors everywhere, and
against the live reference whenever /root/reference is present.  The reference's
own known-answer tests (syncopy/tests/backend/test_timefreq.py,
test_conn.py) are transcribed in tests/test_oracle_known_answers.py.
"""
