"""CPU oracle for the EVStore embedding-lookup hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import or execute it, and there
only as the checker.  The product (``ev-store-dlrm_b200/``) never imports it
and fails loudly when its CUDA library is missing.

Parity pin: ``oracle.evlfu.SeqEvLFU`` is checked request-by-request against the
reference's own ``cache_algo/EvLFU_C1.py`` (imported from /root/reference by
``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/``).  ``oracle.evlfu.BatchEvLFU`` is the batch-granular policy
the CUDA path implements; at batch size 1 it equals ``SeqEvLFU`` on every
request that does not hit the documented same-request re-fetch corner.
"""
