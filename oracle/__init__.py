"""CPU oracle for the PReMVOS hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or as the timed CPU baseline.
The product path (``premvos_b200``) never imports this package and raises when the
CUDA library is missing.
"""
