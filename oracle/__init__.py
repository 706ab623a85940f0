"""CPU oracle for the white-box attack hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``vitatk``) never imports this package and fails
loudly when its CUDA library is missing.
"""
