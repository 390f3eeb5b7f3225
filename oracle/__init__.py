"""ORACLE -- test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from this package, and only as the checker / reported CPU baseline.  The product
path (``loco_asr_b200``) never does and fails loudly when its CUDA extension is missing.
"""
