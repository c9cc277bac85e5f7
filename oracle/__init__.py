"""CPU oracle for the pcb200 hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``pytorch_connectomics_b200/`` may import this package: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the timed baseline.
"""
