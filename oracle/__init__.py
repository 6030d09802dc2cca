"""CPU oracle for the SQD hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package
``qiskit_addon_sqd_b200``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker
or as the timed CPU baseline -- never as a fallback for the CUDA path.
"""
