"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the hydra-pspec Gibbs hot path.

Nothing under ``hydra_pspec_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may use it, and only as the checker / the timed CPU baseline.
"""
