"""CPU oracle for the trust-region inverse-compositional solver path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``deep_prob_feature_track_b200`` may import
this package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU arm being timed -- never on the product path.
"""
