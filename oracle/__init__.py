"""CPU oracle for the nellie structure-enhancement hot path (TEST INFRASTRUCTURE ONLY).

Nothing in ``nellie_b200`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  ``oracle/make_golden.py`` executes the unmodified reference
(``/root/reference/nellie/segmentation/{filtering,labelling}.py`` through the import
shim in ``oracle/ref_shim.py``) in the build container and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
those vectors bit-for-bit, plus the reference's own ``tests/test_labelling.py`` cases.
"""
