"""TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy/SciPy) of the GP_emu_UQSA dense-GP hot path.  Nothing in the
product package ``gp_emu_uqsa_b200`` imports this; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and only as the checker / reported baseline.

Parity status: PINNED.  ``oracle/gp_oracle.py`` was validated in the build container
against the real reference imported from /root/reference (``oracle/ref_loader.py``,
``tests/golden/make_golden.py``); the resulting vectors are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` re-checks the restatement against
them on every run.
"""
