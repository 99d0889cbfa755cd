"""CPU oracle for the JFNK hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or as the timed CPU baseline), never as
the shipped path.  The product path (``iterative-solvers-summer-2020_b200``) never
imports this package and has no CPU fallback.

Parity status: the reference (Shiakaron/Iterative-solvers-summer-2020) ships no tests
and no expected outputs, so the oracle is pinned by
  (a) SciPy 1.18.1 itself (the reference's third-party solver: ``scipy.optimize.
      newton_krylov`` / ``scipy.sparse.linalg.lgmres``), which the oracle CALLS rather
      than restates, and
  (b) golden vectors generated in the build container by importing the reference's own
      Python modules (``tests/golden/make_golden.py``), against which the numpy
      restatements in ``oracle/pma2.py`` and ``oracle/droplet.py`` are checked.
"""
