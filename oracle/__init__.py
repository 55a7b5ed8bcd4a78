"""CPU oracle for the dense GP inference core -- TEST INFRASTRUCTURE ONLY.

This package is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The shipped path (``gpcore`` -> ``libgpcore.so``) never routes
through anything in here and fails loudly when the CUDA library is missing.

Parity status
-------------
* NIGP (``NIGP.py``): PINNED.  The reference module is imported verbatim in the
  build container through ``oracle/gpy_shim`` and its outputs are committed as
  ``tests/golden/nigp_*.npz`` (generator: ``oracle/make_golden.py``).
* GPy ``GPRegression`` / emukit linear multi-fidelity arithmetic: PARITY UNPINNED by golden
  vectors.  Neither GPy nor emukit is vendored in the reference or installed here (versions are
  not pinned by the reference either); ``gp_oracle.py`` restates their published
  algorithms, every third-party quirk being an explicit switch.  What the reference DOES hold at
  that boundary is used: on eleven of twelve bundled data sets the full ``GPTrainers.py`` flow lands
  within 5e-6 (RMSE) / four digits (covariance-weighted MSE) of the numbers the reference published
  with real GPy / emukit (``tests/golden/gp_datasets.npz``, ``tests/test_gpu_widen.py``), and a
  closed-form AR1 known-answer test bypasses this package altogether.
* BASELINE sizes: ``make_golden_scale.py`` freezes the outputs of the reference's own ``NIGP.py`` at
  N = 8192 (``nigp_8192.npz``), its ``NIGP.fit`` (``nigp_fit.npz``) and of this restatement at
  N = 2048 / 4096 / 16384 (``scale_oracle.npz``, literal refit loops for the information gain).
"""
