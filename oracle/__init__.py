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
* GPy ``GPRegression`` / emukit linear multi-fidelity arithmetic: PARITY UNPINNED.
  Neither GPy nor emukit is vendored in the reference or installed here (versions are
  not pinned by the reference either); ``gp_oracle.py`` restates their published
  algorithms, every third-party quirk being an explicit switch.
"""
