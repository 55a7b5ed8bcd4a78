"""Minimal stand-in for the ``GPy`` import of the reference ``NIGP.py`` (line 6).

TEST INFRASTRUCTURE ONLY.  GPy is not installed in this image; the reference NIGP
module only needs ``GPy.kern.RBF(input_dim, variance, lengthscale, ARD, inv_l).K``
(``NIGP.py:18-19``).  This shim restates that one call with GPy's published
``Stationary`` arithmetic so ``NIGP.py`` can be imported verbatim for golden-vector
generation (``oracle/make_golden.py``).
"""
from . import kern  # noqa: F401
