from ...gp_models import LinearMultiFidelityKernel  # noqa: F401
