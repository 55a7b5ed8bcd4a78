from ...gp_models import GPyLinearMultiFidelityModel  # noqa: F401
