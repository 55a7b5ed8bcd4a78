from ..gp_models import GPRegression  # noqa: F401
