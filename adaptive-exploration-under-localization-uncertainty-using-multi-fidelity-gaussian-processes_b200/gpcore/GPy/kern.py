from ..gp_models import RBF, Matern32  # noqa: F401
