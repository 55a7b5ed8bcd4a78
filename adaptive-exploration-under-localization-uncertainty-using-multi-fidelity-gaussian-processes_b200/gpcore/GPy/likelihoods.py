from ..gp_models import Gaussian, MixedNoise  # noqa: F401
