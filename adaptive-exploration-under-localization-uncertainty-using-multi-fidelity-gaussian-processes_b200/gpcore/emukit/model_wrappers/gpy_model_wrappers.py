from ...gp_models import GPyMultiOutputWrapper  # noqa: F401
