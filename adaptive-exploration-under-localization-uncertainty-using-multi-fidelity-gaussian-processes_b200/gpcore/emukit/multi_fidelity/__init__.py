from . import convert_lists_to_array, kernels, models  # noqa: F401
