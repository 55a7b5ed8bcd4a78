from ...gp_models import (convert_x_list_to_array, convert_xy_lists_to_arrays,  # noqa: F401
                          convert_y_list_to_array)
