from distance3d_b200._transforms import (  # noqa: F401
    perpendicular_to_vector, norm_vector, matrix_from_axis_angle,
    active_matrix_from_angle, active_matrix_from_extrinsic_euler_xyz)
