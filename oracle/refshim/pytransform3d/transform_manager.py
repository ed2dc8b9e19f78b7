from distance3d_b200.urdf import TransformManager  # noqa: F401
