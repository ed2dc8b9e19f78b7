from distance3d_b200.urdf import (  # noqa: F401
    UrdfTransformManager, Geometry, Sphere, Box, Cylinder, Mesh)
