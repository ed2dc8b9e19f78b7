"""distance3d_b200 -- B200-native batched narrow- and broad-phase collision engine
with the Python API of AlexanderFabisch/distance3d on its hot path.

Sub-modules mirror the reference: ``colliders``, ``gjk``, ``epa``, ``mpr``,
``aabb_tree``, ``containment``, ``broad_phase``, ``self_collision``, ``urdf_utils``,
``random``, ``geometry``, ``minkowski``, ``mesh``, ``io``, ``utils``, ``hydroelastic_contact`` (broad phase +
tetrahedron pairs); ``pack`` and ``parallel`` are the batched / multi-GPU additions.  All
geometry is computed by the CUDA library ``libd3d_b200.so`` (include/d3d_b200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import pack, colliders, random, urdf, urdf_utils, parallel  # noqa: F401  (host-only imports)
from . import gjk, epa, mpr, aabb_tree, containment, broad_phase, self_collision  # noqa: F401
from . import benchmark, pipeline, utils, geometry, minkowski, mesh, stream, io, hydroelastic_contact  # noqa: F401
