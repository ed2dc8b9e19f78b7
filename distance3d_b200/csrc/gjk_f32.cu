// Opt-in fp32 instantiation of the GJK kernels (same source, `real` = float).
//
// Exports d3d_gjk_distance_f32 / d3d_gjk_intersection_f32 with the signatures of their
// fp64 counterparts (buffers stay fp64 in HBM, conversion happens at load / store).  No
// BLAS / x87 conventions are emulated in this mode; its tolerance against the fp64 path is
// stated in DESIGN.md and tests/test_gjk_gpu.py.
#define D3D_F32 1
#include "gjk.cu"
