/* Packed collider buffers shared by the CUDA library (include/d3d_b200.h) and
 * the CPU oracle (oracle/src/d3d_oracle.h).  Plain C, no torch types.
 *
 * Structure-of-arrays record of N colliders; replaces the Python collider
 * objects of the reference (distance3d/colliders.py:17-646).  For the CUDA
 * library every pointer is a DEVICE pointer, for the oracle a host pointer.
 */
#ifndef D3D_TYPES_H
#define D3D_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Collider type tags (reference class in parentheses). */
enum {
    D3D_SPHERE = 0,    /* colliders.py:243  Sphere(center, radius)                    */
    D3D_CAPSULE = 1,   /* colliders.py:290  Capsule(pose, radius, height)             */
    D3D_BOX = 2,       /* colliders.py:147  Box(pose, size): 8 world vertices in pool  */
    D3D_ELLIPSOID = 3, /* colliders.py:343  Ellipsoid(pose, radii)                    */
    D3D_CYLINDER = 4,  /* colliders.py:390  Cylinder(pose, radius, length)            */
    D3D_HULL = 5,      /* colliders.py:109  ConvexHullVertices: world vertices in pool */
    D3D_MESH = 6,      /* colliders.py:187  MeshGraph: local vertices in pool + pose   */
    D3D_DISK = 7,      /* colliders.py:443  Disk(center, radius, normal)              */
    D3D_ELLIPSE = 8,   /* colliders.py:500  Ellipse(center, axes, radii)              */
    D3D_CONE = 9,      /* colliders.py:554  Cone(pose, radius, height)                */
    D3D_NUM_TYPES = 10
};

/* pose:  row-major 4x4 collider2origin.  Sphere/Disk/Ellipse use only parts of
 *        it (centre = pose[:3,3]; disk normal = pose[:3,2]; ellipse axes =
 *        pose[:3,0], pose[:3,1]).
 * param: sphere (r,-,-); capsule (r,h,-); box (sx,sy,sz); ellipsoid (rx,ry,rz);
 *        cylinder (r,L,-); disk (r,-,-); ellipse (r0,r1,-); cone (r,h,-).
 * vert_off/vert_len: range into verts[M,3] (box: 8 world-frame vertices written
 *        by d3d_prepare; hull: world frame; mesh: local frame).  A hull / mesh has at least
 *        one vertex: the reference's np.argmax raises on an empty array (colliders.py:131-132)
 *        and the Python layer rejects it; the kernels read vert_len < 1 as the single point
 *        (0,0,0) of the collider's frame instead of reading outside the range.
 * margin: optional per-collider Margin (colliders.py:606), NULL = none.
 *
 * MeshGraph support = hill climbing over the triangle graph (mesh.py:12-139).
 * graph_off[i] >= 0 points at the mesh's adjacency record in the int32 pool
 * `graph` (meshes that share vertices may share the record):
 *     g[0]          first_idx = min(triangles)                      (mesh.py:29)
 *     g[1..6]       shortcut vertices: argmax x,y,z, argmin x,y,z   (mesh.py:44-47)
 *     g[7..7+nv]    row pointers (nv+1 entries, relative to g): the neighbours of
 *                   vertex v are g[g[7+v]] .. g[g[7+v+1]-1], in the order the
 *                   reference iterates `connections[v]`             (mesh.py:49-52)
 * graph_off == NULL or graph_off[i] < 0: arg-max over all vertices
 * (MeshSupportFunction, mesh.py:142-191).
 * mesh_start: optional start vertex per collider (the reference object's cached
 *        `first_idx`, mesh.py:85); NULL or < 0 = g[0].  Every pair starts from it
 *        and carries the vertex from one support call to the next, as a fresh
 *        reference object does within one gjk / epa / mpr call.
 * mesh_last: optional OUTPUT, vertex a pair ended on (written per collider; only
 *        meaningful when a collider occurs in one pair - the scalar API).      */
typedef struct d3d_colliders {
    int64_t n;
    const int32_t *type;     /* [n]    */
    const double *pose;      /* [n,16] */
    const double *param;     /* [n,3]  */
    const int32_t *vert_off; /* [n]    */
    const int32_t *vert_len; /* [n]    */
    const double *verts;     /* [M,3]  */
    const double *margin;    /* [n] or NULL */
    const int32_t *graph_off;  /* [n] or NULL */
    const int32_t *graph;      /* [G] adjacency pool or NULL */
    const int32_t *mesh_start; /* [n] or NULL */
    int32_t *mesh_last;        /* [n] or NULL (output) */
} d3d_colliders;

/* Per-pair status codes; 0..3 are the reference's GjkState values
 * (distance3d/gjk/_gjk_jolt.py:22-26). */
enum {
    D3D_NO_INTERSECTION = 0,
    D3D_INTERSECTION = 1,
    D3D_UNKNOWN = 2,
    D3D_CLIPPED = 3,          /* -> (MAX_FLOAT, None, None, None), _gjk_jolt.py:209 */
    D3D_SANITY_FAILED = 4,    /* reference raises AssertionError, _gjk_jolt.py:216   */
    D3D_MONOTONICITY = 5,     /* reference raises AssertionError, _gjk_jolt.py:126,282 */
    D3D_ITER_CAP = 6,         /* no reference equivalent (reference loops forever)  */
    D3D_EPA_MAX_FACES = 7,    /* reference raises AssertionError, epa.py:128        */
    D3D_EPA_BAD_SIMPLEX = 8   /* GJK simplex with < 4 points handed to EPA          */
};

#define D3D_GJK_ITER_CAP 1024

#ifdef __cplusplus
}
#endif
#endif
