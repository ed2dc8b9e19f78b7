/* TEST INFRASTRUCTURE (CPU oracle): forward kinematics of a flattened URDF chain.
 * Restates what the reference obtains from pytransform3d one configuration at a time
 * (UrdfTransformManager.set_joint + get_transform; call sites distance3d/broad_phase.py:111,148):
 * pose(frame k) = prod_s fixed_s * joint_s(q), revolute joints by Rodrigues' formula, joint
 * values clipped to their limits.  Same chain description as d3d_fk_urdf (include/d3d_b200.h). */
#include <math.h>
#include <string.h>
#include "d3d_oracle.h"

static void matmul4(const double *A, const double *B, double *C) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
            C[4 * r + c] = s;
        }
}

void d3do_fk(int n_frames, int n_joints, const double *joint_axis, const double *joint_limits,
             const int32_t *joint_type, const int32_t *chain_off, const double *chain_fixed,
             const int32_t *chain_joint, const double *q, int64_t n_cfg, double *out_pose,
             int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int64_t b = 0; b < n_cfg; ++b)
        for (int k = 0; k < n_frames; ++k) {
            double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, U[16], J[16];
            for (int s = chain_off[k]; s < chain_off[k + 1]; ++s) {
                matmul4(T, chain_fixed + 16 * (int64_t)s, U);
                int j = chain_joint[s];
                if (j < 0) { memcpy(T, U, sizeof(T)); continue; }
                double v = q[b * n_joints + j];
                if (v < joint_limits[2 * j]) v = joint_limits[2 * j];
                if (v > joint_limits[2 * j + 1]) v = joint_limits[2 * j + 1];
                double ux = joint_axis[3 * j], uy = joint_axis[3 * j + 1], uz = joint_axis[3 * j + 2];
                memset(J, 0, sizeof(J));
                J[15] = 1.0;
                if (joint_type[j] == 0) {
                    double c = cos(v), sn = sin(v), ci = 1.0 - c;
                    J[0] = ci * ux * ux + c;       J[1] = ci * ux * uy - uz * sn; J[2] = ci * ux * uz + uy * sn;
                    J[4] = ci * uy * ux + uz * sn; J[5] = ci * uy * uy + c;       J[6] = ci * uy * uz - ux * sn;
                    J[8] = ci * uz * ux - uy * sn; J[9] = ci * uz * uy + ux * sn; J[10] = ci * uz * uz + c;
                } else {
                    J[0] = J[5] = J[10] = 1.0;
                    J[3] = v * ux; J[7] = v * uy; J[11] = v * uz;
                }
                matmul4(U, J, T);
            }
            memcpy(out_pose + 16 * (b * n_frames + k), T, sizeof(T));
        }
}
