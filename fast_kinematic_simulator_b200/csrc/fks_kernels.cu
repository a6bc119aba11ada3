// sm_100a kernels of the batched particle contact simulator.
//
// One warp simulates one particle from start to finish: the controller-step loop
// (simple_particle_contact_simulator.hpp:843-919), the microstep loop and the contact resolver
// (:1546-1816).  Lanes stride over the robot's collision points; reductions over points are warp
// shuffles / votes; D-vectors (one entry per actuated axis) live one entry per lane; link transforms
// live in shared memory.  The grid is persistent: warps pull particle ids from a global counter so
// particles that spend 26 resolver iterations per microstep do not stall their neighbours.
//
// Nothing here is a dense contraction, so no tensor cores: the contended units are the FP64 pipe and
// the LSU/L2 gather path (SDF floats through the read-only path, L2-resident via an access-policy
// window).
//
// Reference line numbers below (spcs = simple_particle_contact_simulator.hpp, tnuva =
// tnuva_robot_models.hpp, unc = simple_uncertainty_models.hpp, pid = simple_pid_controller.hpp)
// say WHAT each function computes; the arithmetic of the un-vendored dependencies (arc_utilities,
// sdf_tools, Eigen) is written from their documented behaviour, see DESIGN.md.

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "fks_device_types.h"
#include "fks_philox.h"

namespace fksdev {

#define FKS_FULL 0xffffffffu

namespace {

constexpr double kPi = 3.14159265358979323846;

struct Ctx {
    const DevRobot* rb;  // shared memory copy
    const double* px;
    const double* py;
    const double* pz;
    const int* plink;
    double* ws;  // this warp's shared block
    WarpLayout wl;
    int lane, L, J, D, P, stride;
    // global scratch of this warp slot
    double* Js;
    int ldj;
    double* selfcorr;
    double* selfwork;
    int* keys;
    unsigned char* sflag;
    unsigned lflags;  // lane-local FKS_FLAG_* bits, OR-reduced when the particle ends
    unsigned cand_links;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FKS_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FKS_FULL, v, o));
    return v;
}

// EigenHelpers::EnforceContinuousRevoluteBounds: wrap to (-pi, pi]
__device__ __forceinline__ double wrap_angle(double value) {
    if ((value <= -kPi) || (value > kPi)) {
        const double remainder = fmod(value, 2.0 * kPi);
        if (remainder <= -kPi) return remainder + (2.0 * kPi);
        if (remainder > kPi) return remainder - (2.0 * kPi);
        return remainder;
    }
    return value;
}

__device__ __forceinline__ void apply_T(const double* T, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = T[0] * x + T[1] * y + T[2] * z + T[3];
    oy = T[4] * x + T[5] * y + T[6] * z + T[7];
    oz = T[8] * x + T[9] * y + T[10] * z + T[11];
}

// Quaterniond(AngleAxisd(angle, axis)).toRotationMatrix(), translation zero
__device__ __forceinline__ void rot_axis(double angle, double ax, double ay, double az, double* R) {
    double s, w;
    sincos(0.5 * angle, &s, &w);
    const double x = s * ax, y = s * ay, z = s * az;
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = 0.0;
    R[4] = txy + twz;
    R[5] = 1.0 - (txx + tzz);
    R[6] = tyz - twx;
    R[7] = 0.0;
    R[8] = txz - twy;
    R[9] = tyz + twx;
    R[10] = 1.0 - (txx + tyy);
    R[11] = 0.0;
}

// C = A * B for rigid transforms stored row-major 3x4
__device__ __forceinline__ void iso_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * r + 0] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
        C[4 * r + 3] = A[4 * r + 0] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
    }
}
__device__ __forceinline__ void iso_inverse(const double* A, double* C) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * c + r];
#pragma unroll
    for (int r = 0; r < 3; r++) C[4 * r + 3] = -(C[4 * r + 0] * A[3] + C[4 * r + 1] * A[7] + C[4 * r + 2] * A[11]);
}

// EigenHelpers::ExpTwist(twist, 1.0) (call sites tnuva:360,378): twist = (v, w)
__device__ __forceinline__ void exp_twist(const double* tw, double* T) {
    const double rn = sqrt(tw[3] * tw[3] + tw[4] * tw[4] + tw[5] * tw[5]);
#pragma unroll
    for (int i = 0; i < 12; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (rn >= 1e-100) {
        const double theta = rn * 1.0;
        const double svx = tw[0] / rn, svy = tw[1] / rn, svz = tw[2] / rn;
        const double wx = tw[3] / rn, wy = tw[4] / rn, wz = tw[5] / rn;
        double s, c;
        sincos(theta, &s, &c);
        const double c1 = 1.0 - c;
        const double K[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
        double K2[9], R[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) K2[3 * r + cc] = K[3 * r + 0] * K[cc] + K[3 * r + 1] * K[3 + cc] + K[3 * r + 2] * K[6 + cc];
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + (K[i] * s) + (K2[i] * c1);
        const double cxv[3] = {wy * svz - wz * svy, wz * svx - wx * svz, wx * svy - wy * svx};
        const double wv = wx * svx + wy * svy + wz * svz;
        const double wa[3] = {wx, wy, wz};
#pragma unroll
        for (int r = 0; r < 3; r++) {
            double acc = 0.0;
#pragma unroll
            for (int cc = 0; cc < 3; cc++) acc += (((r == cc) ? 1.0 : 0.0) - R[3 * r + cc]) * cxv[cc];
            T[4 * r + 3] = acc + (wa[r] * wv) * theta;
#pragma unroll
            for (int cc = 0; cc < 3; cc++) T[4 * r + cc] = R[3 * r + cc];
        }
    } else {
        T[3] = tw[0];
        T[7] = tw[1];
        T[11] = tw[2];
    }
}

__device__ __forceinline__ void safe_normal3(double& x, double& y, double& z) {
    const double n = sqrt(x * x + y * y + z * z);
    if (n > DBL_EPSILON) {
        x = x / n;
        y = y / n;
        z = z / n;
    }
}

// EigenHelpers::TwistBetweenTransforms(a, b) = unhat(log(a^-1 b)) (call site tnuva:389), closed form
__device__ void twist_between(const double* a, const double* b, double* twist) {
    double ai[12], Dm[12];
    iso_inverse(a, ai);
    iso_mul(ai, b, Dm);
    const double tr = Dm[0] + Dm[5] + Dm[10];
    const double axx = Dm[9] - Dm[6], axy = Dm[2] - Dm[8], axz = Dm[4] - Dm[1];
    const double s2 = sqrt(axx * axx + axy * axy + axz * axz);
    const double c = 0.5 * (tr - 1.0);
    const double theta = atan2(0.5 * s2, c);
    const double tx = Dm[3], ty = Dm[7], tz = Dm[11];
    double wx, wy, wz;
    if (theta < 1e-9) {
        wx = axx * 0.5;
        wy = axy * 0.5;
        wz = axz * 0.5;
        twist[0] = tx - (wy * tz - wz * ty) * 0.5;
        twist[1] = ty - (wz * tx - wx * tz) * 0.5;
        twist[2] = tz - (wx * ty - wy * tx) * 0.5;
        twist[3] = wx;
        twist[4] = wy;
        twist[5] = wz;
        return;
    }
    if (kPi - theta < 1e-6) {
        double xx = sqrt(fmax(0.0, 0.5 * (Dm[0] + 1.0)));
        double yy = sqrt(fmax(0.0, 0.5 * (Dm[5] + 1.0)));
        double zz = sqrt(fmax(0.0, 0.5 * (Dm[10] + 1.0)));
        if (xx >= yy && xx >= zz) {
            if (Dm[1] + Dm[4] < 0.0) yy = -yy;
            if (Dm[2] + Dm[8] < 0.0) zz = -zz;
            if (axx < 0.0) { xx = -xx; yy = -yy; zz = -zz; }
        } else if (yy >= zz) {
            if (Dm[1] + Dm[4] < 0.0) xx = -xx;
            if (Dm[6] + Dm[9] < 0.0) zz = -zz;
            if (axy < 0.0) { xx = -xx; yy = -yy; zz = -zz; }
        } else {
            if (Dm[2] + Dm[8] < 0.0) xx = -xx;
            if (Dm[6] + Dm[9] < 0.0) yy = -yy;
            if (axz < 0.0) { xx = -xx; yy = -yy; zz = -zz; }
        }
        safe_normal3(xx, yy, zz);
        wx = xx * theta;
        wy = yy * theta;
        wz = zz * theta;
    } else {
        const double f = theta / s2;
        wx = axx * f;
        wy = axy * f;
        wz = axz * f;
    }
    double st, ct;
    sincos(theta, &st, &ct);
    const double k = (1.0 - (theta * st) / (2.0 * (1.0 - ct))) / (theta * theta);
    const double wxtx = wy * tz - wz * ty, wxty = wz * tx - wx * tz, wxtz = wx * ty - wy * tx;
    const double wwx = wy * wxtz - wz * wxty, wwy = wz * wxtx - wx * wxtz, wwz = wx * wxty - wy * wxtx;
    twist[0] = tx - wxtx * 0.5 + wwx * k;
    twist[1] = ty - wxty * 0.5 + wwy * k;
    twist[2] = tz - wxtz * 0.5 + wwz * k;
    twist[3] = wx;
    twist[4] = wy;
    twist[5] = wz;
}

// TruncatedNormalUncertainVelocityActuator::GetControlValue (unc:70-75 noiseless, :77-90 noisy)
__device__ __forceinline__ double actuate(const DevAxis& ax, double u, bool noisy, double tn, unsigned& lflags) {
    if (isnan(u) || isinf(u)) lflags |= FKS_FLAG_WOULD_ASSERT_NAN;  // assert unc:72-73
    const double vl = ax.vlim;
    const double real_u = fmin(fmax(u, -vl), vl);
    if (!noisy) return real_u;
    const double pb = ax.pnoise * fabs(real_u);
    const double mb = ax.mnoise * vl;
    const double bound = fmax(pb, mb);
    return real_u + tn * bound;
}

// ------------------------------------------------------------------------------------------------
// forward kinematics: SetPosition (call sites spcs:875,1423-1424,1601).  cfg and T are shared memory.
// ------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void forward_kinematics(Ctx& c, double* cfg, double* T) {
    const int lane = c.lane;
    if (KIND == FKS_ROBOT_SE2) {
        if (lane == 0) {
            const double th = wrap_angle(cfg[2]);
            cfg[2] = th;
            double R[12];
            rot_axis(th, 0.0, 0.0, 1.0, R);
            R[3] = cfg[0];
            R[7] = cfg[1];
            R[11] = 0.0;
#pragma unroll
            for (int i = 0; i < 12; i++) T[i] = R[i];
        }
        __syncwarp();
    } else if (KIND == FKS_ROBOT_SE3) {
        if (cfg != T) {
            if (lane < 12) T[lane] = cfg[lane];
            __syncwarp();
        }
    } else {
        const DevRobot* rb = c.rb;
        double* M = c.ws + c.wl.M;
        // joint values (wrap / clamp) and joint motion matrices, one joint per lane
        for (int j = lane; j < c.J; j += 32) {
            const DevJoint& jd = rb->joints[j];
            if (jd.active >= 0) {
                double v = cfg[jd.active];
                if (jd.type == FKS_JOINT_CONTINUOUS) {
                    v = wrap_angle(v);
                } else {
                    if (v > jd.hi) v = jd.hi;
                    else if (v < jd.lo) v = jd.lo;
                }
                cfg[jd.active] = v;
                double R[12];
                if (jd.type == FKS_JOINT_PRISMATIC) {
#pragma unroll
                    for (int i = 0; i < 12; i++) R[i] = (i % 5 == 0) ? 1.0 : 0.0;
                    R[3] = jd.axis[0] * v;
                    R[7] = jd.axis[1] * v;
                    R[11] = jd.axis[2] * v;
                } else {
                    rot_axis(v, jd.axis[0], jd.axis[1], jd.axis[2], R);
                }
#pragma unroll
                for (int i = 0; i < 12; i++) M[12 * j + i] = R[i];
            }
        }
        if (lane < 12) T[lane] = rb->base[lane];
        __syncwarp();
        double* chain = c.ws + c.wl.chain;
        const int r = lane >> 2, cc = lane & 3;
        for (int j = 0; j < c.J; j++) {
            const DevJoint& jd = rb->joints[j];
            if (lane < 12) {
                const double* Tp = T + 12 * jd.parent;
                double v = Tp[4 * r + 0] * jd.T[cc] + Tp[4 * r + 1] * jd.T[4 + cc] + Tp[4 * r + 2] * jd.T[8 + cc];
                if (cc == 3) v += Tp[4 * r + 3];
                chain[lane] = v;
            }
            __syncwarp();
            if (lane < 12) {
                double v;
                if (jd.type == FKS_JOINT_FIXED) {
                    v = chain[lane];
                } else {
                    const double* Mj = M + 12 * j;
                    v = chain[4 * r + 0] * Mj[cc] + chain[4 * r + 1] * Mj[4 + cc] + chain[4 * r + 2] * Mj[8 + cc];
                    if (cc == 3) v += chain[4 * r + 3];
                }
                T[12 * jd.child + lane] = v;
            }
            __syncwarp();
        }
    }
}

// ApplyControlInput(input[, rng]) (tnuva:152-177 SE2, :348-382 SE3, :538-596 linked).
// Reads (cfg_in), writes (cfg_out, T_out); in and out may alias.  tn == nullptr: noiseless overload.
template <int KIND>
__device__ __forceinline__ void apply_control(Ctx& c, const double* cfg_in, double* cfg_out, double* T_out,
                                              const double* u, const double* tn) {
    const int lane = c.lane;
    if (KIND == FKS_ROBOT_SE3) {
        double* stepv = c.ws + c.wl.stepv;
        if (lane < 6) stepv[lane] = actuate(c.rb->axes[lane], u[lane], tn != nullptr, tn ? tn[lane] : 0.0, c.lflags);
        __syncwarp();
        double tw[6], E[12], A[12], Cm[12];
#pragma unroll
        for (int i = 0; i < 6; i++) tw[i] = stepv[i];
        exp_twist(tw, E);
#pragma unroll
        for (int i = 0; i < 12; i++) A[i] = cfg_in[i];
        iso_mul(A, E, Cm);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 12; i++) cfg_out[i] = Cm[i];
        }
        __syncwarp();
        (void)T_out;  // SE3: the configuration IS the link transform
    } else {
        if (lane < c.D) cfg_out[lane] = cfg_in[lane] + actuate(c.rb->axes[lane], u[lane], tn != nullptr, tn ? tn[lane] : 0.0, c.lflags);
        __syncwarp();
        forward_kinematics<KIND>(c, cfg_out, T_out);
    }
}

// ------------------------------------------------------------------------------------------------
// environment queries
// ------------------------------------------------------------------------------------------------
// VoxelGrid::LocationToGridIndex: grid-frame point * (1 / cell), C-cast truncation
__device__ __forceinline__ bool cell_index(const DevEnv& e, double wx, double wy, double wz, int& ix, int& iy, int& iz) {
    const double gx = e.inv_origin[0] * wx + e.inv_origin[1] * wy + e.inv_origin[2] * wz + e.inv_origin[3];
    const double gy = e.inv_origin[4] * wx + e.inv_origin[5] * wy + e.inv_origin[6] * wz + e.inv_origin[7];
    const double gz = e.inv_origin[8] * wx + e.inv_origin[9] * wy + e.inv_origin[10] * wz + e.inv_origin[11];
    ix = (int)(gx * e.inv_sdf_res);
    iy = (int)(gy * e.inv_sdf_res);
    iz = (int)(gz * e.inv_sdf_res);
    return ix >= 0 && iy >= 0 && iz >= 0 && ix < e.nx && iy < e.ny && iz < e.nz;
}
__device__ __forceinline__ float sdf_cell(const DevEnv& e, int x, int y, int z) {
    return __ldg(e.sdf + ((size_t)x * e.ny + y) * e.nz + z);
}

// SignedDistanceField::EstimateDistance4d for an in-bounds point whose cell (x,y,z) holds d0f
__device__ __forceinline__ double estimate_distance(const DevEnv& e, double wx, double wy, double wz, int x, int y, int z, float d0f) {
    const double res = e.sdf_res;
    const double d0 = (double)d0f;
    const double dc = (d0 >= 0.0) ? d0 - (res * 0.5) : d0 + (res * 0.5);
    double g0, g1, g2;
    if (x > 0 && y > 0 && z > 0 && x < e.nx - 1 && y < e.ny - 1 && z < e.nz - 1) {
        g0 = (double)(sdf_cell(e, x + 1, y, z) - sdf_cell(e, x - 1, y, z)) * e.inv_twice_res;
        g1 = (double)(sdf_cell(e, x, y + 1, z) - sdf_cell(e, x, y - 1, z)) * e.inv_twice_res;
        g2 = (double)(sdf_cell(e, x, y, z + 1) - sdf_cell(e, x, y, z - 1)) * e.inv_twice_res;
    } else {
        const int lx = max(0, x - 1), hx = min(e.nx - 1, x + 1);
        const int ly = max(0, y - 1), hy = min(e.ny - 1, y + 1);
        const int lz = max(0, z - 1), hz = min(e.nz - 1, z + 1);
        const double sx = (double)(hx - lx) * res, sy = (double)(hy - ly) * res, sz = (double)(hz - lz) * res;
        g0 = g1 = g2 = 0.0;
        if (sx > 0.0) g0 = ((double)sdf_cell(e, hx, y, z) - (double)sdf_cell(e, lx, y, z)) * (1.0 / sx);
        if (sy > 0.0) g1 = ((double)sdf_cell(e, x, hy, z) - (double)sdf_cell(e, x, ly, z)) * (1.0 / sy);
        if (sz > 0.0) g2 = ((double)sdf_cell(e, x, y, hz) - (double)sdf_cell(e, x, y, lz)) * (1.0 / sz);
    }
    const double cgx = res * ((double)x + 0.5), cgy = res * ((double)y + 0.5), cgz = res * ((double)z + 0.5);
    double cx, cy, cz;
    apply_T(e.origin, cgx, cgy, cgz, cx, cy, cz);
    const double vx = wx - cx, vy = wy - cy, vz = wz - cz;
    const double gg = g0 * g0 + g1 * g1 + g2 * g2;
    double adj = 0.0;
    if (gg > 0.0) adj = (vx * g0 + vy * g1 + vz * g2) / sqrt(gg);
    return dc + adj;
}

// CheckEnvironmentCollision (spcs:921-981) for the link transforms T; collision_threshold = 0.0 (spcs:424)
__device__ __forceinline__ bool check_env(const Ctx& c, const DevEnv& e, const DevSolver& sp, const double* T) {
    const double res = e.sdf_res;
    const double thr = 0.0 - (sp.check_tolerance * res);
    const double thr_deep = thr - res;
    bool hit = false;
#pragma unroll 2
    for (int p = c.lane; p < c.P; p += 32) {
        const double* Tl = T + 12 * c.plink[p];
        double wx, wy, wz;
        apply_T(Tl, c.px[p], c.py[p], c.pz[p], wx, wy, wz);
        int x, y, z;
        if (cell_index(e, wx, wy, wz, x, y, z)) {
            const float f = sdf_cell(e, x, y, z);
            if ((double)f < thr) {
                if ((double)f < thr_deep) hit = true;
                else if (estimate_distance(e, wx, wy, wz, x, y, z, f) < thr) hit = true;
            }
        }
        // out of bounds: the value is the oob value (+inf from the builder -> never a collision, spcs:943-955);
        // EstimateDistance4d out of bounds returns the same value, so both tiers reduce to one compare
        else if ((double)e.oob < thr) hit = true;
    }
    return __any_sync(FKS_FULL, hit);
}

// EstimateMaxControlInputWorkspaceMotion(start_robot, end_robot) (spcs:1492-1527)
__device__ __forceinline__ double max_motion(const Ctx& c, const double* Ta, const double* Tb) {
    double mx = 0.0;
#pragma unroll 2
    for (int p = c.lane; p < c.P; p += 32) {
        const int l = c.plink[p];
        const double x = c.px[p], y = c.py[p], z = c.pz[p];
        double ax, ay, az, bx, by, bz;
        apply_T(Ta + 12 * l, x, y, z, ax, ay, az);
        apply_T(Tb + 12 * l, x, y, z, bx, by, bz);
        const double dx = bx - ax, dy = by - ay, dz = bz - az;
        const double sq = dx * dx + dy * dy + dz * dz;
        if (sq > mx) mx = sq;
    }
    return sqrt(warp_max(mx));
}

// SurfaceNormalGrid::LookupSurfaceNormal + GetBestSurfaceNormal (spcs:186-198,235-256,111-132) for the
// in-bounds cell `li`; (dx,dy,dz) is the SafeNormal'd motion direction.
__device__ __forceinline__ void lookup_normal(const DevEnv& e, long long li, double dx, double dy, double dz,
                                              double& nx, double& ny, double& nz, unsigned& lflags) {
    nx = ny = nz = 0.0;
    const unsigned long long key = (unsigned long long)li + 1ull;
    unsigned long long h = normal_hash((unsigned long long)li) & e.nh_mask;
    uint2 range = make_uint2(0u, 0u);
    while (true) {
        const unsigned long long k = __ldg(e.nh_keys + h);
        if (k == key) {
            range = __ldg(e.nh_vals + h);
            break;
        }
        if (k == 0ull) break;
        h = (h + 1ull) & e.nh_mask;
    }
    if (range.y == 0u) return;  // empty cell -> zero normal
    const double dn = sqrt(dx * dx + dy * dy + dz * dz);
    double ux = 0.0, uy = 0.0, uz = 0.0;
    if (dn > 0.0) {
        ux = dx / dn;
        uy = dy / dn;
        uz = dz / dn;
    } else {
        lflags |= FKS_FLAG_WOULD_ASSERT_NORMAL;  // assert(direction_norm > 0.0) spcs:115
    }
    double best_dot = -INFINITY;
    unsigned best = range.x;
    for (unsigned en = range.x; en < range.x + range.y; en++) {
        const double* q = e.normal_entries + 6 * (size_t)en;
        const double dp = __ldg(q + 0) * ux + __ldg(q + 1) * uy + __ldg(q + 2) * uz;
        if (dp > best_dot) {
            best_dot = dp;
            best = en;
        }
    }
    const double* q = e.normal_entries + 6 * (size_t)best;
    nx = __ldg(q + 3);
    ny = __ldg(q + 4);
    nz = __ldg(q + 5);
}

// ------------------------------------------------------------------------------------------------
// self collisions: CollectSelfCollisions (spcs:1183-1275) + ExtractSelfCollidingPoints (spcs:983-1171)
//
// Two points share a cell of edge map_res only if they are within sqrt(3)*map_res of each other, so
// a bounding-sphere test over the DISALLOWED link pairs decides exactly when the hash-grid pass can
// be skipped (the common case).  The exact pass keeps the reference's per-cell semantics.
// Returns true when at least one point received a correction (self_collision_map non-empty);
// corrections are left in c.selfcorr, per-point flags in c.sflag (bit 0).
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ bool self_collisions_exact(Ctx& c, const DevEnv& e, const DevSolver& sp, const double* Tprev, const double* Tcur);

__device__ __forceinline__ bool collect_self(Ctx& c, const DevEnv& e, const DevSolver& sp, const double* Tprev, const double* Tcur) {
    const DevRobot* rb = c.rb;
    if (rb->n_pairs == 0) return false;  // one link, or every pair allowed (spcs:1186-1197)
    unsigned cand = 0u;
    const double diag = 1.7320508075688772 * e.map_res * (1.0 + 1e-6);
    for (int q = c.lane; q < rb->n_pairs; q += 32) {
        const int a = rb->pair_a[q], b = rb->pair_b[q];
        double ax, ay, az, bx, by, bz;
        apply_T(Tcur + 12 * a, rb->link_center[a][0], rb->link_center[a][1], rb->link_center[a][2], ax, ay, az);
        apply_T(Tcur + 12 * b, rb->link_center[b][0], rb->link_center[b][1], rb->link_center[b][2], bx, by, bz);
        const double dx = ax - bx, dy = ay - by, dz = az - bz;
        const double reach = rb->link_radius[a] + rb->link_radius[b] + diag;
        if (dx * dx + dy * dy + dz * dz <= reach * reach) cand |= (1u << a) | (1u << b);
    }
    cand = __reduce_or_sync(FKS_FULL, cand);
    if (cand == 0u) return false;
    c.cand_links = cand;
    return self_collisions_exact(c, e, sp, Tprev, Tcur);
}

__device__ __forceinline__ bool key_eq(const int* keys, int a, int b) {
    return keys[3 * a] == keys[3 * b] && keys[3 * a + 1] == keys[3 * b + 1] && keys[3 * a + 2] == keys[3 * b + 2];
}

// scan the points of `link` for members of the cell of point `ref`: count, first member, and the
// momentum sum (point velocities added in point order, spcs:1040-1054)
__device__ __forceinline__ void scan_link_cell(const Ctx& c, const double* Tprev, const double* Tcur, int link, int ref,
                                               double time_multiplier, int& count, int& first, double& mx, double& my, double& mz) {
    const DevRobot* rb = c.rb;
    count = 0;
    first = -1;
    mx = my = mz = 0.0;
    const int begin = rb->link_begin[link], end = rb->link_begin[link + 1];
    for (int base = begin; base < end; base += 32) {
        const int q = base + c.lane;
        const bool match = (q < end) && key_eq(c.keys, q, ref);
        double vx = 0.0, vy = 0.0, vz = 0.0;
        if (match) {
            double ax, ay, az, bx, by, bz;
            apply_T(Tcur + 12 * link, c.px[q], c.py[q], c.pz[q], ax, ay, az);
            apply_T(Tprev + 12 * link, c.px[q], c.py[q], c.pz[q], bx, by, bz);
            vx = (ax - bx) * time_multiplier;
            vy = (ay - by) * time_multiplier;
            vz = (az - bz) * time_multiplier;
        }
        unsigned mm = __ballot_sync(FKS_FULL, match);
        if (mm && first < 0) first = base + (__ffs(mm) - 1);
        count += __popc(mm);
        while (mm) {
            const int b = __ffs(mm) - 1;
            mm &= mm - 1u;
            mx += __shfl_sync(FKS_FULL, vx, b);
            my += __shfl_sync(FKS_FULL, vy, b);
            mz += __shfl_sync(FKS_FULL, vz, b);
        }
    }
}

__device__ __noinline__ bool self_collisions_exact(Ctx& c, const DevEnv& e, const DevSolver& sp, const double* Tprev, const double* Tcur) {
    const DevRobot* rb = c.rb;
    const unsigned cand = c.cand_links;
    const int lane = c.lane;
    // cell keys: LocationToExtendedGridIndex (spcs:1173-1181) DIVIDES by the map resolution
    for (int p = lane; p < c.P; p += 32) {
        const int l = c.plink[p];
        int kx = 0, ky = 0, kz = 0;
        if ((cand >> l) & 1u) {
            double wx, wy, wz;
            apply_T(Tcur + 12 * l, c.px[p], c.py[p], c.pz[p], wx, wy, wz);
            const double gx = e.inv_origin[0] * wx + e.inv_origin[1] * wy + e.inv_origin[2] * wz + e.inv_origin[3];
            const double gy = e.inv_origin[4] * wx + e.inv_origin[5] * wy + e.inv_origin[6] * wz + e.inv_origin[7];
            const double gz = e.inv_origin[8] * wx + e.inv_origin[9] * wy + e.inv_origin[10] * wz + e.inv_origin[11];
            kx = (int)(gx / e.map_res);
            ky = (int)(gy / e.map_res);
            kz = (int)(gz / e.map_res);
        }
        c.keys[3 * p] = kx;
        c.keys[3 * p + 1] = ky;
        c.keys[3 * p + 2] = kz;
    }
    __syncwarp();
    // per point: does its cell hold a point of a link it may not touch?  is it the first of its link there?
    bool any = false;
    for (int base = 0; base < c.P; base += 32) {
        const int p = base + lane;
        unsigned char flag = 0;
        if (p < c.P) {
            const int l = c.plink[p];
            if ((cand >> l) & 1u) {
                unsigned dis = rb->disallowed[l] & cand;
                bool collides = false;
                while (dis && !collides) {
                    const int m = __ffs(dis) - 1;
                    dis &= dis - 1u;
                    for (int q = rb->link_begin[m]; q < rb->link_begin[m + 1]; q++)
                        if (key_eq(c.keys, q, p)) {
                            collides = true;
                            break;
                        }
                }
                if (collides) {
                    bool leader = true;
                    for (int q = rb->link_begin[l]; q < p; q++)
                        if (key_eq(c.keys, q, p)) {
                            leader = false;
                            break;
                        }
                    flag = leader ? 3 : 1;
                }
            }
            c.sflag[p] = flag;
        }
        any = any || __any_sync(FKS_FULL, flag != 0);
    }
    __syncwarp();
    if (!any) return false;
    const double time_multiplier = 1.0 / sp.interval;
    double* sw = c.selfwork;
    // one (cell, link) group at a time, the whole warp working on it
    for (int base = 0; base < c.P; base += 32) {
        const int p = base + lane;
        unsigned leaders = __ballot_sync(FKS_FULL, (p < c.P) && (c.sflag[p] == 3));
        while (leaders) {
            const int gp = base + (__ffs(leaders) - 1);
            leaders &= leaders - 1u;
            const int l = c.plink[gp];
            int cnt_i, first_i;
            double mix, miy, miz;
            scan_link_cell(c, Tprev, Tcur, l, gp, time_multiplier, cnt_i, first_i, mix, miy, miz);
            double aix, aiy, aiz;
            apply_T(Tprev + 12 * l, c.px[first_i], c.py[first_i], c.pz[first_i], aix, aiy, aiz);
            const double inv_i = 1.0 / (double)cnt_i;
            const double vix = mix * inv_i, viy = miy * inv_i, viz = miz * inv_i;
            // colliding links in ascending order (std::map iteration, spcs:1000-1017)
            int m = 0;
            unsigned dis = rb->disallowed[l] & cand;
            while (dis) {
                const int s = __ffs(dis) - 1;
                dis &= dis - 1u;
                int cnt_s, first_s;
                double msx, msy, msz;
                scan_link_cell(c, Tprev, Tcur, s, gp, time_multiplier, cnt_s, first_s, msx, msy, msz);
                if (cnt_s == 0) continue;
                double ox, oy, oz;
                apply_T(Tprev + 12 * s, c.px[first_s], c.py[first_s], c.pz[first_s], ox, oy, oz);
                double nx = ox - aix, ny = oy - aiy, nz = oz - aiz;
                safe_normal3(nx, ny, nz);
                const double inv_s = 1.0 / (double)cnt_s;
                const double vsx = msx * inv_s, vsy = msy * inv_s, vsz = msz * inv_s;
                const double rhs = nx * (vsx - vix) + ny * (vsy - viy) + nz * (vsz - viz);
                if (lane == 0) {
                    sw[5 * m + 0] = nx;
                    sw[5 * m + 1] = ny;
                    sw[5 * m + 2] = nz;
                    sw[5 * m + 3] = rhs;
                    sw[5 * m + 4] = rb->link_mass[s];
                }
                m++;
            }
            double* out = sw + 5 * kMaxSelfPartners;  // per-point correction (3)
            __syncwarp();
            if (lane == 0) {
                // A = N^T C^T M^-1 C N (spcs:1135), inverse by partial-pivot Gauss elimination, lambda = A^-1 r
                double* Mx = out + 4;                                  // m x 2m augmented
                double* Ainv = Mx + kMaxSelfPartners * 2 * kMaxSelfPartners;  // m x m
                const double mass_i = rb->link_mass[l];
                const int n = m;
                for (int a = 0; a < n; a++)
                    for (int b = 0; b < n; b++) {
                        double v = (sw[5 * a] * sw[5 * b] + sw[5 * a + 1] * sw[5 * b + 1] + sw[5 * a + 2] * sw[5 * b + 2]) / mass_i;
                        if (a == b) v += (sw[5 * a] * sw[5 * a] + sw[5 * a + 1] * sw[5 * a + 1] + sw[5 * a + 2] * sw[5 * a + 2]) / sw[5 * a + 4];
                        Mx[a * 2 * n + b] = v;
                        Mx[a * 2 * n + n + b] = (a == b) ? 1.0 : 0.0;
                    }
                for (int k = 0; k < n; k++) {
                    int piv = k;
                    double best = fabs(Mx[k * 2 * n + k]);
                    for (int r = k + 1; r < n; r++)
                        if (fabs(Mx[r * 2 * n + k]) > best) {
                            best = fabs(Mx[r * 2 * n + k]);
                            piv = r;
                        }
                    if (piv != k)
                        for (int cc = 0; cc < 2 * n; cc++) {
                            const double t = Mx[k * 2 * n + cc];
                            Mx[k * 2 * n + cc] = Mx[piv * 2 * n + cc];
                            Mx[piv * 2 * n + cc] = t;
                        }
                    const double pv = Mx[k * 2 * n + k];
                    for (int r = k + 1; r < n; r++) {
                        const double f = Mx[r * 2 * n + k] / pv;
                        for (int cc = k; cc < 2 * n; cc++) Mx[r * 2 * n + cc] -= f * Mx[k * 2 * n + cc];
                    }
                }
                for (int cc = 0; cc < n; cc++)
                    for (int r = n - 1; r >= 0; r--) {
                        double s = Mx[r * 2 * n + n + cc];
                        for (int j = r + 1; j < n; j++) s -= Mx[r * 2 * n + j] * Ainv[j * n + cc];
                        Ainv[r * n + cc] = s / Mx[r * 2 * n + r];
                    }
                double cx = 0.0, cy = 0.0, cz = 0.0;
                for (int a = 0; a < n; a++) {
                    double lam = 0.0;
                    for (int b = 0; b < n; b++) lam += Ainv[a * n + b] * sw[5 * b + 3];
                    cx = cx + sw[5 * a] * lam;
                    cy = cy + sw[5 * a + 1] * lam;
                    cz = cz + sw[5 * a + 2] * lam;
                }
                const double im = 1.0 / mass_i;
                cx = cx * im;
                cy = cy * im;
                cz = cz * im;
                out[0] = cx * inv_i;
                out[1] = cy * inv_i;
                out[2] = cz * inv_i;
            }
            __syncwarp();
            const double ppx = out[0], ppy = out[1], ppz = out[2];
            if (isnan(ppx) || isnan(ppy) || isnan(ppz)) c.lflags |= FKS_FLAG_WOULD_ASSERT_NAN;  // asserts spcs:1151-1153
            for (int q = rb->link_begin[l] + lane; q < rb->link_begin[l + 1]; q += 32)
                if (key_eq(c.keys, q, gp)) {
                    c.selfcorr[3 * q] = ppx;
                    c.selfcorr[3 * q + 1] = ppy;
                    c.selfcorr[3 * q + 2] = ppz;
                }
            __syncwarp();
        }
    }
    return true;
}

// CheckCollision (spcs:1418-1436)
template <int KIND>
__device__ __forceinline__ bool check_collision(Ctx& c, const DevEnv& e, const DevSolver& sp, const double* Tprev, const double* Tcur, bool& has_self) {
    const bool envc = check_env(c, e, sp, Tcur);
    has_self = (KIND == FKS_ROBOT_LINKED) ? collect_self(c, e, sp, Tprev, Tcur) : false;
    return envc || has_self;
}

// ------------------------------------------------------------------------------------------------
// CollectPointCorrectionsAndJacobians (spcs:1818-1939): fills the stacked Jacobian (column major,
// leading dimension c.ldj, column D holds the corrections), rows in link-major point-minor order.
// Returns the number of rows.
// ------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ int collect_corrections(Ctx& c, const DevEnv& e, const double* Tprev, const double* Tcur, bool has_self) {
    const int lane = c.lane;
    const DevRobot* rb = c.rb;
    double* A = c.Js;
    const int ld = c.ldj;
    const int D = c.D;
    double* jaxis = c.ws + c.wl.jaxis;
    double* jorig = c.ws + c.wl.jorig;
    if (KIND == FKS_ROBOT_LINKED) {
        // world axis / origin of every joint (frame = child link transform)
        for (int j = lane; j < c.J; j += 32) {
            const DevJoint& jd = rb->joints[j];
            const double* Tj = Tcur + 12 * jd.child;
            jaxis[3 * j + 0] = Tj[0] * jd.axis[0] + Tj[1] * jd.axis[1] + Tj[2] * jd.axis[2];
            jaxis[3 * j + 1] = Tj[4] * jd.axis[0] + Tj[5] * jd.axis[1] + Tj[6] * jd.axis[2];
            jaxis[3 * j + 2] = Tj[8] * jd.axis[0] + Tj[9] * jd.axis[1] + Tj[10] * jd.axis[2];
            jorig[3 * j + 0] = Tj[3];
            jorig[3 * j + 1] = Tj[7];
            jorig[3 * j + 2] = Tj[11];
        }
        __syncwarp();
    }
    const double res = e.sdf_res;
    int npts = 0;
    for (int base = 0; base < c.P; base += 32) {
        const int p = base + lane;
        bool have = false;
        double cx = 0.0, cy = 0.0, cz = 0.0;
        double wx = 0.0, wy = 0.0, wz = 0.0, lx = 0.0, ly = 0.0, lz = 0.0;
        int l = 0;
        if (p < c.P) {
            l = c.plink[p];
            lx = c.px[p];
            ly = c.py[p];
            lz = c.pz[p];
            if (has_self && (c.sflag[p] & 1)) {
                have = true;
                cx = cx + c.selfcorr[3 * p];
                cy = cy + c.selfcorr[3 * p + 1];
                cz = cz + c.selfcorr[3 * p + 2];
            }
            apply_T(Tcur + 12 * l, lx, ly, lz, wx, wy, wz);
            int x, y, z;
            if (cell_index(e, wx, wy, wz, x, y, z)) {
                const float f = sdf_cell(e, x, y, z);
                // EstimateDistance = f -/+ res/2 + (|.| <= sqrt(3)/2 res): cannot be negative when f >= 2 res
                if ((double)f < 2.0 * res) {
                    const double est = estimate_distance(e, wx, wy, wz, x, y, z, f);
                    if (est < 0.0) {  // resolution_distance_threshold_ = 0.0 (spcs:425,1874)
                        double qx, qy, qz;
                        apply_T(Tprev + 12 * l, lx, ly, lz, qx, qy, qz);
                        double dx = wx - qx, dy = wy - qy, dz = wz - qz;
                        safe_normal3(dx, dy, dz);
                        double nx, ny, nz;
                        lookup_normal(e, ((long long)x * e.ny + y) * e.nz + z, dx, dy, dz, nx, ny, nz, c.lflags);
                        safe_normal3(nx, ny, nz);
                        const double pen = fabs(0.0 - est);
                        cx = cx + nx * pen;
                        cy = cy + ny * pen;
                        cz = cz + nz * pen;
                        have = true;
                    }
                }
            }
        }
        const unsigned mask = __ballot_sync(FKS_FULL, have);
        if (have) {
            const int row0 = 3 * (npts + __popc(mask & ((1u << lane) - 1u)));
            double* b = A + (size_t)D * ld + row0;
            b[0] = cx;
            b[1] = cy;
            b[2] = cz;
            // ComputeLinkPointTranslationJacobian (3 x D)
            if (KIND == FKS_ROBOT_SE2) {
                const double* T = Tcur;
                const double rx = wx - T[3], ry = wy - T[7], rz = wz - 0.0;
                (void)rz;
                double* a0 = A + row0;
                a0[0] = 1.0; a0[1] = 0.0; a0[2] = 0.0;
                double* a1 = A + ld + row0;
                a1[0] = 0.0; a1[1] = 1.0; a1[2] = 0.0;
                double* a2 = A + 2 * ld + row0;  // z x (p_world - (x, y, 0))
                a2[0] = 0.0 * rz - 1.0 * ry;
                a2[1] = 1.0 * rx - 0.0 * rz;
                a2[2] = 0.0 * ry - 0.0 * rx;
            } else if (KIND == FKS_ROBOT_SE3) {
                const double* T = Tcur;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const double t0 = T[4 * r + 0], t1 = T[4 * r + 1], t2 = T[4 * r + 2];
                    A[0 * ld + row0 + r] = t0;
                    A[1 * ld + row0 + r] = t1;
                    A[2 * ld + row0 + r] = t2;
                    A[3 * ld + row0 + r] = t1 * (-lz) + t2 * ly;
                    A[4 * ld + row0 + r] = t0 * lz + t2 * (-lx);
                    A[5 * ld + row0 + r] = t0 * (-ly) + t1 * lx;
                }
            } else {
                const unsigned anc = rb->link_ancestors[l];
                for (int a = 0; a < D; a++) {
                    const int j = rb->active_joint[a];
                    double jx = 0.0, jy = 0.0, jz = 0.0;
                    if ((anc >> j) & 1u) {
                        const double ax = jaxis[3 * j], ay = jaxis[3 * j + 1], az = jaxis[3 * j + 2];
                        if (rb->joints[j].type == FKS_JOINT_PRISMATIC) {
                            jx = ax; jy = ay; jz = az;
                        } else {
                            const double rx = wx - jorig[3 * j], ry = wy - jorig[3 * j + 1], rz = wz - jorig[3 * j + 2];
                            jx = ay * rz - az * ry;
                            jy = az * rx - ax * rz;
                            jz = ax * ry - ay * rx;
                        }
                    }
                    double* col = A + (size_t)a * ld + row0;
                    col[0] = jx;
                    col[1] = jy;
                    col[2] = jz;
                }
            }
        }
        npts += __popc(mask);
    }
    __syncwarp();
    return 3 * npts;
}

// ------------------------------------------------------------------------------------------------
// ComputeResolverCorrectionStepStackedJacobian (spcs:1990-1998): x = J.colPivHouseholderQr().solve(c),
// Eigen 3.3 semantics.  A is rows x cols column major (leading dimension ld), b = column
// `cols` of the same store.  Lanes stride over rows.  Result in x (shared memory, cols entries).
// ------------------------------------------------------------------------------------------------
__device__ void colpiv_qr_solve(Ctx& c, double* A, int ld, int rows, int cols, double* x) {
    const int lane = c.lane;
    double* nu = c.ws + c.wl.qr;  // norms updated
    double* nd = nu + cols;       // norms direct
    double* hc = nd + cols;       // householder coefficients
    int* transp = (int*)(hc + cols);
    double* b = A + (size_t)cols * ld;
    const int size = rows < cols ? rows : cols;
    double max_norm = 0.0;
    for (int k = 0; k < cols; k++) {
        const double* ck = A + (size_t)k * ld;
        double s = 0.0;
        for (int r = lane; r < rows; r += 32) s += ck[r] * ck[r];
        const double n = sqrt(warp_sum(s));
        if (lane == 0) {
            nd[k] = n;
            nu[k] = n;
        }
        max_norm = fmax(max_norm, n);
    }
    __syncwarp();
    const double eps = DBL_EPSILON;
    const double threshold_helper = ((max_norm * eps) * (max_norm * eps)) / (double)rows;
    const double norm_downdate_threshold = sqrt(eps);
    int nonzero_pivots = size;
    for (int k = 0; k < size; k++) {
        int biggest = k;
        double big = nu[k];
        for (int j = k + 1; j < cols; j++) {
            const double v = nu[j];
            if (v > big) {
                big = v;
                biggest = j;
            }
        }
        const double big_sq = big * big;
        const double cut = threshold_helper * (double)(rows - k);
        if (nonzero_pivots == size && big_sq < cut) nonzero_pivots = k;
        if (max_norm > 0.0 && big_sq > 0.0 && big_sq < cut * 1e6) c.lflags |= FKS_FLAG_NEAR_RANK_CUT;
        __syncwarp();
        double* ck = A + (size_t)k * ld;
        if (k != biggest) {
            double* cb = A + (size_t)biggest * ld;
            for (int r = lane; r < rows; r += 32) {
                const double t = ck[r];
                ck[r] = cb[r];
                cb[r] = t;
            }
            if (lane == 0) {
                double t = nu[k]; nu[k] = nu[biggest]; nu[biggest] = t;
                t = nd[k]; nd[k] = nd[biggest]; nd[biggest] = t;
            }
        }
        if (lane == 0) transp[k] = biggest;
        __syncwarp();
        // makeHouseholderInPlace on col(k).tail(rows - k)
        double ts = 0.0;
        for (int r = k + 1 + lane; r < rows; r += 32) ts += ck[r] * ck[r];
        const double tail_sq = warp_sum(ts);
        const double c0 = ck[k];
        double tau, beta;
        __syncwarp();
        if (tail_sq <= DBL_MIN) {
            tau = 0.0;
            beta = c0;
            for (int r = k + 1 + lane; r < rows; r += 32) ck[r] = 0.0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0) beta = -beta;
            const double denom = c0 - beta;
            for (int r = k + 1 + lane; r < rows; r += 32) ck[r] = ck[r] / denom;
            tau = (beta - c0) / beta;
        }
        if (lane == 0) {
            hc[k] = tau;
            ck[k] = beta;
        }
        __syncwarp();
        // applyHouseholderOnTheLeft to the trailing columns
        if (rows - k == 1) {
            if (lane == 0)
                for (int j = k + 1; j < cols; j++) A[(size_t)j * ld + k] *= (1.0 - tau);
        } else if (tau != 0.0) {
            for (int j = k + 1; j < cols; j++) {
                double* cj = A + (size_t)j * ld;
                double t = 0.0;
                for (int r = k + 1 + lane; r < rows; r += 32) t += ck[r] * cj[r];
                const double tmp = warp_sum(t) + cj[k];
                __syncwarp();
                if (lane == 0) cj[k] -= tau * tmp;
                for (int r = k + 1 + lane; r < rows; r += 32) cj[r] -= (tau * ck[r]) * tmp;
            }
        }
        __syncwarp();
        // LAPACK-style norm downdate
        for (int j = k + 1; j < cols; j++) {
            const double nuj = nu[j];
            if (nuj != 0.0) {
                const double* cj = A + (size_t)j * ld;
                double temp = fabs(cj[k]) / nuj;
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double ratio = nuj / nd[j];
                const double temp2 = temp * (ratio * ratio);
                double newu, newd = nd[j];
                if (temp2 <= norm_downdate_threshold) {
                    double s = 0.0;
                    for (int r = k + 1 + lane; r < rows; r += 32) s += cj[r] * cj[r];
                    newd = sqrt(warp_sum(s));
                    newu = newd;
                } else {
                    newu = nuj * sqrt(temp);
                }
                __syncwarp();
                if (lane == 0) {
                    nu[j] = newu;
                    nd[j] = newd;
                }
            }
        }
        __syncwarp();
    }
    // c = H_{nz-1} ... H_0 b
    for (int k = 0; k < nonzero_pivots; k++) {
        const double tau = hc[k];
        const double* ck = A + (size_t)k * ld;
        if (rows - k == 1) {
            if (lane == 0) b[k] *= (1.0 - tau);
        } else if (tau != 0.0) {
            double t = 0.0;
            for (int r = k + 1 + lane; r < rows; r += 32) t += ck[r] * b[r];
            const double tmp = warp_sum(t) + b[k];
            __syncwarp();
            if (lane == 0) b[k] -= tau * tmp;
            for (int r = k + 1 + lane; r < rows; r += 32) b[r] -= (tau * ck[r]) * tmp;
        }
        __syncwarp();
    }
    if (lane < cols) x[lane] = 0.0;
    __syncwarp();
    if (lane == 0 && nonzero_pivots > 0) {
        // back substitution on the leading nz x nz upper triangle, then un-permute
        for (int i = nonzero_pivots - 1; i >= 0; i--) {
            double s = b[i];
            for (int j = i + 1; j < nonzero_pivots; j++) s -= A[(size_t)j * ld + i] * b[j];
            b[i] = s / A[(size_t)i * ld + i];
        }
        int perm[kMaxDof];
        for (int j = 0; j < cols; j++) perm[j] = j;
        for (int k = 0; k < size; k++) {
            const int t = perm[k];
            perm[k] = perm[transp[k]];
            perm[transp[k]] = t;
        }
        for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = b[i];
    }
    __syncwarp();
}

struct Stats {
    unsigned long long v[FKS_NUM_STATS];
};

// EstimateMaxControlInputWorkspaceMotion(robot, control_input) (spcs:1538-1544): noiseless apply on a copy
template <int KIND>
__device__ __forceinline__ double max_motion_of_input(Ctx& c, const double* u) {
    double* cfg = c.ws + c.wl.cfg;
    double* tcfg = c.ws + c.wl.tcfg;
    double* Tcur = c.ws + c.wl.Tcur;
    double* Ttmp = c.ws + c.wl.Ttmp;
    apply_control<KIND>(c, cfg, tcfg, Ttmp, u, nullptr);
    return max_motion(c, Tcur, Ttmp);
}

template <int KIND>
__device__ __forceinline__ void copy_state(Ctx& c, int cfg_from, int T_from, int cfg_to, int T_to) {
    if (KIND != FKS_ROBOT_SE3)
        if (c.lane < c.stride) c.ws[cfg_to + c.lane] = c.ws[cfg_from + c.lane];
    for (int i = c.lane; i < c.L * 12; i += 32) c.ws[T_to + i] = c.ws[T_from + i];
    __syncwarp();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kThreadsPerBlock) simulate_kernel(const __grid_constant__ LaunchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- stage the robot description and its points into shared memory, once per CTA -------------
    DevRobot* rb = reinterpret_cast<DevRobot*>(smem_raw);
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(a.robot);
        unsigned* dst = reinterpret_cast<unsigned*>(rb);
        for (int i = threadIdx.x; i < (int)(sizeof(DevRobot) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const int P = a.robot->P;
    size_t off = (sizeof(DevRobot) + 15) & ~(size_t)15;
    double* spx = reinterpret_cast<double*>(smem_raw + off);
    double* spy = spx + P;
    double* spz = spy + P;
    int* splink = reinterpret_cast<int*>(spz + P);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        spx[i] = a.px[i];
        spy[i] = a.py[i];
        spz[i] = a.pz[i];
        splink[i] = a.plink[i];
    }
    off += (size_t)P * 24 + (((size_t)P * 4 + 15) & ~(size_t)15);
    __syncthreads();

    Ctx c;
    c.rb = rb;
    c.px = spx;
    c.py = spy;
    c.pz = spz;
    c.plink = splink;
    c.lane = threadIdx.x & 31;
    c.L = rb->L;
    c.J = rb->J;
    c.D = rb->D;
    c.P = P;
    c.stride = a.cfg_stride;
    c.wl = make_warp_layout(KIND, c.L, c.J, c.D, c.stride);
    const int warp_in_block = threadIdx.x >> 5;
    c.ws = reinterpret_cast<double*>(smem_raw + off) + (size_t)warp_in_block * c.wl.total;
    const ScratchLayout sl = make_scratch_layout(c.D, P);
    char* slot = a.scratch + (size_t)(blockIdx.x * kWarpsPerBlock + warp_in_block) * a.scratch_bytes_per_warp;
    c.Js = reinterpret_cast<double*>(slot + sl.jstore);
    c.ldj = sl.ldj;
    c.selfcorr = reinterpret_cast<double*>(slot + sl.selfcorr);
    c.selfwork = reinterpret_cast<double*>(slot + sl.selfwork);
    c.keys = reinterpret_cast<int*>(slot + sl.keys);
    c.sflag = reinterpret_cast<unsigned char*>(slot + sl.sflag);
    c.cand_links = 0u;

    const int lane = c.lane;
    const int D = c.D;
    const DevEnv& e = a.env;
    const DevSolver& sp = a.sp;
    double* cfg = c.ws + c.wl.cfg;
    double* Tprev = c.ws + c.wl.Tprev;
    double* Tcur = c.ws + c.wl.Tcur;
    double* target = c.ws + c.wl.target;
    double* scfg = c.ws + c.wl.scfg;
    double* act = c.ws + c.wl.act;
    double* ru = c.ws + c.wl.ru;
    double* du = c.ws + c.wl.du;
    double* tn = c.ws + c.wl.tn;
    double* raw = c.ws + c.wl.raw;
    double* stepv = c.ws + c.wl.stepv;

    Stats st;
#pragma unroll
    for (int i = 0; i < FKS_NUM_STATS; i++) st.v[i] = 0ull;

    while (true) {
        unsigned long long pid = 0ull;
        if (lane == 0) pid = (unsigned long long)atomicAdd(a.counter, 1u);
        pid = __shfl_sync(FKS_FULL, pid, 0);
        if (pid >= a.n_particles) break;

        // ---- ForwardSimulateRobot (spcs:824-829): clone + ResetPosition(start) -----------------------
        c.lflags = 0u;
        const double* start = a.starts + (size_t)pid * c.stride;
        const double* tgt = a.targets + (a.n_targets == a.n_particles ? (size_t)pid * c.stride : 0);
        if (lane < c.stride) {
            cfg[lane] = start[lane];
            target[lane] = tgt[lane];
        }
        __syncwarp();
        forward_kinematics<KIND>(c, cfg, Tcur);
        double pid_integral = 0.0, pid_last_error = 0.0;  // lane i owns axis i (pid:98-102 zeroed)
        unsigned long long tape_pos = 0ull, tape_end = 0ull;
        if (a.noise_mode == FKS_NOISE_INJECTED) {
            tape_pos = a.tape_off[pid];
            tape_end = a.tape_off[pid + 1];
        }
        bool collided = false, any_resolve_failed = false;
        unsigned flags = 0u, n_micro_total = 0u, n_iter_total = 0u, n_steps = 0u;

        for (unsigned step = 0; step < sp.n_steps; step++) {
            n_steps++;
            if (!a.allow_contacts) {  // a colliding step is discarded as a whole: keep the step's start (spcs:904-909)
                if (lane < c.stride) scfg[lane] = cfg[lane];
                __syncwarp();
            }
            // ---- GenerateControlAction (tnuva:179-198, :384-412, :598-614) ---------------------------
            if (KIND == FKS_ROBOT_SE3) {
                if (lane == 0) {
                    double cur[12], tg[12], tw[6];
#pragma unroll
                    for (int i = 0; i < 12; i++) {
                        cur[i] = cfg[i];
                        tg[i] = target[i];
                    }
                    twist_between(cur, tg, tw);
#pragma unroll
                    for (int i = 0; i < 6; i++) stepv[i] = tw[i];
                }
                __syncwarp();
            }
            if (lane < D) {
                double err;
                if (KIND == FKS_ROBOT_SE2) {
                    err = target[lane] - cfg[lane];
                    if (lane == 2) err = wrap_angle(err);
                } else if (KIND == FKS_ROBOT_SE3) {
                    err = stepv[lane];
                } else {
                    err = target[lane] - cfg[lane];
                    if (rb->joints[rb->active_joint[lane]].type == FKS_JOINT_CONTINUOUS) err = wrap_angle(err);
                }
                // SimplePIDController::ComputeFeedbackTerm (pid:122-135)
                const DevAxis& ax = rb->axes[lane];
                const double dt = sp.interval;
                const double timestep_error_integral = ((err * 0.5) + (pid_last_error * 0.5)) * dt;
                const double new_error_integral = pid_integral + timestep_error_integral;
                pid_integral = fmax(-ax.iclamp, fmin(ax.iclamp, new_error_integral));
                const double error_derivative = (err - pid_last_error) / dt;
                pid_last_error = err;
                const double term = (err * ax.kp) + (pid_integral * ax.ki) + (error_derivative * ax.kd);
                const double action = actuate(ax, term, false, 0.0, c.lflags);
                act[lane] = action;
                ru[lane] = action * sp.interval;  // real_control_input (spcs:1549)
            }
            __syncwarp();

            // ---- ResolveForwardSimulation (spcs:1546-1816) -------------------------------------------
            const double computed_step_motion = max_motion_of_input<KIND>(c, ru);
            const double target_microstep_distance = e.map_res * 0.125;
            const double allowed_microstep_distance = e.map_res * 1.0;
            const double ratio = computed_step_motion / target_microstep_distance;
            unsigned number_microsteps = (unsigned)ceil(ratio);
            if (number_microsteps < 1u) number_microsteps = 1u;
            if (lane < D) du[lane] = ru[lane] / (double)number_microsteps;
            __syncwarp();
            const double computed_microstep_motion = max_motion_of_input<KIND>(c, du);
            if (computed_microstep_motion > allowed_microstep_distance) flags |= FKS_FLAG_WOULD_ASSERT_MICROSTEP;

            bool step_collided = false, step_failed = false, step_stopped = false;
            bool has_self = false;
            for (unsigned micro = 0; micro < number_microsteps; micro++) {
                n_micro_total++;
                copy_state<KIND>(c, c.wl.cfg, c.wl.Tcur, c.wl.pcfg, c.wl.Tprev);  // previous_configuration (spcs:1597)
                // actuator noise: one truncated-normal draw per axis, in axis order
                if (lane < D) {
                    double v = 0.0;
                    if (a.noise_mode == FKS_NOISE_INJECTED) {
                        if (tape_pos + lane < tape_end) v = a.tape[tape_pos + lane];
                        else c.lflags |= FKS_FLAG_TAPE_EXHAUSTED;
                    } else if (a.noise_mode == FKS_NOISE_PHILOX) {
                        v = fks_philox_truncated_normal(a.seed, a.first_id + pid, step, micro, (uint32_t)lane, rb->axes[lane].sigma);
                    }
                    tn[lane] = v;
                }
                tape_pos += (unsigned long long)D;
                __syncwarp();
                apply_control<KIND>(c, cfg, cfg, Tcur, du, tn);  // spcs:1599-1601
                bool in_collision = check_collision<KIND>(c, e, sp, Tprev, Tcur, has_self);  // spcs:1608
                if (in_collision) step_collided = true;
                if (in_collision && a.allow_contacts) {
                    unsigned resolver_iterations = 0u;
                    double scaling = sp.initial_step;
                    while (in_collision) {
                        const int rows = collect_corrections<KIND>(c, e, Tprev, Tcur, has_self);  // spcs:1627
                        st.v[FKS_STAT_TOTAL_CORRECTED_POINTS] += (unsigned long long)(rows / 3);
                        if (rows == 0) {
                            // Eigen would return an empty vector and ApplyControlInput would assert; documented
                            // device behaviour: zero correction step
                            flags |= FKS_FLAG_EMPTY_JACOBIAN;
                            if (lane < D) raw[lane] = 0.0;
                            __syncwarp();
                        } else {
                            colpiv_qr_solve(c, c.Js, c.ldj, rows, D, raw);  // spcs:1629,1990-1998
                        }
                        const double est = max_motion_of_input<KIND>(c, raw);  // spcs:1630
                        const double step_fraction = fmax(est / allowed_microstep_distance, 1.0);  // spcs:1681
                        if (lane < D) stepv[lane] = (raw[lane] / step_fraction) * fabs(scaling);  // spcs:1682
                        __syncwarp();
                        if (KIND == FKS_ROBOT_SE3) {
                            // apply_control<SE3> uses stepv as its own temporary: hand the step over in `act`
                            // (the controller action of this step is no longer needed)
                            if (lane < D) act[lane] = stepv[lane];
                            __syncwarp();
                            apply_control<KIND>(c, cfg, cfg, Tcur, act, nullptr);  // spcs:1689
                        } else {
                            apply_control<KIND>(c, cfg, cfg, Tcur, stepv, nullptr);
                        }
                        in_collision = check_collision<KIND>(c, e, sp, Tprev, Tcur, has_self);  // spcs:1694-1698
                        resolver_iterations++;
                        n_iter_total++;
                        if (resolver_iterations > sp.max_iters) {  // spcs:1705-1746
                            st.v[FKS_STAT_UNSUCCESSFUL_RESOLVES]++;
                            if (has_self) st.v[FKS_STAT_UNSUCCESSFUL_SELF_COLLISION_RESOLVES]++;
                            else st.v[FKS_STAT_UNSUCCESSFUL_ENV_COLLISION_RESOLVES]++;
                            copy_state<KIND>(c, c.wl.pcfg, c.wl.Tprev, c.wl.cfg, c.wl.Tcur);  // return previous_configuration
                            step_collided = true;
                            step_failed = true;
                            break;
                        }
                        if ((resolver_iterations % sp.decay_iters) == 0u) {  // spcs:1747-1761
                            if (scaling >= 0.0) {
                                scaling = scaling * sp.decay_rate;
                                if (scaling < sp.min_scaling) scaling = -sp.min_scaling;
                            } else {
                                scaling = -sp.min_scaling;
                            }
                        }
                    }
                    if (step_failed) break;
                } else if (in_collision && !a.allow_contacts) {  // spcs:1769-1786
                    st.v[FKS_STAT_SUCCESSFUL_RESOLVES]++;
                    copy_state<KIND>(c, c.wl.pcfg, c.wl.Tprev, c.wl.cfg, c.wl.Tcur);
                    step_stopped = true;
                    break;
                }
            }
            if (!step_failed && !step_stopped) {  // spcs:1802-1814
                st.v[FKS_STAT_SUCCESSFUL_RESOLVES]++;
                if (step_collided) st.v[FKS_STAT_COLLISION_RESOLVES]++;
                else st.v[FKS_STAT_FREE_RESOLVES]++;
            }

            // ---- back in ForwardSimulateMutableRobot (spcs:873-909) ------------------------------------
            if (a.allow_contacts || !step_collided) {
                if (step_collided) collided = true;
                if (step_failed) {
                    flags |= FKS_FLAG_RESOLVE_FAILED;
                    if (sp.failed_ends_motion) {
                        flags |= FKS_FLAG_ENDED_BY_FAILURE;
                        break;
                    }
                    any_resolve_failed = true;
                } else if (any_resolve_failed) {
                    st.v[FKS_STAT_RECOVERED_UNSUCCESSFUL_RESOLVES]++;
                }
                if (sp.shortcut_distance > 0.0) {  // ComputeConfigurationDistanceTo (spcs:898); a distance is never < 0
                    double dist;
                    if (KIND == FKS_ROBOT_SE2) {
                        const double dx = fabs(target[0] - cfg[0]), dy = fabs(target[1] - cfg[1]);
                        const double dr = fabs(wrap_angle(target[2] - cfg[2]));
                        dist = (sqrt(dx * dx + dy * dy) * rb->pos_w) + (dr * rb->rot_w);
                    } else if (KIND == FKS_ROBOT_SE3) {
                        double cur[12], tg[12], ci[12], Dm[12];
#pragma unroll
                        for (int i = 0; i < 12; i++) {
                            cur[i] = cfg[i];
                            tg[i] = target[i];
                        }
                        const double dx = tg[3] - cur[3], dy = tg[7] - cur[7], dz = tg[11] - cur[11];
                        iso_inverse(cur, ci);
                        iso_mul(ci, tg, Dm);
                        const double cc = fmin(fmax(0.5 * (Dm[0] + Dm[5] + Dm[10] - 1.0), -1.0), 1.0);
                        dist = (sqrt(dx * dx + dy * dy + dz * dz) * rb->pos_w) + (acos(cc) * rb->rot_w);
                    } else {
                        double s = 0.0;
                        for (int j = 0; j < c.J; j++) {
                            const DevJoint& jd = rb->joints[j];
                            if (jd.active < 0) continue;
                            double dj = target[jd.active] - cfg[jd.active];
                            if (jd.type == FKS_JOINT_CONTINUOUS) dj = wrap_angle(dj);
                            const double wd = dj * jd.weight;
                            s += wd * wd;
                        }
                        dist = sqrt(s);
                    }
                    if (dist < sp.shortcut_distance) {
                        flags |= FKS_FLAG_ENDED_BY_SHORTCUT;
                        break;
                    }
                }
            } else {
                // robot->SetPosition(resolved_configuration) is skipped: the particle stays where the step began
                if (lane < c.stride) cfg[lane] = scfg[lane];
                __syncwarp();
                forward_kinematics<KIND>(c, cfg, Tcur);
                flags |= FKS_FLAG_ENDED_BY_NOCONTACT;
                break;
            }
        }
        if (collided) flags |= FKS_FLAG_DID_CONTACT;
        flags |= __reduce_or_sync(FKS_FULL, c.lflags);
        // ---- result record: cfg_stride doubles + fks_result_tail --------------------------------------
        char* rec = a.results + (size_t)pid * a.rec_stride;
        if (lane < c.stride) reinterpret_cast<double*>(rec)[lane] = cfg[lane];
        if (lane == 0) {
            unsigned* tail = reinterpret_cast<unsigned*>(rec + (size_t)c.stride * 8);  // fks_result_tail
            tail[0] = flags;
            tail[1] = n_micro_total;
            tail[2] = n_iter_total;
            tail[3] = n_steps;
        }
        st.v[FKS_STAT_TOTAL_MICROSTEPS] += n_micro_total;
        st.v[FKS_STAT_TOTAL_RESOLVER_ITERATIONS] += n_iter_total;
        __syncwarp();
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < FKS_NUM_STATS; i++)
            if (st.v[i]) atomicAdd(a.stats + i, st.v[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// roofline micro-benchmarks: FP64 FMA throughput and random 4-byte gather rate
// ------------------------------------------------------------------------------------------------
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

__global__ void gather_kernel(const float* __restrict__ data, unsigned long long n_mask, float* out, int iters) {
    unsigned long long x = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345ull;
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        // 4 independent gathers per iteration (xorshift addresses)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            acc += __ldg(data + (x & n_mask));
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
size_t simulate_dyn_smem(int kind, int L, int J, int D, int P, int stride) {
    const WarpLayout wl = make_warp_layout(kind, L, J, D, stride);
    size_t off = (sizeof(DevRobot) + 15) & ~(size_t)15;
    off += (size_t)P * 24 + (((size_t)P * 4 + 15) & ~(size_t)15);
    off += (size_t)kWarpsPerBlock * wl.total * 8;
    return off;
}

static const void* kernel_ptr(int kind) {
    switch (kind) {
        case FKS_ROBOT_SE2: return (const void*)simulate_kernel<FKS_ROBOT_SE2>;
        case FKS_ROBOT_SE3: return (const void*)simulate_kernel<FKS_ROBOT_SE3>;
        case FKS_ROBOT_LINKED: return (const void*)simulate_kernel<FKS_ROBOT_LINKED>;
    }
    return nullptr;
}

int simulate_kernel_info(int kind, size_t dyn_smem, KernelInfo* out) {
    const void* fn = kernel_ptr(kind);
    if (!fn) return (int)cudaErrorInvalidValue;
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    if (err != cudaSuccess) return (int)err;
    cudaFuncAttributes fa;
    err = cudaFuncGetAttributes(&fa, fn);
    if (err != cudaSuccess) return (int)err;
    int nb = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kThreadsPerBlock, dyn_smem);
    if (err != cudaSuccess) return (int)err;
    out->regs = fa.numRegs;
    out->static_smem = (int)fa.sharedSizeBytes;
    out->local_bytes = (int)fa.localSizeBytes;
    out->max_blocks_per_sm = nb;
    out->dyn_smem = dyn_smem;
    return 0;
}

int launch_simulate(int kind, const LaunchArgs& args, int grid, size_t dyn_smem, void* stream,
                    const void* l2_window_base, size_t l2_window_bytes) {
    const void* fn = kernel_ptr(kind);
    if (!fn) return (int)cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreadsPerBlock);
    cfg.dynamicSmemBytes = dyn_smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (l2_window_base && l2_window_bytes) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(l2_window_base);
        attr[0].val.accessPolicyWindow.num_bytes = l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    void* params[1] = {const_cast<LaunchArgs*>(&args)};
    return (int)cudaLaunchKernelExC(&cfg, fn, params);
}

int launch_fp64_peak(double* out, int grid, int iters, void* stream) {
    fp64_peak_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, iters);
    return (int)cudaGetLastError();
}
int launch_gather(const float* data, unsigned long long n_mask, float* out, int grid, int iters, void* stream) {
    gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(data, n_mask, out, iters);
    return (int)cudaGetLastError();
}

}  // namespace fksdev
