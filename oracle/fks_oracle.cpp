// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A scalar C++ restatement of the reference's batched particle forward-simulation path
// (calderpg/fast_kinematic_simulator).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product (libfksgpu.so) never does.
//
// PARITY UNPINNED (except the PID controller): the reference ships no tests, golden vectors or
// fixtures, and cannot be compiled here (Eigen, ROS, arc_utilities, sdf_tools and
// uncertainty_planning_core are absent and unpinned).  Control flow and constants
// follow the reference line by line (cited below, paths relative to /root/reference/include/
// fast_kinematic_simulator/: spcs = simple_particle_contact_simulator.hpp, tnuva =
// tnuva_robot_models.hpp, unc = simple_uncertainty_models.hpp, pid = simple_pid_controller.hpp).
// Arithmetic that lives in the un-vendored dependencies is restated from their published
// behaviour; each such choice is marked "RESTATEMENT" and listed in DESIGN.md.
// The PID restatement IS pinned: oracle/Makefile builds oracle/_ref/pid_ref from the reference's
// own simple_pid_controller.hpp and tests/test_oracle_pid.py compares against it.
//
// Besides results, the oracle records (a) the tape of truncated-normal draws so the GPU can replay
// the exact noise (injection mode), and (b) a per-particle SENSITIVITY mask: which discrete
// decisions came within a tolerance of flipping (cell boundaries, contact thresholds, rank cut,
// pivot ties ...).  Parity tests demand exact flags/counters for insensitive particles and
// enumerate the sensitive ones.

#include "fksgpu.h"
#include "fks_philox.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

// CPU model of the DEVICE solver (fks_qr_model.cpp), used only when a study asks for it (oracle_set_qr_model)
extern "C" void oracle_qr_device_model(const double* A_colmajor, const double* b, int rows, int cols, int opt_bits, double* x, double* out4);
extern "C" void oracle_qr_device_model_forced(const double* A_colmajor, const double* b, int rows, int cols, int opt_bits, uint64_t order,
                                              int rank, double* x);

namespace {

// ------------------------------------------------------------------------------------------------
// sensitivity bits (oracle only)
// ------------------------------------------------------------------------------------------------
enum {
    SENS_CELL_BOUNDARY = 1u << 0,  // a grid coordinate within tol of an integer
    SENS_EST_THRESHOLD = 1u << 1,  // EstimateDistance within tol of the collision threshold (spcs:968)
    SENS_EST_ZERO = 1u << 2,       // EstimateDistance within tol of 0 in the resolver (spcs:1874)
    SENS_RANK_CUT = 1u << 3,       // QR pivot within 1e4x of the rank threshold
    SENS_PIVOT_TIE = 1u << 4,      // two pivot candidates within rel tol, not identical
    SENS_NMICRO = 1u << 5,         // motion/target within tol of an integer (ceil, spcs:1562)
    SENS_NORMAL_TIE = 1u << 6,     // best two entry-direction dots within tol (spcs:100,123)
    SENS_ANGLE_WRAP = 1u << 7,     // an angle within tol of +-pi at a wrap
    SENS_SELF_COLLISION = 1u << 8, // a self-collision candidate cell was evaluated
    SENS_RAW_THRESHOLD = 1u << 9,  // (unused: raw float compares only change with the cell)
    SENS_STEP_FRACTION = 1u << 10, // m/res within tol of 1 (max() kink, spcs:1681)
    SENS_ILL_CONDITIONED = 1u << 11 // a stacked system with condition estimate > kIllConditioned was solved (and not replaced through
                                    // the decision tape): the step is reproducible to cond x epsilon only, and later solves amplify
                                    // the difference.  Measured (tests/test_oracle_decision_tape.py): with another solver arithmetic
                                    // every particle whose largest estimate stays below 1e3 reproduces to 1e-9, the others mostly do
};
constexpr double kIllConditioned = 1e3;

const double kPi = 3.14159265358979323846;

struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
// EigenHelpers::SafeNormal (RESTATEMENT): v/|v| if |v| > DBL_EPSILON else v
inline V3 safe_normal(V3 v) {
    const double n = norm(v);
    if (n > DBL_EPSILON) return {v.x / n, v.y / n, v.z / n};
    return v;
}

// rigid transform, row-major 3x4 [R|t]
struct Iso {
    double m[12];
};
inline Iso iso_identity() { return {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}}; }
inline V3 iso_apply(const Iso& T, V3 p) {
    return {T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z + T.m[3], T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z + T.m[7],
            T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z + T.m[11]};
}
inline V3 iso_rotate(const Iso& T, V3 p) {
    return {T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z, T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z,
            T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z};
}
inline V3 iso_translation(const Iso& T) { return {T.m[3], T.m[7], T.m[11]}; }
// Isometry3d product: (Ra,ta)(Rb,tb) = (Ra Rb, Ra tb + ta); sums in k order
inline Iso iso_mul(const Iso& A, const Iso& B) {
    Iso C;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
            C.m[4 * r + c] = A.m[4 * r + 0] * B.m[c] + A.m[4 * r + 1] * B.m[4 + c] + A.m[4 * r + 2] * B.m[8 + c];
        C.m[4 * r + 3] = A.m[4 * r + 0] * B.m[3] + A.m[4 * r + 1] * B.m[7] + A.m[4 * r + 2] * B.m[11] + A.m[4 * r + 3];
    }
    return C;
}
inline Iso iso_inverse(const Iso& A) {
    Iso C;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C.m[4 * r + c] = A.m[4 * c + r];
    for (int r = 0; r < 3; r++) C.m[4 * r + 3] = -(C.m[4 * r + 0] * A.m[3] + C.m[4 * r + 1] * A.m[7] + C.m[4 * r + 2] * A.m[11]);
    return C;
}

// Quaterniond(AngleAxisd(angle, axis)).toRotationMatrix() (RESTATEMENT of Eigen: w = cos(a/2),
// v = sin(a/2) axis; toRotationMatrix with the tx/ty/tz products)
inline Iso rotation_about_axis(double angle, V3 axis) {
    const double ha = 0.5 * angle;
    const double w = std::cos(ha), s = std::sin(ha);
    const double x = s * axis.x, y = s * axis.y, z = s * axis.z;
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Iso R;
    R.m[0] = 1.0 - (tyy + tzz);
    R.m[1] = txy - twz;
    R.m[2] = txz + twy;
    R.m[3] = 0.0;
    R.m[4] = txy + twz;
    R.m[5] = 1.0 - (txx + tzz);
    R.m[6] = tyz - twx;
    R.m[7] = 0.0;
    R.m[8] = txz - twy;
    R.m[9] = tyz + twx;
    R.m[10] = 1.0 - (txx + tyy);
    R.m[11] = 0.0;
    return R;
}

// EigenHelpers::EnforceContinuousRevoluteBounds (RESTATEMENT): wrap to (-pi, pi]
inline double wrap_angle(double value) {
    if ((value <= -kPi) || (value > kPi)) {
        const double remainder = std::fmod(value, 2.0 * kPi);
        if (remainder <= -kPi) return remainder + (2.0 * kPi);
        if (remainder > kPi) return remainder - (2.0 * kPi);
        return remainder;
    }
    return value;
}

// EigenHelpers::ExpTwist(twist, 1.0) (RESTATEMENT; call sites tnuva:360,378): twist = (v, w)
Iso exp_twist(const double* twist) {
    const V3 tv = {twist[0], twist[1], twist[2]};
    const V3 rv = {twist[3], twist[4], twist[5]};
    const double rn = norm(rv);
    Iso T = iso_identity();
    if (rn >= 1e-100) {
        const double theta = rn * 1.0;
        const V3 sv = {tv.x / rn, tv.y / rn, tv.z / rn};
        const V3 w = {rv.x / rn, rv.y / rn, rv.z / rn};
        // ExpMatrixExact: I + hat(w) sin(theta) + hat(w)^2 (1 - cos(theta))
        const double s = std::sin(theta), c1 = 1.0 - std::cos(theta);
        const double K[9] = {0.0, -w.z, w.y, w.z, 0.0, -w.x, -w.y, w.x, 0.0};
        double K2[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) K2[3 * r + c] = K[3 * r + 0] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
        double R[9];
        for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + (K[i] * s) + (K2[i] * c1);
        // (I - R) (w x v) + w (w . v) theta
        const V3 wxv = cross(w, sv);
        const double wv = dot(w, sv);
        double t[3];
        const double wxva[3] = {wxv.x, wxv.y, wxv.z};
        const double wa[3] = {w.x, w.y, w.z};
        for (int r = 0; r < 3; r++) {
            double acc = 0.0;
            for (int c = 0; c < 3; c++) acc += (((r == c) ? 1.0 : 0.0) - R[3 * r + c]) * wxva[c];
            t[r] = acc + (wa[r] * wv) * theta;
        }
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) T.m[4 * r + c] = R[3 * r + c];
            T.m[4 * r + 3] = t[r];
        }
    } else {
        T.m[3] = tv.x;
        T.m[7] = tv.y;
        T.m[11] = tv.z;
    }
    return T;
}

// EigenHelpers::TwistBetweenTransforms(a, b) = unhat(log(a^-1 b)) (RESTATEMENT; call site tnuva:389).
// Upstream evaluates Eigen's generic matrix logarithm; this is the closed-form SE(3) logarithm.
void twist_between(const Iso& a, const Iso& b, double* twist) {
    const Iso D = iso_mul(iso_inverse(a), b);
    const double tr = D.m[0] + D.m[5] + D.m[10];
    const V3 ax = {D.m[9] - D.m[6], D.m[2] - D.m[8], D.m[4] - D.m[1]};  // 2 sin(theta) * axis
    const double s2 = norm(ax);                                          // 2 sin(theta)
    const double c = 0.5 * (tr - 1.0);
    const double theta = std::atan2(0.5 * s2, c);
    const V3 t = iso_translation(D);
    V3 w;
    if (theta < 1e-9) {
        w = ax * 0.5;
        const V3 v = t - cross(w, t) * 0.5;
        twist[0] = v.x; twist[1] = v.y; twist[2] = v.z;
        twist[3] = w.x; twist[4] = w.y; twist[5] = w.z;
        return;
    }
    if (kPi - theta < 1e-6) {
        // near pi: axis from the diagonal of (R + I)/2 = axis axis^T (to first order)
        double xx = std::sqrt(std::max(0.0, 0.5 * (D.m[0] + 1.0)));
        double yy = std::sqrt(std::max(0.0, 0.5 * (D.m[5] + 1.0)));
        double zz = std::sqrt(std::max(0.0, 0.5 * (D.m[10] + 1.0)));
        // fix signs from the largest component
        if (xx >= yy && xx >= zz) {
            if (D.m[1] + D.m[4] < 0.0) yy = -yy;
            if (D.m[2] + D.m[8] < 0.0) zz = -zz;
            if (ax.x < 0.0) { xx = -xx; yy = -yy; zz = -zz; }
        } else if (yy >= zz) {
            if (D.m[1] + D.m[4] < 0.0) xx = -xx;
            if (D.m[6] + D.m[9] < 0.0) zz = -zz;
            if (ax.y < 0.0) { xx = -xx; yy = -yy; zz = -zz; }
        } else {
            if (D.m[2] + D.m[8] < 0.0) xx = -xx;
            if (D.m[6] + D.m[9] < 0.0) yy = -yy;
            if (ax.z < 0.0) { xx = -xx; yy = -yy; zz = -zz; }
        }
        const V3 axis = safe_normal({xx, yy, zz});
        w = axis * theta;
    } else {
        w = ax * (theta / s2);
    }
    // V^-1 = I - 1/2 hat(w) + k hat(w)^2,  k = (1 - theta sin(theta) / (2 (1 - cos(theta)))) / theta^2
    const double st = std::sin(theta), ct = std::cos(theta);
    const double k = (1.0 - (theta * st) / (2.0 * (1.0 - ct))) / (theta * theta);
    const V3 wxt = cross(w, t);
    const V3 wxwxt = cross(w, wxt);
    const V3 v = t - wxt * 0.5 + wxwxt * k;
    twist[0] = v.x; twist[1] = v.y; twist[2] = v.z;
    twist[3] = w.x; twist[4] = w.y; twist[5] = w.z;
}

inline double clamp_value(double v, double lo, double hi) { return std::min(std::max(v, lo), hi); }  // arc_helpers::ClampValue

// ------------------------------------------------------------------------------------------------
// PID (pid:98-135) -- literal
// ------------------------------------------------------------------------------------------------
struct Pid {
    double kp, ki, kd, iclamp, integral, last_error;
    void init(double p, double i, double d, double c) {  // pid:104-113
        kp = std::abs(p);
        ki = std::abs(i);
        kd = std::abs(d);
        iclamp = std::abs(c);
        integral = 0.0;
        last_error = 0.0;
    }
    void zero() { last_error = 0.0; integral = 0.0; }  // pid:98-102
    double feedback(double current_error, double timestep) {  // pid:122-135
        const double timestep_error_integral = ((current_error * 0.5) + (last_error * 0.5)) * timestep;
        const double new_error_integral = integral + timestep_error_integral;
        integral = std::max(-iclamp, std::min(iclamp, new_error_integral));
        const double error_derivative = (current_error - last_error) / timestep;
        last_error = current_error;
        return (current_error * kp) + (integral * ki) + (error_derivative * kd);
    }
};

// ------------------------------------------------------------------------------------------------
// Environment: SDF + normals (spcs:44-256; sdf_tools restated)
// ------------------------------------------------------------------------------------------------
struct Env {
    fks_env_desc d;
    Iso origin, inv_origin;
    std::unordered_map<int64_t, std::pair<uint32_t, uint32_t>> normal_cells;
    double sens_tol;

    void init(const fks_env_desc* desc) {
        d = *desc;
        std::memcpy(origin.m, desc->origin, sizeof(origin.m));
        std::memcpy(inv_origin.m, desc->inverse_origin, sizeof(inv_origin.m));
        normal_cells.reserve((size_t)desc->n_normal_cells * 2 + 16);
        for (int64_t i = 0; i < desc->n_normal_cells; i++)
            normal_cells[desc->normal_cell_index[i]] = {desc->normal_cell_start[i], desc->normal_cell_start[i + 1]};
        sens_tol = 1e-9;
    }
    // VoxelGrid::LocationToGridIndex4d (RESTATEMENT): grid-frame point * (1/cell), C-cast
    inline bool index_of(V3 p, double res, int64_t* ix, int64_t* iy, int64_t* iz, uint32_t* sens) const {
        const V3 g = iso_apply(inv_origin, p);
        const double inv = 1.0 / res;
        const double gx = g.x * inv, gy = g.y * inv, gz = g.z * inv;
        *ix = (int64_t)gx;
        *iy = (int64_t)gy;
        *iz = (int64_t)gz;
        if (sens) {
            if (std::abs(gx - std::nearbyint(gx)) < sens_tol || std::abs(gy - std::nearbyint(gy)) < sens_tol ||
                std::abs(gz - std::nearbyint(gz)) < sens_tol)
                *sens |= SENS_CELL_BOUNDARY;
        }
        return *ix >= 0 && *iy >= 0 && *iz >= 0 && *ix < d.nx && *iy < d.ny && *iz < d.nz;
    }
    inline float cell(int64_t x, int64_t y, int64_t z) const { return d.sdf[(size_t)((x * d.ny + y) * d.nz + z)]; }
    // SignedDistanceField::GetImmutable4d (spcs:941)
    inline std::pair<float, bool> get4d(V3 p, uint32_t* sens) const {
        int64_t x, y, z;
        if (index_of(p, d.sdf_resolution, &x, &y, &z, sens)) return {cell(x, y, z), true};
        return {d.oob_value, false};
    }
    // SignedDistanceField::GetGradient(idx, true) (RESTATEMENT)
    inline void gradient(int64_t x, int64_t y, int64_t z, double* g) const {
        const double res = d.sdf_resolution;
        if (x > 0 && y > 0 && z > 0 && x < d.nx - 1 && y < d.ny - 1 && z < d.nz - 1) {
            const double inv_twice_res = 1.0 / (2.0 * res);
            g[0] = (double)(cell(x + 1, y, z) - cell(x - 1, y, z)) * inv_twice_res;
            g[1] = (double)(cell(x, y + 1, z) - cell(x, y - 1, z)) * inv_twice_res;
            g[2] = (double)(cell(x, y, z + 1) - cell(x, y, z - 1)) * inv_twice_res;
            return;
        }
        const int64_t lx = std::max<int64_t>(0, x - 1), hx = std::min<int64_t>(d.nx - 1, x + 1);
        const int64_t ly = std::max<int64_t>(0, y - 1), hy = std::min<int64_t>(d.ny - 1, y + 1);
        const int64_t lz = std::max<int64_t>(0, z - 1), hz = std::min<int64_t>(d.nz - 1, z + 1);
        const double ix = (double)(hx - lx) * res, iy = (double)(hy - ly) * res, iz = (double)(hz - lz) * res;
        g[0] = g[1] = g[2] = 0.0;
        if (ix > 0.0) g[0] = ((double)cell(hx, y, z) - (double)cell(lx, y, z)) * (1.0 / ix);
        if (iy > 0.0) g[1] = ((double)cell(x, hy, z) - (double)cell(x, ly, z)) * (1.0 / iy);
        if (iz > 0.0) g[2] = ((double)cell(x, y, hz) - (double)cell(x, y, lz)) * (1.0 / iz);
    }
    // SignedDistanceField::EstimateDistance4d (RESTATEMENT): centre distance shrunk by
    // half a cell, plus the signed length of (p - cell centre) projected on the gradient.
    inline std::pair<double, bool> estimate_distance(V3 p, uint32_t* sens) const {
        int64_t x, y, z;
        if (!index_of(p, d.sdf_resolution, &x, &y, &z, sens)) return {(double)d.oob_value, false};
        const double res = d.sdf_resolution;
        const double d0 = (double)cell(x, y, z);
        const double dc = (d0 >= 0.0) ? d0 - (res * 0.5) : d0 + (res * 0.5);
        double g[3];
        gradient(x, y, z, g);
        const V3 centre_grid = {res * ((double)x + 0.5), res * ((double)y + 0.5), res * ((double)z + 0.5)};
        const V3 centre = iso_apply(origin, centre_grid);
        const V3 v = p - centre;
        const double gg = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        double adj = 0.0;
        if (gg > 0.0) adj = (v.x * g[0] + v.y * g[1] + v.z * g[2]) / std::sqrt(gg);
        return {dc + adj, true};
    }
    // SurfaceNormalGrid::LookupSurfaceNormal(Vector4d location, Vector4d direction)
    // (spcs:186-198,206-210,235-256) + GetBestSurfaceNormal (spcs:111-132).
    // Returns found flag (in bounds); would_assert set when the reference's asserts (:113-115) fire.
    inline bool lookup_normal(V3 p, V3 dir, V3* out, bool* would_assert, uint32_t* sens) const {
        int64_t x, y, z;
        *out = {0.0, 0.0, 0.0};
        if (!index_of(p, d.sdf_resolution, &x, &y, &z, nullptr)) return false;
        const int64_t li = (x * d.ny + y) * d.nz + z;
        auto it = normal_cells.find(li);
        if (it == normal_cells.end() || it->second.first == it->second.second) return true;  // empty -> zero normal
        const double dn = norm(dir);
        if (!(dn > 0.0)) {  // assert(direction_norm > 0.0) spcs:115
            *would_assert = true;
            // documented device behaviour: treat the unit direction as zero -> first entry wins
        }
        const V3 u = (dn > 0.0) ? V3{dir.x / dn, dir.y / dn, dir.z / dn} : V3{0.0, 0.0, 0.0};
        int best = -1;
        double best_dot = -std::numeric_limits<double>::infinity(), second = -std::numeric_limits<double>::infinity();
        for (uint32_t e = it->second.first; e < it->second.second; e++) {
            const double* en = d.normal_entries + 7 * (size_t)e;
            const double dp = en[0] * u.x + en[1] * u.y + en[2] * u.z + en[3] * 0.0;
            if (dp > best_dot) {
                second = best_dot;
                best_dot = dp;
                best = (int)e;
            } else if (dp > second) {
                second = dp;
            }
        }
        if (sens && (best_dot - second) < sens_tol && it->second.second - it->second.first > 1) {
            // identical normals in tied entries would be harmless, but flag anyway
            *sens |= SENS_NORMAL_TIE;
        }
        const double* en = d.normal_entries + 7 * (size_t)best;
        *out = {en[4], en[5], en[6]};
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// Robots (tnuva.hpp + arc_utilities PointSphereBasic{SE2,SE3,Linked}Robot restated)
// ------------------------------------------------------------------------------------------------
const int kMaxDof = 16;
const int kMaxLinks = 16;
const int kMaxJoints = 16;

// Immutable description shared by every clone (the reference shares geometry through shared_ptr).
struct Model {
    int kind, L, J, D;
    std::vector<V3> points;
    std::vector<int> point_link;
    std::vector<int> link_begin;  // L+1
    std::vector<fks_axis_params> axes;
    Iso base;
    std::vector<fks_joint_desc> joints;
    std::vector<int> joint_active_index;   // joint -> active index or -1
    std::vector<int> link_parent_joint;    // link -> joint whose child it is, or -1
    std::vector<uint8_t> allowed;
    double pos_w, rot_w;

    bool init(const fks_robot_desc* r) {
        kind = r->kind;
        L = r->n_links;
        J = r->n_joints;
        D = r->n_dof;
        if (D > kMaxDof || L < 1 || L > kMaxLinks || J > kMaxJoints) return false;
        points.resize((size_t)r->n_points);
        point_link.resize((size_t)r->n_points);
        link_begin.assign((size_t)L + 1, 0);
        for (int64_t i = 0; i < r->n_points; i++) {
            points[(size_t)i] = {r->points_xyz[3 * i], r->points_xyz[3 * i + 1], r->points_xyz[3 * i + 2]};
            point_link[(size_t)i] = r->point_link[i];
            if (i > 0 && r->point_link[i] < r->point_link[i - 1]) return false;
            if (r->point_link[i] < 0 || r->point_link[i] >= L) return false;
            link_begin[(size_t)r->point_link[i] + 1]++;
        }
        for (int l = 0; l < L; l++) link_begin[(size_t)l + 1] += link_begin[(size_t)l];
        axes.assign(r->axes, r->axes + D);
        std::memcpy(base.m, r->base_transform, sizeof(base.m));
        joints.clear();
        joint_active_index.clear();
        link_parent_joint.assign((size_t)L, -1);
        int active = 0;
        for (int j = 0; j < J; j++) {
            joints.push_back(r->joints[j]);
            joint_active_index.push_back(r->joints[j].type == FKS_JOINT_FIXED ? -1 : active++);
            link_parent_joint[(size_t)r->joints[j].child_link] = j;
        }
        if (kind == FKS_ROBOT_LINKED && active != D) return false;  // tnuva:503-516
        allowed.assign((size_t)L * L, 1);
        if (r->allowed_self_collision) allowed.assign(r->allowed_self_collision, r->allowed_self_collision + (size_t)L * L);
        pos_w = r->position_distance_weight;
        rot_w = r->rotation_distance_weight;
        return true;
    }
};

struct Robot {
    const Model* mdl;
    int kind, L, J, D;
    // state
    double cfg[kMaxDof];  // SE2: x,y,theta; SE3: 12; linked: D values
    Iso link_T[kMaxLinks];
    double joint_values[kMaxJoints];  // all joints (fixed ones hold their clamped value)

    int cfg_stride() const { return kind == FKS_ROBOT_SE2 ? 3 : (kind == FKS_ROBOT_SE3 ? 12 : D); }

    void bind(const Model* m) {
        mdl = m;
        kind = m->kind;
        L = m->L;
        J = m->J;
        D = m->D;
        for (int l = 0; l < kMaxLinks; l++) link_T[l] = iso_identity();
        std::memset(cfg, 0, sizeof(cfg));
        std::memset(joint_values, 0, sizeof(joint_values));
    }

    // SetPosition (arc_utilities; call sites spcs:875,1423-1424,1601): store config (wrap / limit) + FK
    void set_position(const double* c, uint32_t* sens) {
        if (kind == FKS_ROBOT_SE2) {
            cfg[0] = c[0];
            cfg[1] = c[1];
            if (sens && std::abs(std::abs(c[2]) - kPi) < 1e-9) *sens |= SENS_ANGLE_WRAP;
            cfg[2] = wrap_angle(c[2]);
            // Translation3d(x, y, 0) * Quaterniond(AngleAxisd(theta, UnitZ))
            Iso T = rotation_about_axis(cfg[2], {0.0, 0.0, 1.0});
            T.m[3] = cfg[0];
            T.m[7] = cfg[1];
            T.m[11] = 0.0;
            link_T[0] = T;
        } else if (kind == FKS_ROBOT_SE3) {
            std::memcpy(cfg, c, 12 * sizeof(double));
            std::memcpy(link_T[0].m, c, 12 * sizeof(double));
        } else {
            // SimpleJointModel::SetValue: continuous -> wrap, others -> clamp to limits
            for (int j = 0; j < J; j++) {
                const int a = mdl->joint_active_index[(size_t)j];
                const fks_joint_desc& jd = mdl->joints[(size_t)j];
                if (a < 0) {
                    joint_values[(size_t)j] = clamp_value(0.0, jd.lower_limit, jd.upper_limit);
                    continue;
                }
                double v = c[a];
                if (jd.type == FKS_JOINT_CONTINUOUS) {
                    if (sens && std::abs(std::abs(v) - kPi) < 1e-9) *sens |= SENS_ANGLE_WRAP;
                    v = wrap_angle(v);
                } else {
                    if (v > jd.upper_limit) v = jd.upper_limit;
                    else if (v < jd.lower_limit) v = jd.lower_limit;
                }
                cfg[a] = v;
                joint_values[(size_t)j] = v;
            }
            // UpdateTransforms (RESTATEMENT): child = (parent * joint_transform) * motion(value)
            link_T[0] = mdl->base;
            for (int j = 0; j < J; j++) {
                const fks_joint_desc& jd = mdl->joints[(size_t)j];
                Iso jt;
                std::memcpy(jt.m, jd.transform, sizeof(jt.m));
                const Iso complete = iso_mul(link_T[(size_t)jd.parent_link], jt);
                const V3 axis = {jd.axis[0], jd.axis[1], jd.axis[2]};
                if (jd.type == FKS_JOINT_REVOLUTE || jd.type == FKS_JOINT_CONTINUOUS) {
                    link_T[(size_t)jd.child_link] = iso_mul(complete, rotation_about_axis(joint_values[(size_t)j], axis));
                } else if (jd.type == FKS_JOINT_PRISMATIC) {
                    Iso tr = iso_identity();
                    const V3 t = axis * joint_values[(size_t)j];
                    tr.m[3] = t.x;
                    tr.m[7] = t.y;
                    tr.m[11] = t.z;
                    link_T[(size_t)jd.child_link] = iso_mul(complete, tr);
                } else {
                    link_T[(size_t)jd.child_link] = complete;
                }
            }
        }
    }

    // TruncatedNormalUncertainVelocityActuator::GetControlValue (unc:70-75 / 77-90).
    // `tn` = output of noise_distribution_(rng); nullptr = noiseless overload.
    static inline double actuate_axis(const fks_axis_params& a, double u, const double* tn, bool* nan_seen) {
        if (std::isnan(u) || std::isinf(u)) *nan_seen = true;  // assert unc:72-73
        const double vl = std::abs(a.velocity_limit);
        const double real_u = clamp_value(u, -vl, vl);
        if (!tn) return real_u;
        const double pb = std::abs(a.proportional_noise) * std::abs(real_u);
        const double mb = std::abs(a.minimum_noise) * vl;
        const double bound = std::max(pb, mb);
        const double real_noise = (*tn) * bound;
        return real_u + real_noise;
    }
    inline double actuate(int axis, double u, const double* tn, bool* nan_seen) const {
        return actuate_axis(mdl->axes[(size_t)axis], u, tn, nan_seen);
    }

    // ApplyControlInput(input[, rng]) (tnuva:152-177 SE2, :348-382 SE3, :538-596 linked)
    void apply_control(const double* u, const double* tn, bool* nan_seen, uint32_t* sens) {
        double r[kMaxDof];
        for (int i = 0; i < D; i++) r[i] = actuate(i, u[i], tn ? tn + i : nullptr, nan_seen);
        if (kind == FKS_ROBOT_SE2) {
            const double nc[3] = {cfg[0] + r[0], cfg[1] + r[1], cfg[2] + r[2]};
            set_position(nc, sens);
        } else if (kind == FKS_ROBOT_SE3) {
            Iso cur;
            std::memcpy(cur.m, cfg, sizeof(cur.m));
            const Iso nc = iso_mul(cur, exp_twist(r));
            set_position(nc.m, sens);
        } else {
            double nc[kMaxDof];
            for (int i = 0; i < D; i++) nc[i] = cfg[i] + r[i];
            set_position(nc, sens);
        }
    }

    // GenerateControlAction(target, dt) (tnuva:179-198, :384-412, :598-614)
    void control_action(const double* target, double dt, Pid* pids, double* out, bool* nan_seen) const {
        double err[kMaxDof];
        if (kind == FKS_ROBOT_SE2) {
            // ComputePerDimensionConfigurationSignedDistance (RESTATEMENT): (dx, dy, shortest angle)
            err[0] = target[0] - cfg[0];
            err[1] = target[1] - cfg[1];
            err[2] = wrap_angle(target[2] - cfg[2]);
        } else if (kind == FKS_ROBOT_SE3) {
            Iso cur, tgt;
            std::memcpy(cur.m, cfg, sizeof(cur.m));
            std::memcpy(tgt.m, target, sizeof(tgt.m));
            twist_between(cur, tgt, err);
        } else {
            // ComputeUnweightedPerDimensionConfigurationRawDistance (RESTATEMENT)
            for (int j = 0; j < J; j++) {
                const int a = mdl->joint_active_index[(size_t)j];
                if (a < 0) continue;
                if (mdl->joints[(size_t)j].type == FKS_JOINT_CONTINUOUS) err[a] = wrap_angle(target[a] - cfg[a]);
                else err[a] = target[a] - cfg[a];
            }
        }
        for (int i = 0; i < D; i++) {
            const double term = pids[i].feedback(err[i], dt);
            out[i] = actuate(i, term, nullptr, nan_seen);
        }
    }

    // ComputeConfigurationDistanceTo (RESTATEMENT; only matters when simulation_shortcut_distance > 0)
    double distance_to(const double* target) const {
        if (kind == FKS_ROBOT_SE2) {
            const double dx = std::abs(target[0] - cfg[0]), dy = std::abs(target[1] - cfg[1]);
            const double dr = std::abs(wrap_angle(target[2] - cfg[2]));
            return (std::sqrt(dx * dx + dy * dy) * mdl->pos_w) + (dr * mdl->rot_w);
        } else if (kind == FKS_ROBOT_SE3) {
            Iso cur, tgt;
            std::memcpy(cur.m, cfg, sizeof(cur.m));
            std::memcpy(tgt.m, target, sizeof(tgt.m));
            const V3 dt3 = iso_translation(tgt) - iso_translation(cur);
            const Iso Dm = iso_mul(iso_inverse(cur), tgt);
            const double c = clamp_value(0.5 * (Dm.m[0] + Dm.m[5] + Dm.m[10] - 1.0), -1.0, 1.0);
            return (norm(dt3) * mdl->pos_w) + (std::acos(c) * mdl->rot_w);
        }
        double s = 0.0;
        for (int j = 0; j < J; j++) {
            const int a = mdl->joint_active_index[(size_t)j];
            if (a < 0) continue;
            double dj = target[a] - cfg[a];
            if (mdl->joints[(size_t)j].type == FKS_JOINT_CONTINUOUS) dj = wrap_angle(dj);
            const double wd = dj * mdl->joints[(size_t)j].distance_weight;
            s += wd * wd;
        }
        return std::sqrt(s);
    }

    // ComputeLinkPointTranslationJacobian (RESTATEMENT); J is 3 x D row-major
    void point_jacobian(int link, V3 pl, double* Jm) const {
        for (int i = 0; i < 3 * D; i++) Jm[i] = 0.0;
        if (kind == FKS_ROBOT_SE2) {
            const V3 pw = iso_apply(link_T[0], pl);
            const V3 cur = {link_T[0].m[3], link_T[0].m[7], 0.0};
            const V3 c2 = cross({0.0, 0.0, 1.0}, pw - cur);
            Jm[0 * D + 0] = 1.0;
            Jm[1 * D + 1] = 1.0;
            Jm[0 * D + 2] = c2.x;
            Jm[1 * D + 2] = c2.y;
            Jm[2 * D + 2] = c2.z;
        } else if (kind == FKS_ROBOT_SE3) {
            // R * [I, -skew(p)]
            const Iso& T = link_T[0];
            const double S[9] = {0.0, pl.z, -pl.y, -pl.z, 0.0, pl.x, pl.y, -pl.x, 0.0};  // -skew(p)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    Jm[r * D + c] = T.m[4 * r + c];
                    Jm[r * D + 3 + c] = T.m[4 * r + 0] * S[c] + T.m[4 * r + 1] * S[3 + c] + T.m[4 * r + 2] * S[6 + c];
                }
        } else {
            const V3 pw = iso_apply(link_T[(size_t)link], pl);
            int j = mdl->link_parent_joint[(size_t)link];
            while (j >= 0) {
                const fks_joint_desc& jd = mdl->joints[(size_t)j];
                const Iso& Tj = link_T[(size_t)jd.child_link];
                const int a = mdl->joint_active_index[(size_t)j];
                if (a >= 0) {
                    const V3 axis = {jd.axis[0], jd.axis[1], jd.axis[2]};
                    if (jd.type == FKS_JOINT_PRISMATIC) {
                        const V3 aw = iso_rotate(Tj, axis);
                        Jm[0 * D + a] += aw.x;
                        Jm[1 * D + a] += aw.y;
                        Jm[2 * D + a] += aw.z;
                    } else {
                        const V3 aw = iso_rotate(Tj, axis);
                        const V3 col = cross(aw, pw - iso_translation(Tj));
                        Jm[0 * D + a] += col.x;
                        Jm[1 * D + a] += col.y;
                        Jm[2 * D + a] += col.z;
                    }
                }
                j = mdl->link_parent_joint[(size_t)jd.parent_link];
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Column-pivoted Householder QR solve with Eigen 3.3 semantics (RESTATEMENT)
// A is rows x cols COLUMN-major (destroyed); b has rows entries (destroyed); x has cols entries.
// ------------------------------------------------------------------------------------------------
// debug histogram of log10(pivot_norm^2 / rank_cut) for pivots near the rank decision (oracle_debug_rank_hist)
long long g_rank_hist[42];
long long g_rows_hist[16];  // debug: histogram of stacked-Jacobian row counts, bucket = min(15, rows / 24)

// The discrete decisions of one solve (recorded on the DECISION TAPE next to the noise tape): Eigen's nonzero_pivots and
// the pivot order.  `roundoff_seen`: some pivot was at round-off level (0 < |pivot|^2 < 1e8 x Eigen's rank threshold; the
// histogram of |pivot|^2 / threshold is bimodal -- <= 10 or >= 1e17 on every workload here -- so 1e8 separates the modes);
// `roundoff_kept`: such a pivot was NOT cut, i.e. the solution divides by round-off and no other arithmetic (x86 with a
// different summation order, a GPU) reproduces it.
struct QrInfo {
    int rank, size;
    uint64_t order;  // nibble k: ORIGINAL index of the column picked at step k
    bool roundoff_seen, roundoff_kept, tie;
    double cond_est;  // |R(0,0)| / min |R(k,k)| over the kept pivots: the usual estimate of the pivoted triangle's condition
};
constexpr double kRoundoffPivotBand = 1e8;

// Every reduction over rows (squaredNorm, the reflector's dot products) goes through sum4.  RESTATEMENT: Eigen reduces
// with SSE2 packets -- two packets of two doubles, i.e. FOUR interleaved partial sums that are added pairwise at the end
// (redux_impl<..., LinearVectorizedTraversal>; the reference is built without -march, CMakeLists.txt:66) -- and where its
// packets start depends on the alignment of the column in memory, which no second implementation can reproduce.  Fixed
// here: partial sum i takes the entries whose ROW index is congruent to i modulo 4, in ascending row order, and the result
// is (p0 + p2) + (p1 + p3) (packet_res0 + packet_res1, then predux).  The device solver uses the same order (one lane per
// partial sum), which is what makes it return these bits (tests/test_gpu_qr_solver.py).
template <class Term>
inline double sum4(int lo, int hi, Term term) {
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    int r = lo;
    for (; r < hi && (r & 3) != 0; r++) {  // rows up to the next multiple of 4, each into the partial sum of its residue
        const double t = term(r);
        if ((r & 3) == 1) p1 += t;
        else if ((r & 3) == 2) p2 += t;
        else p3 += t;
    }
    for (; r + 3 < hi; r += 4) {
        p0 += term(r);
        p1 += term(r + 1);
        p2 += term(r + 2);
        p3 += term(r + 3);
    }
    if (r < hi) p0 += term(r);
    if (r + 1 < hi) p1 += term(r + 1);
    if (r + 2 < hi) p2 += term(r + 2);
    return (p0 + p2) + (p1 + p3);
}

void colpiv_qr_solve(double* A, double* b, int rows, int cols, double* x, uint32_t* sens, QrInfo* info = nullptr) {
    const int size = std::min(rows, cols);
    int colidx[kMaxDof];
    for (int j = 0; j < cols; j++) colidx[j] = j;
    QrInfo qi = {size, size, 0ull, false, false, false, 1.0};
    double hcoeff[kMaxDof];
    int transp[kMaxDof];
    double norms_updated[kMaxDof], norms_direct[kMaxDof];
    auto col = [&](int j) { return A + (size_t)j * rows; };
    double max_norm = 0.0;
    for (int k = 0; k < cols; k++) {
        const double* ck = col(k);
        norms_direct[k] = std::sqrt(sum4(0, rows, [&](int r) { return ck[r] * ck[r]; }));
        norms_updated[k] = norms_direct[k];
        max_norm = std::max(max_norm, norms_updated[k]);
    }
    const double eps = DBL_EPSILON;
    const double threshold_helper = ((max_norm * eps) * (max_norm * eps)) / (double)rows;
    const double norm_downdate_threshold = std::sqrt(eps);
    int nonzero_pivots = size;
    for (int k = 0; k < size; k++) {
        int biggest = k;
        double big = norms_updated[k];
        for (int j = k + 1; j < cols; j++)
            if (norms_updated[j] > big) {
                big = norms_updated[j];
                biggest = j;
            }
        for (int j = k; j < cols; j++)
            if (j != biggest && norms_updated[j] != big && std::abs(norms_updated[j] - big) <= 1e-9 * big) {
                if (sens) *sens |= SENS_PIVOT_TIE;
                qi.tie = true;
            }
        const double big_sq = big * big;
        const double cut = threshold_helper * (double)(rows - k);
        if (nonzero_pivots == size && big_sq < cut) nonzero_pivots = k;
        if (max_norm > 0.0 && big_sq < cut * kRoundoffPivotBand && big_sq > 0.0) {
            if (sens) *sens |= SENS_RANK_CUT;
            qi.roundoff_seen = true;
            if (nonzero_pivots == size) qi.roundoff_kept = true;
        }
        qi.order |= (uint64_t)colidx[biggest] << (4 * k);
        if (sens && max_norm > 0.0) {
            int bucket = 41;  // exact zero
            if (big_sq > 0.0) {
                const double lg = std::log10(big_sq / cut);
                bucket = (int)std::floor(std::min(std::max(lg, -20.0), 19.0)) + 20;
            }
#pragma omp atomic
            g_rank_hist[bucket]++;
        }
        transp[k] = biggest;
        std::swap(colidx[k], colidx[biggest]);
        if (k != biggest) {
            for (int r = 0; r < rows; r++) std::swap(col(k)[r], col(biggest)[r]);
            std::swap(norms_updated[k], norms_updated[biggest]);
            std::swap(norms_direct[k], norms_direct[biggest]);
        }
        // makeHouseholderInPlace on col(k).tail(rows-k)
        double* ck = col(k);
        const double tail_sq = sum4(k + 1, rows, [&](int r) { return ck[r] * ck[r]; });
        const double c0 = ck[k];
        double tau, beta;
        if (tail_sq <= DBL_MIN) {
            tau = 0.0;
            beta = c0;
            for (int r = k + 1; r < rows; r++) ck[r] = 0.0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0) beta = -beta;
            const double denom = c0 - beta;
            for (int r = k + 1; r < rows; r++) ck[r] = ck[r] / denom;
            tau = (beta - c0) / beta;
        }
        hcoeff[k] = tau;
        ck[k] = beta;
        // applyHouseholderOnTheLeft to bottomRightCorner(rows-k, cols-k-1)
        if (rows - k == 1) {
            for (int j = k + 1; j < cols; j++) col(j)[k] *= (1.0 - tau);
        } else if (tau != 0.0) {
            for (int j = k + 1; j < cols; j++) {
                double* cj = col(j);
                double tmp = sum4(k + 1, rows, [&](int r) { return ck[r] * cj[r]; });
                tmp += cj[k];
                cj[k] -= tau * tmp;
                for (int r = k + 1; r < rows; r++) cj[r] -= (tau * ck[r]) * tmp;
            }
        }
        // LAPACK-style norm downdate
        for (int j = k + 1; j < cols; j++) {
            if (norms_updated[j] != 0.0) {
                double temp = std::abs(col(j)[k]) / norms_updated[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double ratio = norms_updated[j] / norms_direct[j];
                const double temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    const double* cj = col(j);
                    norms_direct[j] = std::sqrt(sum4(k + 1, rows, [&](int r) { return cj[r] * cj[r]; }));
                    norms_updated[j] = norms_direct[j];
                } else {
                    norms_updated[j] *= std::sqrt(temp);
                }
            }
        }
    }
    int perm[kMaxDof];
    for (int j = 0; j < cols; j++) perm[j] = j;
    for (int k = 0; k < size; k++) std::swap(perm[k], perm[transp[k]]);
    for (int j = 0; j < cols; j++) x[j] = 0.0;
    qi.rank = nonzero_pivots;
    for (int k = 1; k < nonzero_pivots; k++) {
        const double dk = std::abs(col(k)[k]);
        qi.cond_est = std::max(qi.cond_est, dk > 0.0 ? std::abs(col(0)[0]) / dk : INFINITY);
    }
    if (info) *info = qi;
    if (nonzero_pivots == 0) return;
    // c = H_{nz-1} ... H_0 b
    for (int k = 0; k < nonzero_pivots; k++) {
        const double tau = hcoeff[k];
        const double* ck = col(k);
        if (rows - k == 1) {
            b[k] *= (1.0 - tau);
        } else if (tau != 0.0) {
            double tmp = sum4(k + 1, rows, [&](int r) { return ck[r] * b[r]; });
            tmp += b[k];
            b[k] -= tau * tmp;
            for (int r = k + 1; r < rows; r++) b[r] -= (tau * ck[r]) * tmp;
        }
    }
    // back substitution on the leading nz x nz upper triangle
    for (int i = nonzero_pivots - 1; i >= 0; i--) {
        double s = b[i];
        for (int j = i + 1; j < nonzero_pivots; j++) s -= col(j)[i] * b[j];
        b[i] = s / col(i)[i];
    }
    for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = b[i];
}

// small dense inverse by partial-pivot LU (Eigen dynamic MatrixXd::inverse, RESTATEMENT); n <= 16
bool lu_inverse(const double* A, int n, double* inv) {
    double M[16 * 32];
    for (int r = 0; r < n; r++) {
        for (int c = 0; c < n; c++) M[r * 2 * n + c] = A[r * n + c];
        for (int c = 0; c < n; c++) M[r * 2 * n + n + c] = (r == c) ? 1.0 : 0.0;
    }
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::abs(M[k * 2 * n + k]);
        for (int r = k + 1; r < n; r++)
            if (std::abs(M[r * 2 * n + k]) > best) {
                best = std::abs(M[r * 2 * n + k]);
                piv = r;
            }
        if (piv != k)
            for (int c = 0; c < 2 * n; c++) std::swap(M[k * 2 * n + c], M[piv * 2 * n + c]);
        const double p = M[k * 2 * n + k];
        for (int r = k + 1; r < n; r++) {
            const double f = M[r * 2 * n + k] / p;
            for (int c = k; c < 2 * n; c++) M[r * 2 * n + c] -= f * M[k * 2 * n + c];
        }
    }
    for (int c = 0; c < n; c++) {
        for (int r = n - 1; r >= 0; r--) {
            double s = M[r * 2 * n + n + c];
            for (int j = r + 1; j < n; j++) s -= M[r * 2 * n + j] * inv[j * n + c];
            inv[r * n + c] = s / M[r * 2 * n + r];
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// noise sources
// ------------------------------------------------------------------------------------------------
// arc_helpers::TruncatedNormalDistribution(0, sigma, -1, 1) (RESTATEMENT):
// sigma == 0 -> mean; standardised bounds straddle 0 -> naive accept/reject on
// std::normal_distribution<double>, which is libstdc++'s own (the cached second variate is state).
struct TruncNormal {
    double stddev, lo, hi;
    bool none;
    std::normal_distribution<double> nd;
    void init(double sigma) {
        stddev = clamp_value(std::abs(sigma), 0.0, 1.0);  // unc:61
        none = (std::abs(stddev) == 0.0);
        if (!none) {
            lo = (-1.0 - 0.0) / stddev;
            hi = (1.0 - 0.0) / stddev;
        }
        nd = std::normal_distribution<double>(0.0, 1.0);
    }
    double operator()(std::mt19937_64& rng) {
        if (none) return 0.0;
        while (true) {
            const double draw = nd(rng);
            if ((draw <= hi) && (draw >= lo)) return 0.0 + stddev * draw;
        }
    }
};

enum { ORACLE_NOISE_MT19937 = 3 };

struct NoiseSource {
    int mode;
    // injected
    const double* tape;
    uint64_t tape_pos, tape_end;
    bool exhausted;
    // mt19937
    std::mt19937_64* rng;
    TruncNormal tn[kMaxDof];
    // philox
    uint64_t seed, particle_id;
    // recording
    std::vector<double>* record;
    std::vector<uint64_t>* record_decisions = nullptr;  // decision tape (fks_noise_tape.decisions), see QrInfo
    double* max_condition = nullptr;                    // largest QrInfo::cond_est among this particle's solves
    const uint64_t* dec_tape = nullptr;                 // replayed decision tape (studies with the device-solver model)
    uint64_t dec_pos = 0, dec_end = 0;

    void begin_step(const Robot& robot) {
        // each controller step works on a fresh Clone() of the robot (spcs:1548), whose actuator
        // distributions have never drawn: the cached normal variate does not survive a step.
        if (mode == ORACLE_NOISE_MT19937)
            for (int i = 0; i < robot.D; i++) tn[i].init(robot.mdl->axes[(size_t)i].noise_sigma);
    }
    void draw(const Robot& robot, uint32_t step, uint32_t micro, double* out) {
        for (int i = 0; i < robot.D; i++) {
            double v = 0.0;
            if (mode == FKS_NOISE_INJECTED) {
                if (tape_pos < tape_end) v = tape[tape_pos++];
                else exhausted = true;
            } else if (mode == ORACLE_NOISE_MT19937) {
                v = tn[i](*rng);
            } else if (mode == FKS_NOISE_PHILOX) {
                v = fks_philox_truncated_normal(seed, particle_id, step, micro, (uint32_t)i,
                                                robot.mdl->axes[(size_t)i].noise_sigma);
            }
            out[i] = v;
            if (record) record->push_back(v);
        }
    }
};

// debug capture of the stacked systems the resolver solves (oracle_debug_capture_systems): rows, cols, A column major, b
struct CapturedSystem {
    int rows, cols;
    std::vector<double> A, b;
};
std::vector<CapturedSystem> g_captured;
size_t g_capture_limit = 0;
void capture_system(const double* A, const double* b, int rows, int cols) {
    if (g_capture_limit == 0) return;
#pragma omp critical(fks_capture)
    {
        if (g_captured.size() < g_capture_limit) {
            CapturedSystem cs;
            cs.rows = rows;
            cs.cols = cols;
            cs.A.assign(A, A + (size_t)rows * cols);
            cs.b.assign(b, b + rows);
            g_captured.push_back(std::move(cs));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The simulator (spcs:371-1999)
// ------------------------------------------------------------------------------------------------
struct SelfKey {
    int64_t x, y, z;
    bool operator<(const SelfKey& o) const {
        if (x != o.x) return x < o.x;
        if (y != o.y) return y < o.y;
        return z < o.z;
    }
};
typedef std::map<std::pair<int, int>, V3> SelfMap;  // (link, point-in-link) -> correction

struct Sim {
    Env env;
    Model model;
    Robot proto;
    fks_solver_params sp;
    double freq, interval;
    uint64_t seed;
    int num_threads;
    std::vector<std::mt19937_64> rngs;
    uint64_t stats[FKS_NUM_STATS];
    // recorded tape of the last call
    std::vector<std::vector<double>> recorded;
    std::vector<std::vector<uint64_t>> recorded_decisions;
    std::vector<uint32_t> sensitivity;
    std::vector<double> max_condition;
    double decision_cond_limit = INFINITY;  // solves whose cond_est exceeds it are recorded with their solution (oracle_set_decision_cond_limit)
    int qr_model = -1;  // >= 0: solve with the CPU model of the DEVICE solver (fks_qr_model.cpp), option bits = value (studies only)

    // CheckEnvironmentCollision (spcs:921-981)
    bool check_env(const Robot& r, double collision_threshold, uint32_t* sens) const {
        const double res = env.d.sdf_resolution;
        const double thr = collision_threshold - (sp.environment_collision_check_tolerance * res);
        bool hit = false;
        for (size_t p = 0; p < r.mdl->points.size(); p++) {
            const V3 pw = iso_apply(r.link_T[(size_t)r.mdl->point_link[p]], r.mdl->points[p]);
            const std::pair<float, bool> chk = env.get4d(pw, sens);
            if ((double)chk.first < thr) {
                if ((double)chk.first < (thr - res)) {
                    hit = true;
                } else {
                    const double est = env.estimate_distance(pw, sens).first;
                    if (sens && std::abs(est - thr) < 1e-9 * res) *sens |= SENS_EST_THRESHOLD;
                    if (est < thr) hit = true;
                }
            }
            // no early return: the sensitivity scan wants every point (result is an OR, spcs:963,974)
        }
        return hit;
    }

    // CollectSelfCollisions (spcs:1183-1275) + ExtractSelfCollidingPoints (spcs:983-1171)
    void collect_self(const Robot& prev, const Robot& cur, double time_interval, SelfMap* out, uint32_t* sens) const {
        out->clear();
        const int L = cur.L;
        if (L == 1) return;
        if (L == 2 && cur.mdl->allowed[0 * L + 1]) return;
        std::map<SelfKey, std::vector<std::pair<int, int>>> cells;
        bool any_candidate = false;
        const double res = env.d.map_resolution;
        for (int l = 0; l < L; l++) {
            for (int p = cur.mdl->link_begin[(size_t)l]; p < cur.mdl->link_begin[(size_t)l + 1]; p++) {
                const V3 pw = iso_apply(cur.link_T[(size_t)l], cur.mdl->points[(size_t)p]);
                // LocationToExtendedGridIndex (spcs:1173-1181): DIVISION by the resolution, C-cast
                const V3 g = iso_apply(env.inv_origin, pw);
                const double gx = g.x / res, gy = g.y / res, gz = g.z / res;
                const SelfKey key = {(int64_t)gx, (int64_t)gy, (int64_t)gz};
                std::vector<std::pair<int, int>>& cell = cells[key];
                if (cell.size() > 1) any_candidate = true;
                else if (cell.size() == 1 && cell[0].first != l) any_candidate = true;
                cell.push_back({l, p - cur.mdl->link_begin[(size_t)l]});
                if (sens) {
                    if (std::abs(gx - std::nearbyint(gx)) < 1e-9 || std::abs(gy - std::nearbyint(gy)) < 1e-9 ||
                        std::abs(gz - std::nearbyint(gz)) < 1e-9)
                        *sens |= SENS_CELL_BOUNDARY;
                }
            }
        }
        if (!any_candidate) return;
        std::vector<double> mass((size_t)L);
        double acc = 0.0;
        for (int l = L - 1; l >= 0; l--) {  // spcs:1244-1255
            const double m = (double)(cur.mdl->link_begin[(size_t)l + 1] - cur.mdl->link_begin[(size_t)l]);
            mass[(size_t)l] = m + acc;
            acc += m;
        }
        for (auto& kv : cells) {
            const std::vector<std::pair<int, int>>& cand = kv.second;
            if (cand.size() <= 1) continue;
            std::map<int, std::vector<int>> by_link;
            for (auto& c : cand) by_link[c.first].push_back(c.second);
            if (by_link.size() < 2) continue;
            std::map<int, std::vector<int>> link_collisions;
            for (auto& f : by_link)
                for (auto& s : by_link)
                    if (f.first != s.first && !cur.mdl->allowed[(size_t)f.first * L + s.first])
                        link_collisions[f.first].push_back(s.first);
            if (link_collisions.size() < 2) continue;
            if (sens) *sens |= SENS_SELF_COLLISION;
            const double time_multiplier = 1.0 / time_interval;
            std::map<int, V3> momentum;
            for (auto& li : by_link) {
                const int l = li.first;
                if (link_collisions.find(l) == link_collisions.end()) continue;
                V3 m = {0.0, 0.0, 0.0};
                for (int pi : li.second) {
                    const V3 pl = cur.mdl->points[(size_t)(cur.mdl->link_begin[(size_t)l] + pi)];
                    const V3 motion = iso_apply(cur.link_T[(size_t)l], pl) - iso_apply(prev.link_T[(size_t)l], pl);
                    m = m + motion * time_multiplier;
                }
                momentum[l] = m;
            }
            for (auto& lc : link_collisions) {
                const int l = lc.first;
                const std::vector<int>& others = lc.second;
                const V3 anchor = iso_apply(prev.link_T[(size_t)l], cur.mdl->points[(size_t)(cur.mdl->link_begin[(size_t)l] + by_link[l].front())]);
                const double cnt = (double)by_link[l].size();
                const V3 v_i = momentum[l] * (1.0 / cnt) ;
                const int m = (int)others.size();
                if (m > 15) continue;
                V3 nrm[16];
                double A[16 * 16], rhs[16], Ainv[16 * 16];
                for (int a = 0; a < m; a++) {
                    const int ol = others[(size_t)a];
                    const V3 other_anchor = iso_apply(prev.link_T[(size_t)ol], cur.mdl->points[(size_t)(cur.mdl->link_begin[(size_t)ol] + by_link[ol].front())]);
                    nrm[a] = safe_normal(other_anchor - anchor);
                    const V3 v_a = momentum[ol] * (1.0 / (double)by_link[ol].size());
                    rhs[a] = dot(nrm[a], v_a - v_i);
                }
                for (int a = 0; a < m; a++)
                    for (int b = 0; b < m; b++) {
                        double v = dot(nrm[a], nrm[b]) / mass[(size_t)l];
                        if (a == b) v += dot(nrm[a], nrm[a]) / mass[(size_t)others[(size_t)a]];
                        A[a * m + b] = v;
                    }
                lu_inverse(A, m, Ainv);
                V3 corr = {0.0, 0.0, 0.0};
                for (int a = 0; a < m; a++) {
                    double lam = 0.0;
                    for (int b = 0; b < m; b++) lam += Ainv[a * m + b] * rhs[b];
                    corr = corr + nrm[a] * lam;
                }
                corr = corr * (1.0 / mass[(size_t)l]);
                const V3 per_point = corr * (1.0 / cnt);
                for (int pi : by_link[l]) (*out)[{l, pi}] = per_point;
            }
        }
    }

    // CheckSelfCollisions + CheckPointsForSelfCollision (spcs:1277-1396): true when some cell of edge check_resolution
    // holds points of two links whose pair is not allowed
    bool check_self_bool(const Robot& cur, double check_resolution) const {
        const int L = cur.L;
        if (L == 1) return false;
        if (L == 2 && cur.mdl->allowed[0 * L + 1]) return false;
        std::map<SelfKey, std::vector<int>> cells;  // cell -> links of its points
        for (int l = 0; l < L; l++)
            for (int p = cur.mdl->link_begin[(size_t)l]; p < cur.mdl->link_begin[(size_t)l + 1]; p++) {
                const V3 pw = iso_apply(cur.link_T[(size_t)l], cur.mdl->points[(size_t)p]);
                const V3 g = iso_apply(env.inv_origin, pw);
                const SelfKey key = {(int64_t)(g.x / check_resolution), (int64_t)(g.y / check_resolution), (int64_t)(g.z / check_resolution)};
                cells[key].push_back(l);
            }
        for (auto& kv : cells) {
            const std::vector<int>& links = kv.second;
            if (links.size() <= 1) continue;
            for (int a : links)
                for (int b : links)
                    if (a != b && !cur.mdl->allowed[(size_t)a * L + b]) return true;
        }
        return false;
    }

    // CheckConfigCollision (spcs:1398-1416)
    bool check_config_collision(const double* config, double inflation_ratio) const {
        Robot robot = proto;
        robot.set_position(config, nullptr);
        const double environment_collision_distance_threshold = inflation_ratio * env.d.map_resolution;
        const double self_collision_check_resolution = (inflation_ratio + 1.0) * env.d.map_resolution;
        const bool env_collision = check_env(robot, environment_collision_distance_threshold, nullptr);
        const bool self_collision = check_self_bool(robot, self_collision_check_resolution);
        return env_collision || self_collision;
    }

    // CheckCollision (spcs:1418-1436)
    bool check_collision(const Robot& prev, const Robot& cur, double time_interval, SelfMap* self, uint32_t* sens) const {
        const bool envc = check_env(cur, 0.0 /*contact_distance_threshold_, spcs:424*/, sens);
        collect_self(prev, cur, time_interval, self, sens);
        return envc || !self->empty();
    }

    // EstimateMaxControlInputWorkspaceMotion(start_robot, end_robot) (spcs:1492-1527)
    double max_motion(const Robot& a, const Robot& b) const {
        double mx = 0.0;
        for (size_t p = 0; p < a.mdl->points.size(); p++) {
            const int l = a.mdl->point_link[p];
            const V3 d = iso_apply(b.link_T[(size_t)l], a.mdl->points[p]) - iso_apply(a.link_T[(size_t)l], a.mdl->points[p]);
            const double sq = dot(d, d);
            if (sq > mx) mx = sq;
        }
        return std::sqrt(mx);
    }
    // (robot, control_input) overload (spcs:1538-1544): noiseless apply on a clone
    double max_motion_of_input(const Robot& r, const double* u, bool* nan_seen) const {
        Robot next = r;
        next.apply_control(u, nullptr, nan_seen, nullptr);
        return max_motion(r, next);
    }

    // CollectPointCorrectionsAndJacobians (spcs:1818-1939): rows appended link-major, point-minor
    void collect_corrections(const Robot& prev, const Robot& cur, const SelfMap& self, std::vector<double>* Jrows,
                             std::vector<double>* corr, uint32_t* flags, uint32_t* sens) const {
        Jrows->clear();
        corr->clear();
        const int D = cur.D;
        double Jp[3 * kMaxDof];
        for (int l = 0; l < cur.L; l++) {
            for (int p = cur.mdl->link_begin[(size_t)l]; p < cur.mdl->link_begin[(size_t)l + 1]; p++) {
                const V3 pl = cur.mdl->points[(size_t)p];
                bool have_self = false, have_env = false;
                V3 self_c = {0, 0, 0}, env_c = {0, 0, 0};
                auto it = self.find({l, p - cur.mdl->link_begin[(size_t)l]});
                if (it != self.end()) {
                    have_self = true;
                    self_c = it->second;
                }
                const V3 p_prev = iso_apply(prev.link_T[(size_t)l], pl);
                const V3 p_cur = iso_apply(cur.link_T[(size_t)l], pl);
                const std::pair<double, bool> chk = env.estimate_distance(p_cur, sens);
                if (chk.second && sens && std::abs(chk.first) < 1e-9 * env.d.sdf_resolution) *sens |= SENS_EST_ZERO;
                if (chk.first < 0.0 /*resolution_distance_threshold_, spcs:425*/ && chk.second) {
                    const V3 motion = p_cur - p_prev;
                    const V3 dir = safe_normal(motion);
                    V3 nrm;
                    bool would_assert = false;
                    const bool found = env.lookup_normal(p_cur, dir, &nrm, &would_assert, sens);
                    if (!found || would_assert) *flags |= FKS_FLAG_WOULD_ASSERT_NORMAL;  // spcs:1882, :115
                    const V3 g = safe_normal(nrm);
                    const double pen = std::abs(0.0 - chk.first);
                    env_c = g * pen;
                    have_env = true;
                }
                if (have_self || have_env) {
                    cur.point_jacobian(l, pl, Jp);
                    for (int i = 0; i < 3 * D; i++) Jrows->push_back(Jp[i]);
                    V3 pc = {0, 0, 0};
                    if (have_self) pc = pc + self_c;
                    if (have_env) pc = pc + env_c;
                    corr->push_back(pc.x);
                    corr->push_back(pc.y);
                    corr->push_back(pc.z);
                }
            }
        }
    }

    struct StepResult {
        bool collided, failed;
    };

    // ForwardSimulationStepTrace (filled at spcs:1583-1617, :1703, :1714, :1778), flat: one record per push_back /
    // assignment of the reference, in its order; layout = fks_trace_header + max(cfg_stride, D) doubles
    mutable std::vector<char>* trace = nullptr;  // set for a traced single-particle call only
    void trace_push(uint32_t kind, uint32_t step, uint32_t micro, uint32_t iter, const double* values, int n) const {
        if (!trace) return;
        const int width = std::max(proto.cfg_stride(), proto.D);
        fks_trace_header h = {kind, step, micro, iter};
        std::vector<double> v((size_t)width, 0.0);
        for (int i = 0; i < n; i++) v[(size_t)i] = values[i];
        const size_t at = trace->size();
        trace->resize(at + sizeof(h) + (size_t)width * 8);
        std::memcpy(trace->data() + at, &h, sizeof(h));
        std::memcpy(trace->data() + at + sizeof(h), v.data(), (size_t)width * 8);
    }

    // ResolveForwardSimulation (spcs:1546-1816).  `robot` enters at the step's start configuration and
    // leaves at the resolved configuration (the reference returns it and the caller SetPosition()s).
    StepResult resolve(Robot& robot, const double* control_input, double controller_interval, NoiseSource& noise,
                       bool allow_contacts, uint32_t step, uint32_t* flags, uint32_t* n_micro_total,
                       uint32_t* n_iter_total, uint64_t* st, uint32_t* sens) const {
        const int D = robot.D;
        double real_u[kMaxDof], du[kMaxDof];
        bool nan_seen = false;
        for (int i = 0; i < D; i++) real_u[i] = control_input[i] * controller_interval;
        const double map_res = env.d.map_resolution;
        const double computed_step_motion = max_motion_of_input(robot, real_u, &nan_seen);
        const double target_microstep_distance = map_res * 0.125;
        const double allowed_microstep_distance = map_res * 1.0;
        const double ratio = computed_step_motion / target_microstep_distance;
        if (sens && ratio > 1e-6 && std::abs(ratio - std::nearbyint(ratio)) < 1e-9) *sens |= SENS_NMICRO;
        const uint32_t number_microsteps = std::max(1u, (uint32_t)std::ceil(ratio));
        for (int i = 0; i < D; i++) du[i] = real_u[i] / (double)number_microsteps;
        const double computed_microstep_motion = max_motion_of_input(robot, du, &nan_seen);
        if (computed_microstep_motion > allowed_microstep_distance) *flags |= FKS_FLAG_WOULD_ASSERT_MICROSTEP;  // spcs:1570-1575
        bool collided = false;
        SelfMap self;
        std::vector<double> Jrows, corr, Acm, bvec;
        noise.begin_step(robot);
        trace_push(FKS_TRACE_CONTROL_INPUT, step, 0, 0, real_u, D);        // spcs:1583-1587
        trace_push(FKS_TRACE_CONTROL_INPUT_STEP, step, 0, 0, du, D);
        for (uint32_t micro = 0; micro < number_microsteps; micro++) {
            (*n_micro_total)++;
            const Robot previous = robot;  // previous_configuration (spcs:1597)
            double tn[kMaxDof];
            noise.draw(robot, step, micro, tn);
            robot.apply_control(du, tn, &nan_seen, sens);  // spcs:1599-1601
            bool in_collision = check_collision(previous, robot, controller_interval, &self, sens);  // spcs:1608
            if (in_collision) collided = true;
            trace_push(FKS_TRACE_POST_ACTION, step, micro, 0, robot.cfg, robot.cfg_stride());  // spcs:1615-1618
            if (in_collision && allow_contacts) {
                uint32_t resolver_iterations = 0;
                double scaling = sp.resolve_correction_initial_step_size;
                while (in_collision) {
                    collect_corrections(previous, robot, self, &Jrows, &corr, flags, sens);
                    const int rows = (int)corr.size();
                    st[FKS_STAT_TOTAL_CORRECTED_POINTS] += (uint64_t)(rows / 3);
#pragma omp atomic
                    g_rows_hist[std::min(15, rows / 24)]++;
                    double raw[kMaxDof];
                    for (int i = 0; i < D; i++) raw[i] = 0.0;
                    if (rows == 0) {
                        // Eigen: colPivHouseholderQr of a 0x0 matrix, solve -> empty vector; the reference
                        // would then fail the size assert in ApplyControlInput.  Device behaviour: zero step.
                        *flags |= FKS_FLAG_EMPTY_JACOBIAN;
                    } else {
                        Acm.resize((size_t)rows * D);
                        bvec = corr;
                        for (int r = 0; r < rows; r++)
                            for (int c = 0; c < D; c++) Acm[(size_t)c * rows + r] = Jrows[(size_t)r * D + c];
                        capture_system(Acm.data(), bvec.data(), rows, D);
                        if (qr_model >= 0 && (qr_model & 4) && noise.dec_tape) {
                            // what the GPU does in parity mode: the decisions (and flagged solutions) of the recorded run
                            // injected into the device solver's arithmetic
                            const uint64_t rec_words = 2 + (uint64_t)D;
                            if (noise.dec_pos < noise.dec_end) {
                                const uint64_t* rec = noise.dec_tape + noise.dec_pos * rec_words;
                                noise.dec_pos++;
                                const int rank = (int)(rec[0] & 0xFFull);
                                if ((int)((rec[0] >> 16) & 0xFFFFFFFFull) != rows) *flags |= FKS_FLAG_DECISION_DESYNC;
                                if (rec[0] & FKS_DECISION_OVERRIDE_SOLUTION) {
                                    for (int i = 0; i < D; i++) std::memcpy(&raw[i], &rec[2 + i], 8);
                                } else {
                                    oracle_qr_device_model_forced(Acm.data(), bvec.data(), rows, D, qr_model & 3, rec[1], rank, raw);
                                }
                            } else {
                                *flags |= FKS_FLAG_DECISION_DESYNC;
                                oracle_qr_device_model(Acm.data(), bvec.data(), rows, D, qr_model & 3, raw, nullptr);
                            }
                        } else if (qr_model >= 0) {
                            oracle_qr_device_model(Acm.data(), bvec.data(), rows, D, qr_model, raw, nullptr);
                        } else {
                            QrInfo qi;
                            colpiv_qr_solve(Acm.data(), bvec.data(), rows, D, raw, sens, &qi);  // spcs:1629,1990-1998
                            if (noise.max_condition && !qi.roundoff_kept) *noise.max_condition = std::max(*noise.max_condition, qi.cond_est);
                            if (noise.record_decisions) {
                                // one record of 2 + D words per solve (fksgpu.h, fks_noise_tape.decisions)
                                std::vector<uint64_t>& dv = *noise.record_decisions;
                                uint64_t w0 = (uint64_t)qi.rank | ((uint64_t)rows << 16);
                                if (qi.roundoff_kept || qi.cond_est > decision_cond_limit) w0 |= FKS_DECISION_OVERRIDE_SOLUTION;
                                if (qi.cond_est > kIllConditioned && qi.cond_est <= decision_cond_limit && !qi.roundoff_kept && sens)
                                    *sens |= SENS_ILL_CONDITIONED;
                                if (qi.roundoff_seen) w0 |= FKS_DECISION_ROUNDOFF_PIVOT;
                                if (qi.tie) w0 |= FKS_DECISION_PIVOT_TIE;
                                dv.push_back(w0);
                                dv.push_back(qi.order);
                                for (int i = 0; i < D; i++) {
                                    uint64_t bits;
                                    std::memcpy(&bits, &raw[i], 8);
                                    dv.push_back(bits);
                                }
                            }
                        }
                    }
                    static const bool trace = std::getenv("FKS_ORACLE_TRACE") != nullptr;
                    if (trace) {
                        std::fprintf(stderr, "step %u micro %u iter %u rows %d scaling %.4f corr:", step, micro, resolver_iterations, rows, scaling);
                        for (int i = 0; i < rows && i < 9; i++) std::fprintf(stderr, " %.3e", corr[(size_t)i]);
                        std::fprintf(stderr, " | raw:");
                        for (int i = 0; i < D; i++) std::fprintf(stderr, " %.3e", raw[i]);
                        std::fprintf(stderr, "\n");
                    }
                    const double est = max_motion_of_input(robot, raw, &nan_seen);  // spcs:1630
                    const double frac_raw = est / allowed_microstep_distance;
                    if (sens && std::abs(frac_raw - 1.0) < 1e-9) *sens |= SENS_STEP_FRACTION;
                    const double step_fraction = std::max(frac_raw, 1.0);  // spcs:1681
                    double real_step[kMaxDof];
                    for (int i = 0; i < D; i++) real_step[i] = (raw[i] / step_fraction) * std::abs(scaling);  // spcs:1682
                    robot.apply_control(real_step, nullptr, &nan_seen, sens);  // spcs:1689
                    in_collision = check_collision(previous, robot, controller_interval, &self, sens);  // spcs:1694-1698
                    resolver_iterations++;
                    (*n_iter_total)++;
                    trace_push(FKS_TRACE_RESOLUTION_STEP, step, micro, resolver_iterations, robot.cfg, robot.cfg_stride());  // spcs:1703
                    if (resolver_iterations > sp.max_resolver_iterations) {  // spcs:1705-1746
                        trace_push(FKS_TRACE_RETURNED_PREVIOUS, step, micro, resolver_iterations, previous.cfg, previous.cfg_stride());  // spcs:1714
                        st[FKS_STAT_UNSUCCESSFUL_RESOLVES]++;
                        if (!self.empty()) st[FKS_STAT_UNSUCCESSFUL_SELF_COLLISION_RESOLVES]++;
                        else st[FKS_STAT_UNSUCCESSFUL_ENV_COLLISION_RESOLVES]++;
                        robot = previous;
                        if (nan_seen) *flags |= FKS_FLAG_WOULD_ASSERT_NAN;
                        return {true, true};
                    }
                    if ((resolver_iterations % sp.resolve_correction_step_scaling_decay_iterations) == 0) {  // spcs:1747-1761
                        if (scaling >= 0.0) {
                            scaling = scaling * sp.resolve_correction_step_scaling_decay_rate;
                            if (scaling < sp.resolve_correction_min_step_scaling) scaling = -sp.resolve_correction_min_step_scaling;
                        } else {
                            scaling = -sp.resolve_correction_min_step_scaling;
                        }
                    }
                }
            } else if (in_collision && !allow_contacts) {  // spcs:1769-1786
                trace_push(FKS_TRACE_RETURNED_PREVIOUS, step, micro, 0, previous.cfg, previous.cfg_stride());  // spcs:1778
                st[FKS_STAT_SUCCESSFUL_RESOLVES]++;
                robot = previous;
                if (nan_seen) *flags |= FKS_FLAG_WOULD_ASSERT_NAN;
                return {true, false};
            }
        }
        st[FKS_STAT_SUCCESSFUL_RESOLVES]++;  // spcs:1802-1814
        if (collided) st[FKS_STAT_COLLISION_RESOLVES]++;
        else st[FKS_STAT_FREE_RESOLVES]++;
        if (nan_seen) *flags |= FKS_FLAG_WOULD_ASSERT_NAN;
        return {collided, false};
    }

    // ForwardSimulateRobot + ForwardSimulateMutableRobot (spcs:824-829, 843-919)
    void simulate_particle(const double* start, const double* target, bool allow_contacts, NoiseSource& noise,
                           double* out_cfg, fks_result_tail* tail, uint64_t* st, uint32_t* sens) const {
        Robot robot = proto;  // Clone (spcs:826)
        Pid pids[kMaxDof];
        for (int i = 0; i < robot.D; i++) {
            const fks_axis_params& a = robot.mdl->axes[(size_t)i];
            pids[i].init(a.kp, a.ki, a.kd, a.integral_clamp);
        }
        robot.set_position(start, sens);  // ResetPosition (tnuva:139-143): zero controllers + SetPosition
        bool collided = false, any_resolve_failed = false;
        uint32_t flags = 0, n_micro = 0, n_iter = 0, n_steps = 0;
        const uint32_t steps = std::max((uint32_t)(sp.forward_simulation_time * freq), 1u);  // spcs:856
        for (uint32_t step = 0; step < steps; step++) {
            n_steps++;
            double action[kMaxDof];
            bool nan_seen = false;
            robot.control_action(target, interval, pids, action, &nan_seen);  // spcs:868
            if (nan_seen) flags |= FKS_FLAG_WOULD_ASSERT_NAN;
            Robot work = robot;  // ResolveForwardSimulation clones (spcs:1548)
            const StepResult res = resolve(work, action, interval, noise, allow_contacts, step, &flags, &n_micro, &n_iter, st, sens);
            if (allow_contacts || !res.collided) {  // spcs:873
                robot.set_position(work.cfg, sens);  // spcs:875
                if (res.collided) collided = true;
                if (res.failed) {
                    flags |= FKS_FLAG_RESOLVE_FAILED;
                    if (sp.failed_resolves_end_motion) {
                        flags |= FKS_FLAG_ENDED_BY_FAILURE;
                        break;
                    }
                    any_resolve_failed = true;
                } else if (any_resolve_failed) {
                    st[FKS_STAT_RECOVERED_UNSUCCESSFUL_RESOLVES]++;
                }
                const double target_distance = robot.distance_to(target);  // spcs:898
                if (target_distance < sp.simulation_shortcut_distance) {
                    flags |= FKS_FLAG_ENDED_BY_SHORTCUT;
                    break;
                }
            } else {
                flags |= FKS_FLAG_ENDED_BY_NOCONTACT;
                // The reference's SimulationResult carries the local `collided`, which is only set inside
                // the branch above (spcs:877-880): a no-contact stop reports did_contact == false.
                break;
            }
        }
        if (collided) flags |= FKS_FLAG_DID_CONTACT;
        if (noise.mode == FKS_NOISE_INJECTED && noise.exhausted) flags |= FKS_FLAG_TAPE_EXHAUSTED;
        const int stride = robot.cfg_stride();
        for (int i = 0; i < stride; i++) out_cfg[i] = robot.cfg[i];
        tail->flags = flags;
        tail->n_microsteps = n_micro;
        tail->n_resolver_iters = n_iter;
        tail->n_steps = n_steps;
        st[FKS_STAT_TOTAL_MICROSTEPS] += n_micro;
        st[FKS_STAT_TOTAL_RESOLVER_ITERATIONS] += n_iter;
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// C interface for ctypes
// ------------------------------------------------------------------------------------------------
extern "C" {

struct oracle_sim {
    Sim s;
};

int oracle_noise_mode_mt19937(void) { return ORACLE_NOISE_MT19937; }

oracle_sim* oracle_create(const fks_env_desc* env, const fks_robot_desc* robot, const fks_solver_params* params,
                          double simulation_controller_frequency, uint64_t prng_seed, int num_threads) {
    oracle_sim* o = new oracle_sim();
    o->s.env.init(env);
    if (!o->s.model.init(robot)) {
        delete o;
        return nullptr;
    }
    o->s.proto.bind(&o->s.model);
    o->s.sp = *params;
    o->s.freq = std::abs(simulation_controller_frequency);      // spcs:426
    o->s.interval = 1.0 / simulation_controller_frequency;       // spcs:427 (sign kept)
    o->s.seed = prng_seed;
#ifdef _OPENMP
    o->s.num_threads = num_threads > 0 ? num_threads : omp_get_max_threads();
#else
    o->s.num_threads = 1;
#endif
    // spcs:431-441: one RNG per thread, seeded through uniform_int_distribution<uint64_t>
    std::mt19937_64 prng(prng_seed);
    std::uniform_int_distribution<uint64_t> seed_dist(0, std::numeric_limits<uint64_t>::max());
    for (int t = 0; t < o->s.num_threads; t++) o->s.rngs.push_back(std::mt19937_64(seed_dist(prng)));
    std::memset(o->s.stats, 0, sizeof(o->s.stats));
    return o;
}

void oracle_destroy(oracle_sim* o) { delete o; }

int oracle_num_threads(const oracle_sim* o) { return o->s.num_threads; }

int oracle_config_stride(const oracle_sim* o) { return o->s.proto.cfg_stride(); }

// ForwardSimulateRobots (spcs:788-804).  noise_mode: FKS_NOISE_* or ORACLE_NOISE_MT19937 (the
// reference's own generator: per-thread std::mt19937_64, static schedule).  record_tape != 0 keeps
// every draw for oracle_copy_tape.  results: n records of (8*stride + 16) bytes.
int oracle_forward_simulate(oracle_sim* o, const double* starts, const double* targets, size_t n, size_t n_targets,
                            int allow_contacts, int noise_mode, const fks_noise_tape* tape, uint64_t first_particle_id,
                            int record_tape, void* results) {
    Sim& s = o->s;
    if (n > 0 && !(n_targets == 1 || n_targets == n)) return FKS_ERR_INVALID_ARGUMENT;  // assert spcs:792
    if (noise_mode == FKS_NOISE_INJECTED && !tape) return FKS_ERR_INVALID_ARGUMENT;
    const int stride = s.proto.cfg_stride();
    const size_t rec = (size_t)stride * 8 + sizeof(fks_result_tail);
    s.recorded.assign(record_tape ? n : 0, std::vector<double>());
    s.recorded_decisions.assign(record_tape ? n : 0, std::vector<uint64_t>());
    s.sensitivity.assign(n, 0);
    s.max_condition.assign(n, 1.0);
    const int T = s.num_threads;
    std::vector<std::vector<uint64_t>> tstats((size_t)T, std::vector<uint64_t>(FKS_NUM_STATS, 0));
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t idx = 0; idx < (int64_t)n; idx++) {
#ifdef _OPENMP
        const int th = omp_get_thread_num();
#else
        const int th = 0;
#endif
        NoiseSource ns;
        ns.mode = noise_mode;
        ns.tape = nullptr;
        ns.tape_pos = ns.tape_end = 0;
        ns.exhausted = false;
        ns.rng = &s.rngs[(size_t)th];
        ns.seed = s.seed;
        ns.particle_id = first_particle_id + (uint64_t)idx;
        ns.record = record_tape ? &s.recorded[(size_t)idx] : nullptr;
        ns.record_decisions = record_tape ? &s.recorded_decisions[(size_t)idx] : nullptr;
        ns.max_condition = &s.max_condition[(size_t)idx];
        if (noise_mode == FKS_NOISE_INJECTED) {
            ns.tape = tape->draws;
            ns.tape_pos = tape->offsets[idx];
            ns.tape_end = tape->offsets[idx + 1];
            if (tape->decisions && tape->decision_offsets) {
                ns.dec_tape = tape->decisions;
                ns.dec_pos = tape->decision_offsets[idx];
                ns.dec_end = tape->decision_offsets[idx + 1];
            }
        }
        const double* start = starts + (size_t)idx * stride;
        const double* target = targets + (n_targets == n ? (size_t)idx * stride : 0);
        char* out = (char*)results + (size_t)idx * rec;
        double cfg[kMaxDof];
        fks_result_tail tail;
        s.simulate_particle(start, target, allow_contacts != 0, ns, cfg, &tail, tstats[(size_t)th].data(), &s.sensitivity[(size_t)idx]);
        std::memcpy(out, cfg, (size_t)stride * 8);
        std::memcpy(out + (size_t)stride * 8, &tail, sizeof(tail));
    }
    for (int t = 0; t < T; t++)
        for (int k = 0; k < FKS_NUM_STATS; k++) s.stats[k] += tstats[(size_t)t][(size_t)k];
    return FKS_OK;
}

// ForwardSimulateRobot with enable_tracing (spcs:824-829) for one particle; returns the number of trace records and copies
// up to `capacity` of them (fks_trace_header + max(stride, D) doubles each)
size_t oracle_trace_stride(const oracle_sim* o) { return sizeof(fks_trace_header) + (size_t)std::max(o->s.proto.cfg_stride(), o->s.proto.D) * 8; }
size_t oracle_forward_simulate_traced(oracle_sim* o, const double* start, const double* target, int allow_contacts, int noise_mode,
                                      const fks_noise_tape* tape, uint64_t particle_id, void* result, void* records, size_t capacity) {
    Sim& s = o->s;
    std::vector<char> buf;
    s.trace = &buf;
    NoiseSource ns;
    ns.mode = noise_mode;
    ns.tape = nullptr;
    ns.tape_pos = ns.tape_end = 0;
    ns.exhausted = false;
    ns.rng = &s.rngs[0];
    ns.seed = s.seed;
    ns.particle_id = particle_id;
    ns.record = nullptr;
    if (noise_mode == FKS_NOISE_INJECTED && tape) {
        ns.tape = tape->draws;
        ns.tape_pos = tape->offsets[0];
        ns.tape_end = tape->offsets[1];
    }
    double cfg[kMaxDof];
    fks_result_tail tail;
    std::vector<uint64_t> st(FKS_NUM_STATS, 0);
    uint32_t sens = 0;
    s.simulate_particle(start, target, allow_contacts != 0, ns, cfg, &tail, st.data(), &sens);
    s.trace = nullptr;
    const int stride = s.proto.cfg_stride();
    std::memcpy(result, cfg, (size_t)stride * 8);
    std::memcpy((char*)result + (size_t)stride * 8, &tail, sizeof(tail));
    const size_t rec = oracle_trace_stride(o), n = buf.size() / rec;
    std::memcpy(records, buf.data(), std::min(n, capacity) * rec);
    return n;
}

// CheckConfigCollision (spcs:1398-1416) for n configurations
void oracle_check_config_collision(const oracle_sim* o, const double* configs, size_t n, double inflation_ratio, uint8_t* out) {
    const int stride = o->s.proto.cfg_stride();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) out[i] = o->s.check_config_collision(configs + (size_t)i * stride, inflation_ratio) ? 1 : 0;
}

// SURVEY 8(f)-3, the first consumer of a batch (uncertainty_planning_core.cpp:97-99): particles split by did_contact
// (ascending ids inside each part, no-contact part first) and their pairwise configuration distances
// (robot->ComputeConfigurationDistanceTo, the call at spcs:898; RESTATEMENT of the upstream robot classes like distance_to).
void oracle_end_states_partition(const uint32_t* flags, size_t n, uint32_t* order, uint64_t* counts2) {
    size_t k = 0;
    for (size_t i = 0; i < n; i++)
        if (!(flags[i] & FKS_FLAG_DID_CONTACT)) order[k++] = (uint32_t)i;
    counts2[0] = k;
    for (size_t i = 0; i < n; i++)
        if (flags[i] & FKS_FLAG_DID_CONTACT) order[k++] = (uint32_t)i;
    counts2[1] = n - counts2[0];
}
void oracle_pairwise_config_distance(const oracle_sim* o, const double* configs, size_t n, double* out) {
    const int stride = o->s.proto.cfg_stride();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Robot robot = o->s.proto;
        robot.set_position(configs + (size_t)i * stride, nullptr);
        for (size_t j = 0; j < n; j++) out[(size_t)i * n + j] = robot.distance_to(configs + j * stride);
    }
}

uint64_t oracle_tape_total(const oracle_sim* o) {
    uint64_t t = 0;
    for (auto& v : o->s.recorded) t += v.size();
    return t;
}
// offsets has n+1 entries
void oracle_copy_tape(const oracle_sim* o, double* draws, uint64_t* offsets) {
    uint64_t pos = 0;
    size_t i = 0;
    for (auto& v : o->s.recorded) {
        offsets[i++] = pos;
        std::memcpy(draws + pos, v.data(), v.size() * sizeof(double));
        pos += v.size();
    }
    offsets[i] = pos;
}
// decision tape of the last recorded call: total 64-bit words, then the flat words + per-particle offsets in RECORDS
uint64_t oracle_decision_words(const oracle_sim* o) {
    uint64_t t = 0;
    for (auto& v : o->s.recorded_decisions) t += v.size();
    return t;
}
void oracle_copy_decisions(const oracle_sim* o, uint64_t* words, uint64_t* offsets) {
    const uint64_t rec_words = 2 + (uint64_t)o->s.proto.D;
    uint64_t pos = 0;
    size_t i = 0;
    for (auto& v : o->s.recorded_decisions) {
        offsets[i++] = pos / rec_words;
        std::memcpy(words + pos, v.data(), v.size() * sizeof(uint64_t));
        pos += v.size();
    }
    offsets[i] = pos / rec_words;
}
// studies: solve with the CPU model of the device solver (opt_bits >= 0) instead of the Eigen-semantics solver (-1)
void oracle_set_qr_model(oracle_sim* o, int opt_bits) { o->s.qr_model = opt_bits; }
// decision tape: also record (FKS_DECISION_OVERRIDE_SOLUTION) the solution of every solve whose condition estimate exceeds `limit`
void oracle_set_decision_cond_limit(oracle_sim* o, double limit) { o->s.decision_cond_limit = limit; }
// studies: keep the next `limit` stacked systems the resolver solves
void oracle_debug_capture_systems(size_t limit) {
    g_captured.clear();
    g_capture_limit = limit;
}
size_t oracle_debug_captured_count(void) { return g_captured.size(); }
// sizes of system i; with non-null pointers also copies A (rows x cols column major) and b
void oracle_debug_captured_system(size_t i, int* rows, int* cols, double* A, double* b) {
    const CapturedSystem& cs = g_captured[i];
    *rows = cs.rows;
    *cols = cs.cols;
    if (A) std::memcpy(A, cs.A.data(), cs.A.size() * sizeof(double));
    if (b) std::memcpy(b, cs.b.data(), cs.b.size() * sizeof(double));
}
// per particle of the last call: the largest condition estimate among its solves (1 when it never solved)
void oracle_copy_max_condition(const oracle_sim* o, double* out) {
    std::memcpy(out, o->s.max_condition.data(), o->s.max_condition.size() * sizeof(double));
}
void oracle_copy_sensitivity(const oracle_sim* o, uint32_t* out) {
    std::memcpy(out, o->s.sensitivity.data(), o->s.sensitivity.size() * sizeof(uint32_t));
}
void oracle_get_statistics(const oracle_sim* o, uint64_t* out) { std::memcpy(out, o->s.stats, sizeof(o->s.stats)); }
void oracle_reset_statistics(oracle_sim* o) { std::memset(o->s.stats, 0, sizeof(o->s.stats)); }

void oracle_debug_rows_hist(long long* out16, int reset) {
    for (int i = 0; i < 16; i++) {
        out16[i] = g_rows_hist[i];
        if (reset) g_rows_hist[i] = 0;
    }
}
void oracle_debug_rank_hist(long long* out42, int reset) {
    for (int i = 0; i < 42; i++) {
        out42[i] = g_rank_hist[i];
        if (reset) g_rank_hist[i] = 0;
    }
}
// ---- unit-test entry points for the restated primitives -----------------------------------------
void oracle_colpiv_qr_solve(const double* A_colmajor, const double* b, int rows, int cols, double* x, uint32_t* sens) {
    std::vector<double> A(A_colmajor, A_colmajor + (size_t)rows * cols), bb(b, b + rows);
    colpiv_qr_solve(A.data(), bb.data(), rows, cols, x, sens);
}
// same, also returning the decisions: out4 = {rank, size, order nibbles, flags (1 round-off pivot seen, 2 kept, 4 tie)}
void oracle_colpiv_qr_solve_info(const double* A_colmajor, const double* b, int rows, int cols, double* x, uint64_t* out4) {
    std::vector<double> A(A_colmajor, A_colmajor + (size_t)rows * cols), bb(b, b + rows);
    QrInfo qi;
    colpiv_qr_solve(A.data(), bb.data(), rows, cols, x, nullptr, &qi);
    out4[0] = (uint64_t)qi.rank;
    out4[1] = (uint64_t)qi.size;
    out4[2] = qi.order;
    out4[3] = (qi.roundoff_seen ? 1u : 0u) | (qi.roundoff_kept ? 2u : 0u) | (qi.tie ? 4u : 0u);
}
void oracle_pid_run(double kp, double ki, double kd, double iclamp, const double* errors, const double* timesteps, int n,
                    double* out) {
    Pid p;
    p.init(kp, ki, kd, iclamp);
    for (int i = 0; i < n; i++) out[i] = p.feedback(errors[i], timesteps[i]);
}
// the oracle's actuator (Robot::actuate_axis) on n (control, draw) pairs: noiseless and noisy value per pair -- pinned
// against the reference's own simple_uncertainty_models.hpp (oracle/_ref/unc_ref) by tests/test_oracle_actuator.py
void oracle_actuate_run(double velocity_limit, double proportional_noise, double minimum_noise, const double* controls,
                        const double* draws, int n, double* out_quiet, double* out_noisy) {
    fks_axis_params a;
    std::memset(&a, 0, sizeof(a));
    a.velocity_limit = velocity_limit;
    a.proportional_noise = proportional_noise;
    a.minimum_noise = minimum_noise;
    bool nan_seen = false;
    for (int i = 0; i < n; i++) {
        out_quiet[i] = Robot::actuate_axis(a, controls[i], nullptr, &nan_seen);
        out_noisy[i] = Robot::actuate_axis(a, controls[i], draws + i, &nan_seen);
    }
}
void oracle_exp_twist(const double* twist, double* out12) {
    const Iso T = exp_twist(twist);
    std::memcpy(out12, T.m, sizeof(T.m));
}
void oracle_twist_between(const double* a12, const double* b12, double* twist6) {
    Iso a, b;
    std::memcpy(a.m, a12, sizeof(a.m));
    std::memcpy(b.m, b12, sizeof(b.m));
    twist_between(a, b, twist6);
}
double oracle_wrap_angle(double v) { return wrap_angle(v); }
// truncated normal of the reference generator: n draws from TN(0, sigma, [-1, 1]) with mt19937_64(seed)
void oracle_truncated_normal(uint64_t seed, double sigma, int n, double* out) {
    std::mt19937_64 rng(seed);
    TruncNormal tn;
    tn.init(sigma);
    for (int i = 0; i < n; i++) out[i] = tn(rng);
}
void oracle_philox_raw(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
    fks_philox4x32_10(ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1], out4);
}
double oracle_philox_truncated_normal(uint64_t seed, uint64_t particle, uint32_t step, uint32_t micro, uint32_t dof, double sigma) {
    return fks_philox_truncated_normal(seed, particle, step, micro, dof, sigma);
}
// SDF queries against an environment description
int oracle_env_query(const fks_env_desc* env, const double* p3, float* raw, double* est, double* grad3) {
    Env e;
    e.init(env);
    const V3 p = {p3[0], p3[1], p3[2]};
    const std::pair<float, bool> r = e.get4d(p, nullptr);
    *raw = r.first;
    const std::pair<double, bool> d = e.estimate_distance(p, nullptr);
    *est = d.first;
    int64_t x, y, z;
    if (e.index_of(p, env->sdf_resolution, &x, &y, &z, nullptr)) e.gradient(x, y, z, grad3);
    return r.second ? 1 : 0;
}
// robot kinematics: link transforms (L*12) and the 3xD Jacobian of one point at a configuration
int oracle_robot_kinematics(const fks_robot_desc* robot, const double* cfg, int link, const double* point3,
                            double* link_transforms, double* jac3xD, double* stored_cfg) {
    Model m;
    if (!m.init(robot)) return 1;
    Robot r;
    r.bind(&m);
    r.set_position(cfg, nullptr);
    for (int l = 0; l < r.L; l++) std::memcpy(link_transforms + 12 * l, r.link_T[(size_t)l].m, 12 * sizeof(double));
    r.point_jacobian(link, {point3[0], point3[1], point3[2]}, jac3xD);
    for (int i = 0; i < r.cfg_stride(); i++) stored_cfg[i] = r.cfg[i];
    return 0;
}

}  // extern "C"
