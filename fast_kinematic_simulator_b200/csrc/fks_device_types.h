// Device-side data model of the particle contact simulator (shared by fks_kernels.cu and fks_api.cu).
//
// Layout in HBM (uploaded once per environment / robot, read-only afterwards):
//   SDF            float[nx*ny*nz], x-major (same linear index as VoxelGrid), L2-persisting window
//   normal table   open-addressing hash  (cell index + 1) -> (first entry, entry count); entries are
//                  6 doubles (entry direction xyz, normal xyz)
//   robot          one DevRobot struct + SoA point arrays (x[], y[], z[], link[]), staged into shared
//                  memory once by every CTA of the persistent kernel
//   particles      starts/targets as flat doubles (cfg_stride per particle), results as fixed records
//   scratch        one slot per resident warp: stacked Jacobian (column major) + self-collision work
#ifndef FKS_DEVICE_TYPES_H
#define FKS_DEVICE_TYPES_H

#include <stdint.h>

#include "fksgpu.h"

namespace fksdev {

constexpr int kMaxDof = 16;
constexpr int kMaxLinks = 16;
constexpr int kMaxJoints = 16;
constexpr int kMaxPairs = (kMaxLinks * (kMaxLinks - 1)) / 2;
constexpr int kMaxSelfPartners = kMaxLinks - 1;
constexpr int kWarpsPerBlock = 4;
constexpr int kThreadsPerBlock = kWarpsPerBlock * 32;

struct DevAxis {
    double kp, ki, kd, iclamp;  // |.| already applied (pid.hpp:104-113)
    double vlim;                // |velocity_limit| (unc.hpp:61)
    double pnoise, mnoise;      // |.| applied
    double sigma;
};

struct DevJoint {
    int parent, child, type, active;  // active = index among non-fixed joints, -1 if fixed
    double T[12];
    double axis[3];
    double lo, hi, weight;
};

struct DevRobot {
    int kind, L, J, D, P, n_pairs;
    double base[12];
    double pos_w, rot_w;
    DevAxis axes[kMaxDof];
    DevJoint joints[kMaxJoints];
    int active_joint[kMaxDof];        // active dof -> joint index
    int link_begin[kMaxLinks + 1];
    unsigned link_ancestors[kMaxLinks];  // bit j: joint j is on the path from the root to this link
    unsigned disallowed[kMaxLinks];      // bit j: self collision between this link and j is NOT allowed
    unsigned char pair_a[kMaxPairs], pair_b[kMaxPairs];  // the disallowed pairs, a < b
    double link_center[kMaxLinks][3];    // bounding sphere of the link's points (link frame)
    double link_radius[kMaxLinks];
    double link_mass[kMaxLinks];         // cumulative point counts (spcs.hpp:1244-1255)
};

struct DevEnv {
    double origin[12];
    double inv_origin[12];
    double map_res, sdf_res;
    double inv_sdf_res;      // 1.0 / sdf_res (VoxelGrid multiplies by the inverse cell size)
    double inv_twice_res;    // 1.0 / (2.0 * sdf_res)
    int nx, ny, nz;
    float oob;
    const float* sdf;
    // normal hash: keys hold (linear cell index + 1), 0 = empty slot
    const unsigned long long* nh_keys;
    const uint2* nh_vals;
    unsigned long long nh_mask;
    const double* normal_entries;  // 6 doubles: entry direction xyz, normal xyz
};

struct DevSolver {
    double interval;          // 1 / frequency, sign kept (spcs.hpp:427)
    double shortcut_distance;
    double check_tolerance;
    double decay_rate, initial_step, min_scaling;
    unsigned max_iters, decay_iters, n_steps;
    int failed_ends_motion;
};

struct LaunchArgs {
    const DevRobot* robot;
    const double* px;
    const double* py;
    const double* pz;
    const int* plink;
    DevEnv env;
    DevSolver sp;
    const double* starts;
    const double* targets;
    unsigned long long n_particles, n_targets;
    int allow_contacts, noise_mode;
    const double* tape;
    const unsigned long long* tape_off;
    unsigned long long seed, first_id;
    char* results;
    int cfg_stride, rec_stride;
    unsigned long long* stats;
    unsigned int* counter;
    char* scratch;                   // per-warp-slot global scratch
    unsigned long long scratch_bytes_per_warp;
    int ldj;                         // leading dimension (rows) of the stacked Jacobian columns
};

// hash of a linear cell index into the normal table (same function on host and device)
inline __host__ __device__ unsigned long long normal_hash(unsigned long long cell) {
    unsigned long long h = cell * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}

// ---- per-warp shared memory layout (doubles) -------------------------------------------------
struct WarpLayout {
    int Tprev, Tcur, Ttmp;   // L*12 each
    int M;                   // J*12 joint motion matrices
    int jaxis, jorig;        // J*3 each (world joint axes / origins for the Jacobian)
    int chain;               // 12
    int cfg, pcfg, tcfg;     // cfg_stride each (SE3: alias of the T arrays)
    int target;              // cfg_stride
    int scfg;                // cfg_stride: configuration at the start of the controller step (allow_contacts == false)
    int act, ru, du, tn, raw, stepv;  // D each
    int qr;                  // 3*D doubles + D ints (norms updated/direct, hcoeff, transpositions)
    int total;
};

inline __host__ __device__ WarpLayout make_warp_layout(int kind, int L, int J, int D, int stride) {
    WarpLayout w;
    int o = 0;
    w.Tprev = o; o += L * 12;
    w.Tcur = o; o += L * 12;
    w.Ttmp = o; o += L * 12;
    w.M = o; o += J * 12;
    w.jaxis = o; o += J * 3;
    w.jorig = o; o += J * 3;
    w.chain = o; o += 12;
    if (kind == FKS_ROBOT_SE3) {
        w.pcfg = w.Tprev; w.cfg = w.Tcur; w.tcfg = w.Ttmp;
    } else {
        w.cfg = o; o += stride;
        w.pcfg = o; o += stride;
        w.tcfg = o; o += stride;
    }
    w.target = o; o += stride;
    w.scfg = o; o += stride;
    w.act = o; o += D;
    w.ru = o; o += D;
    w.du = o; o += D;
    w.tn = o; o += D;
    w.raw = o; o += D;
    w.stepv = o; o += D;
    w.qr = o; o += 4 * D;
    w.total = (o + 1) & ~1;
    return w;
}

// ---- per-warp global scratch layout (bytes) --------------------------------------------------
struct ScratchLayout {
    unsigned long long jstore;    // (D+1) * ldj doubles
    unsigned long long selfcorr;  // 3*P doubles
    unsigned long long selfwork;  // small dense solve workspace
    unsigned long long keys;      // 3*P ints
    unsigned long long sflag;     // P bytes
    unsigned long long total;
    int ldj;
};

constexpr int kSelfWorkDoubles = kMaxSelfPartners * 5 + 4 + kMaxSelfPartners * (2 * kMaxSelfPartners) +
                                 kMaxSelfPartners * kMaxSelfPartners + 2 * kMaxSelfPartners;

inline __host__ __device__ ScratchLayout make_scratch_layout(int D, int P) {
    ScratchLayout s;
    s.ldj = ((3 * P + 3) / 4) * 4;
    unsigned long long o = 0;
    s.jstore = o; o += (unsigned long long)(D + 1) * s.ldj * 8;
    s.selfcorr = o; o += (unsigned long long)3 * P * 8;
    s.selfwork = o; o += (unsigned long long)kSelfWorkDoubles * 8;
    s.keys = o; o += (((unsigned long long)3 * P * 4 + 7) / 8) * 8;
    s.sflag = o; o += (((unsigned long long)P + 7) / 8) * 8;
    s.total = ((o + 127) / 128) * 128;
    return s;
}

// launch interface implemented in fks_kernels.cu
struct KernelInfo {
    int regs, static_smem, local_bytes, max_blocks_per_sm;
    size_t dyn_smem;
};
int simulate_kernel_info(int kind, size_t dyn_smem, KernelInfo* out);
int launch_simulate(int kind, const LaunchArgs& args, int grid, size_t dyn_smem, void* stream,
                    const void* l2_window_base, size_t l2_window_bytes);
size_t simulate_dyn_smem(int kind, int L, int J, int D, int P, int stride);
int launch_fp64_peak(double* out, int grid, int iters, void* stream);
int launch_gather(const float* data, unsigned long long n_mask, float* out, int grid, int iters, void* stream);

}  // namespace fksdev

#endif
