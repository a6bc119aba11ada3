// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see fks_oracle.cpp for the rules).
//
// Scalar restatement of simulator_environment_builder::BuildCompleteEnvironment
// (/root/reference/src/fast_kinematic_simulator/simulator_environment_builder.cpp:470-476, "envb" below), the checker of
// the device environment builder (fks_env_build_device) and of the host builder (fks_build_environment).
// It keeps the reference's own structure: DiscretizeObstacle materialises every sample (envb:21-46), BuildEnvironment
// runs the if / else-if bounds update over all of them (envb:85-124) and fills the grid by SetValue (envb:150-155),
// BuildSurfaceNormalsGrid keeps one std::vector of stored normals per cell, fills pass 1 from the SDF gradient
// (envb:262-277) and then walks the obstacles through the literal 26-way chain (envb:302-461) with
// UpdateSurfaceNormalGridCell (envb:162-187) clearing and rewriting cells in loop order.
//
// PARITY UNPINNED: the reference ships no fixtures for the builder and cannot be compiled here.  RESTATEMENT choices for
// the un-vendored dependencies (same as DESIGN.md section 2): VoxelGrid cell counts = ceil(size / res), index =
// (int64)(grid coordinate * (1 / res)); ExtractSignedDistanceField = distance to the nearest filled cell centre minus
// distance to the nearest free cell centre, as float, from an exact Euclidean transform -- computed here by three
// separable passes that minimise over the WHOLE line for every cell (no envelope, no windowing), i.e. a third algorithm
// next to the host's lower-envelope passes and the device's windowed search; GetGradient(x, y, z, true) = float central
// differences scaled in double, one-sided at the grid faces.

#include "fksgpu.h"

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

namespace {

struct Vec3 {
    double x, y, z;
};

inline Vec3 transform_point(const double* T, const Vec3& p) {
    return {T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7], T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]};
}
inline Vec3 rotate_vector(const double* T, const Vec3& p) {
    return {T[0] * p.x + T[1] * p.y + T[2] * p.z, T[4] * p.x + T[5] * p.y + T[6] * p.z, T[8] * p.x + T[9] * p.y + T[10] * p.z};
}

// StoredSurfaceNormal(normal3, direction3) (spcs.hpp:59-63): both SafeNormal'd, the direction as a 4-vector with w = 0
struct StoredNormal {
    double entry[4];
    double normal[3];
};
StoredNormal make_stored(const Vec3& normal, const Vec3& direction) {
    StoredNormal s;
    const double nn = std::sqrt(normal.x * normal.x + normal.y * normal.y + normal.z * normal.z);
    if (nn > DBL_EPSILON) {
        s.normal[0] = normal.x / nn;
        s.normal[1] = normal.y / nn;
        s.normal[2] = normal.z / nn;
    } else {
        s.normal[0] = normal.x;
        s.normal[1] = normal.y;
        s.normal[2] = normal.z;
    }
    const double d4[4] = {direction.x, direction.y, direction.z, 0.0};
    const double dn = std::sqrt(d4[0] * d4[0] + d4[1] * d4[1] + d4[2] * d4[2] + d4[3] * d4[3]);
    for (int i = 0; i < 4; i++) s.entry[i] = dn > DBL_EPSILON ? d4[i] / dn : d4[i];
    return s;
}

struct RawNormal {  // RawCellSurfaceNormal (envb.hpp): normal, entry direction
    Vec3 normal, entry;
};

struct VoxelFrame {  // VoxelGrid geometry (RESTATEMENT)
    double origin[12], inverse_origin[12];
    double res;
    int64_t nx, ny, nz;
    bool index_of(const Vec3& world, int64_t* x, int64_t* y, int64_t* z) const {
        const Vec3 g = transform_point(inverse_origin, world);
        const double inv = 1.0 / res;
        *x = (int64_t)(g.x * inv);
        *y = (int64_t)(g.y * inv);
        *z = (int64_t)(g.z * inv);
        return *x >= 0 && *y >= 0 && *z >= 0 && *x < nx && *y < ny && *z < nz;
    }
    size_t linear(int64_t x, int64_t y, int64_t z) const { return (size_t)((x * ny + y) * nz + z); }
};

// one separable pass: out[q] = min_p (in[p] + (q - p)^2) over the whole line
const int64_t kNoSeed = (int64_t)1 << 29;
void full_line_pass(std::vector<int64_t>& d2, int64_t n_lines_outer, int64_t outer_stride, int64_t n_lines_inner, int64_t inner_stride, int64_t n,
                    int64_t stride) {
#pragma omp parallel
    {
        std::vector<int64_t> in((size_t)n);
#pragma omp for collapse(2) schedule(static)
        for (int64_t a = 0; a < n_lines_outer; a++)
            for (int64_t b = 0; b < n_lines_inner; b++) {
                int64_t* line = &d2[(size_t)(a * outer_stride + b * inner_stride)];
                for (int64_t p = 0; p < n; p++) in[(size_t)p] = line[p * stride];
                for (int64_t q = 0; q < n; q++) {
                    int64_t best = kNoSeed;
                    for (int64_t p = 0; p < n; p++) {
                        if (in[(size_t)p] >= kNoSeed) continue;
                        const int64_t c = in[(size_t)p] + (q - p) * (q - p);
                        if (c < best) best = c;
                    }
                    line[q * stride] = best;
                }
            }
    }
}
void squared_distance_to_seeds(const std::vector<uint8_t>& seed, const VoxelFrame& f, std::vector<int64_t>& d2) {
    d2.resize(seed.size());
    for (size_t i = 0; i < seed.size(); i++) d2[i] = seed[i] ? 0 : kNoSeed;
    full_line_pass(d2, f.nx, f.ny * f.nz, f.ny, f.nz, f.nz, 1);      // along z
    full_line_pass(d2, f.nx, f.ny * f.nz, f.nz, 1, f.ny, f.nz);      // along y
    full_line_pass(d2, f.ny, f.nz, f.nz, 1, f.nx, f.ny * f.nz);      // along x
}

}  // namespace

struct oracle_env {
    fks_env_desc desc;
    std::vector<float> sdf;
    std::vector<uint8_t> occupancy;
    std::vector<int64_t> cell_index;
    std::vector<uint32_t> cell_start;
    std::vector<double> entries;
};

extern "C" {

void oracle_env_destroy(oracle_env* e) { delete e; }
const fks_env_desc* oracle_env_desc(const oracle_env* e) { return &e->desc; }
const uint8_t* oracle_env_occupancy(const oracle_env* e) { return e->occupancy.data(); }

oracle_env* oracle_build_environment(const fks_obstacle* obstacles, size_t n_obstacles, double resolution) {
    oracle_env* env = new oracle_env();
    VoxelFrame f;
    f.res = resolution;
    double x_min = 0.0, y_min = 0.0, z_min = 0.0, x_max = 0.0, y_max = 0.0, z_max = 0.0;
    double grid_x_size = 10.0, grid_y_size = 10.0, grid_z_size = 10.0;  // envb:54-56
    std::vector<Vec3> all_obstacle_cells;
    if (n_obstacles > 0) {
        bool xyz_bounds_initialized = false;
        for (size_t idx = 0; idx < n_obstacles; idx++) {
            const fks_obstacle& obstacle = obstacles[idx];
            // DiscretizeObstacle (envb:21-46)
            const double effective_resolution = resolution * 0.5;
            std::vector<Vec3> cells;
            const int32_t x_cells = (int32_t)(obstacle.extents[0] * 2.0 * (1.0 / effective_resolution));
            const int32_t y_cells = (int32_t)(obstacle.extents[1] * 2.0 * (1.0 / effective_resolution));
            const int32_t z_cells = (int32_t)(obstacle.extents[2] * 2.0 * (1.0 / effective_resolution));
            for (int32_t xidx = 0; xidx < x_cells; xidx++)
                for (int32_t yidx = 0; yidx < y_cells; yidx++)
                    for (int32_t zidx = 0; zidx < z_cells; zidx++) {
                        const double x_location = -(obstacle.extents[0] - (resolution * 0.5)) + (effective_resolution * xidx);
                        const double y_location = -(obstacle.extents[1] - (resolution * 0.5)) + (effective_resolution * yidx);
                        const double z_location = -(obstacle.extents[2] - (resolution * 0.5)) + (effective_resolution * zidx);
                        cells.push_back({x_location, y_location, z_location});
                    }
            for (size_t cidx = 0; cidx < cells.size(); cidx++) {
                const Vec3 real_location = transform_point(obstacle.pose, cells[cidx]);
                all_obstacle_cells.push_back(real_location);
                if (xyz_bounds_initialized) {  // envb:88-113
                    if (real_location.x < x_min) x_min = real_location.x;
                    else if (real_location.x > x_max) x_max = real_location.x;
                    if (real_location.y < y_min) y_min = real_location.y;
                    else if (real_location.y > y_max) y_max = real_location.y;
                    if (real_location.z < z_min) z_min = real_location.z;
                    else if (real_location.z > z_max) z_max = real_location.z;
                } else {
                    x_min = x_max = real_location.x;
                    y_min = y_max = real_location.y;
                    z_min = z_max = real_location.z;
                    xyz_bounds_initialized = true;
                }
            }
        }
        x_min -= (resolution * 0.5);  // envb:128-139
        y_min -= (resolution * 0.5);
        z_min -= (resolution * 0.5);
        x_min -= (resolution * 3.0);
        y_min -= (resolution * 3.0);
        z_min -= (resolution * 3.0);
        x_max += (resolution * 3.0);
        y_max += (resolution * 3.0);
        z_max += (resolution * 3.0);
        grid_x_size = x_max - x_min;
        grid_y_size = y_max - y_min;
        grid_z_size = z_max - z_min;
    }
    const double origin[12] = {1, 0, 0, x_min, 0, 1, 0, y_min, 0, 0, 1, z_min};
    const double inverse[12] = {1, 0, 0, -x_min, 0, 1, 0, -y_min, 0, 0, 1, -z_min};
    std::memcpy(f.origin, origin, sizeof(origin));
    std::memcpy(f.inverse_origin, inverse, sizeof(inverse));
    f.nx = (int64_t)std::ceil(grid_x_size / resolution);
    f.ny = (int64_t)std::ceil(grid_y_size / resolution);
    f.nz = (int64_t)std::ceil(grid_z_size / resolution);
    const size_t ncells = (size_t)(f.nx * f.ny * f.nz);
    env->occupancy.assign(ncells, 0);
    for (size_t idx = 0; idx < all_obstacle_cells.size(); idx++) {  // envb:150-155
        int64_t x, y, z;
        if (f.index_of(all_obstacle_cells[idx], &x, &y, &z)) env->occupancy[f.linear(x, y, z)] = 1;
    }
    all_obstacle_cells.clear();
    all_obstacle_cells.shrink_to_fit();

    // ExtractSignedDistanceField(+inf, {}, true, false) (envb:473) -- RESTATEMENT
    {
        std::vector<uint8_t> free_cells(ncells);
        for (size_t i = 0; i < ncells; i++) free_cells[i] = env->occupancy[i] ? 0 : 1;
        std::vector<int64_t> to_filled, to_free;
        squared_distance_to_seeds(env->occupancy, f, to_filled);
        squared_distance_to_seeds(free_cells, f, to_free);
        env->sdf.resize(ncells);
        for (size_t i = 0; i < ncells; i++) {
            const double distance_to_filled = std::sqrt((double)to_filled[i]) * resolution;
            const double distance_to_free = std::sqrt((double)to_free[i]) * resolution;
            env->sdf[i] = (float)(distance_to_filled - distance_to_free);
        }
    }
    const std::vector<float>& sdf = env->sdf;
    auto sdf_at = [&](int64_t x, int64_t y, int64_t z) -> float { return sdf[f.linear(x, y, z)]; };

    // BuildSurfaceNormalsGrid (envb:258-468)
    std::vector<std::vector<StoredNormal>> grid(ncells);
    for (int64_t x_idx = 0; x_idx < f.nx; x_idx++)
        for (int64_t y_idx = 0; y_idx < f.ny; y_idx++)
            for (int64_t z_idx = 0; z_idx < f.nz; z_idx++) {
                const float distance = sdf_at(x_idx, y_idx, z_idx);
                if (distance < 0.0) {
                    // GetGradient(x, y, z, true) -- RESTATEMENT
                    Vec3 gradient = {0.0, 0.0, 0.0};
                    if (x_idx > 0 && y_idx > 0 && z_idx > 0 && x_idx < f.nx - 1 && y_idx < f.ny - 1 && z_idx < f.nz - 1) {
                        const double inv_twice_resolution = 1.0 / (2.0 * resolution);
                        gradient.x = (double)(sdf_at(x_idx + 1, y_idx, z_idx) - sdf_at(x_idx - 1, y_idx, z_idx)) * inv_twice_resolution;
                        gradient.y = (double)(sdf_at(x_idx, y_idx + 1, z_idx) - sdf_at(x_idx, y_idx - 1, z_idx)) * inv_twice_resolution;
                        gradient.z = (double)(sdf_at(x_idx, y_idx, z_idx + 1) - sdf_at(x_idx, y_idx, z_idx - 1)) * inv_twice_resolution;
                    } else {
                        const int64_t low_x = x_idx > 0 ? x_idx - 1 : 0, high_x = x_idx < f.nx - 1 ? x_idx + 1 : f.nx - 1;
                        const int64_t low_y = y_idx > 0 ? y_idx - 1 : 0, high_y = y_idx < f.ny - 1 ? y_idx + 1 : f.ny - 1;
                        const int64_t low_z = z_idx > 0 ? z_idx - 1 : 0, high_z = z_idx < f.nz - 1 ? z_idx + 1 : f.nz - 1;
                        const double span_x = (double)(high_x - low_x) * resolution, span_y = (double)(high_y - low_y) * resolution,
                                     span_z = (double)(high_z - low_z) * resolution;
                        if (span_x > 0.0) gradient.x = ((double)sdf_at(high_x, y_idx, z_idx) - (double)sdf_at(low_x, y_idx, z_idx)) * (1.0 / span_x);
                        if (span_y > 0.0) gradient.y = ((double)sdf_at(x_idx, high_y, z_idx) - (double)sdf_at(x_idx, low_y, z_idx)) * (1.0 / span_y);
                        if (span_z > 0.0) gradient.z = ((double)sdf_at(x_idx, y_idx, high_z) - (double)sdf_at(x_idx, y_idx, low_z)) * (1.0 / span_z);
                    }
                    grid[f.linear(x_idx, y_idx, z_idx)].push_back(make_stored(gradient, {0.0, 0.0, 0.0}));
                }
            }
    // the 26 cases in the order of the reference's if / else-if chain (envb:302-461).  Condition per axis: 0 = index == 0,
    // 1 = index == count - 1, 2 = not tested.  Normals: Eigen's -UnitX() is (-1, -0, -0).
    struct Case {
        int cx, cy, cz;
    };
    static const Case chain[26] = {
        {0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1},  // corners :302-357
        {0, 0, 2}, {0, 1, 2}, {1, 0, 2}, {1, 1, 2},                                              // x-y edges :359-382
        {0, 2, 0}, {0, 2, 1}, {1, 2, 0}, {1, 2, 1},                                              // x-z edges :383-406
        {2, 0, 0}, {2, 0, 1}, {2, 1, 0}, {2, 1, 1},                                              // y-z edges :407-430
        {0, 2, 2}, {1, 2, 2}, {2, 0, 2}, {2, 1, 2}, {2, 2, 0}, {2, 2, 1}};                       // faces :432-461
    const Vec3 unit[3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
    auto negated = [](const Vec3& v) -> Vec3 { return {-v.x, -v.y, -v.z}; };
    for (size_t idx = 0; idx < n_obstacles; idx++) {
        const fks_obstacle& current_obstacle = obstacles[idx];
        const double effective_resolution = resolution * 0.5;
        const int32_t x_cells = (int32_t)(current_obstacle.extents[0] * 2.0 * (1.0 / effective_resolution));
        const int32_t y_cells = (int32_t)(current_obstacle.extents[1] * 2.0 * (1.0 / effective_resolution));
        const int32_t z_cells = (int32_t)(current_obstacle.extents[2] * 2.0 * (1.0 / effective_resolution));
        for (int32_t xidx = 0; xidx < x_cells; xidx++)
            for (int32_t yidx = 0; yidx < y_cells; yidx++)
                for (int32_t zidx = 0; zidx < z_cells; zidx++) {
                    if (!((xidx == 0) || (yidx == 0) || (zidx == 0) || (xidx == (x_cells - 1)) || (yidx == (y_cells - 1)) || (zidx == (z_cells - 1))))
                        continue;
                    const double x_location = -(current_obstacle.extents[0] - effective_resolution) + (effective_resolution * xidx);
                    const double y_location = -(current_obstacle.extents[1] - effective_resolution) + (effective_resolution * yidx);
                    const double z_location = -(current_obstacle.extents[2] - effective_resolution) + (effective_resolution * zidx);
                    const Vec3 local_cell_location = {x_location, y_location, z_location};
                    const int32_t index[3] = {xidx, yidx, zidx}, count[3] = {x_cells, y_cells, z_cells};
                    std::vector<RawNormal> raw_surface_normals;
                    for (int c = 0; c < 26; c++) {
                        const int cond[3] = {chain[c].cx, chain[c].cy, chain[c].cz};
                        bool match = true;
                        for (int a = 0; a < 3; a++) {
                            if (cond[a] == 0 && index[a] != 0) match = false;
                            if (cond[a] == 1 && index[a] != count[a] - 1) match = false;
                        }
                        if (!match) continue;
                        for (int a = 0; a < 3; a++) {
                            if (cond[a] == 0) raw_surface_normals.push_back({negated(unit[a]), unit[a]});
                            if (cond[a] == 1) raw_surface_normals.push_back({unit[a], negated(unit[a])});
                        }
                        break;  // else-if chain: the first matching case only
                    }
                    // UpdateSurfaceNormalGridCell (envb:162-187)
                    const Vec3 world_location = transform_point(current_obstacle.pose, local_cell_location);
                    int64_t cx, cy, cz;
                    const bool in_bounds = f.index_of(world_location, &cx, &cy, &cz);
                    const float distance = in_bounds ? sdf_at(cx, cy, cz) : std::numeric_limits<float>::infinity();
                    if (distance > -(resolution * 1.5)) {
                        if (!in_bounds) continue;  // Clear / Insert on a location outside the grid return false and do nothing
                        std::vector<StoredNormal>& cell = grid[f.linear(cx, cy, cz)];
                        cell.clear();
                        for (size_t k = 0; k < raw_surface_normals.size(); k++) {
                            const Vec3 real_surface_normal = rotate_vector(current_obstacle.pose, raw_surface_normals[k].normal);
                            const Vec3 real_entry_direction = rotate_vector(current_obstacle.pose, raw_surface_normals[k].entry);
                            cell.push_back(make_stored(real_surface_normal, real_entry_direction));
                        }
                    }
                }
    }
    // flatten to the fks_env_desc layout: non-empty cells in ascending linear index
    env->cell_start.push_back(0);
    for (size_t i = 0; i < ncells; i++) {
        if (grid[i].empty()) continue;
        env->cell_index.push_back((int64_t)i);
        for (const StoredNormal& s : grid[i]) {
            for (int k = 0; k < 4; k++) env->entries.push_back(s.entry[k]);
            for (int k = 0; k < 3; k++) env->entries.push_back(s.normal[k]);
        }
        env->cell_start.push_back((uint32_t)(env->entries.size() / 7));
    }
    fks_env_desc& d = env->desc;
    std::memset(&d, 0, sizeof(d));
    std::memcpy(d.origin, f.origin, sizeof(d.origin));
    std::memcpy(d.inverse_origin, f.inverse_origin, sizeof(d.inverse_origin));
    d.map_resolution = resolution;
    d.sdf_resolution = resolution;
    d.nx = f.nx;
    d.ny = f.ny;
    d.nz = f.nz;
    d.sdf = env->sdf.data();
    d.oob_value = std::numeric_limits<float>::infinity();
    d.n_normal_cells = (int64_t)env->cell_index.size();
    d.normal_cell_index = env->cell_index.data();
    d.normal_cell_start = env->cell_start.data();
    d.normal_entries = env->entries.data();
    return env;
}

}  // extern "C"
