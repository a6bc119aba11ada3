// Host-side environment builder: cuboid obstacles -> occupancy grid -> signed distance field ->
// surface-normal table, in the flat formats the device path reads (include/fksgpu.h fks_env_desc).
//
// Behaviour follows simulator_environment_builder::BuildCompleteEnvironment
// (src/fast_kinematic_simulator/simulator_environment_builder.cpp:470-476):
//   BuildEnvironment        envb.cpp:49-160   (DiscretizeObstacle :21-46)
//   ExtractSignedDistanceField(+inf, {}, true, false)   envb.cpp:473 (sdf_tools, not vendored:
//       value = dist-to-nearest-filled-centre - dist-to-nearest-free-centre, stored as float)
//   BuildSurfaceNormalsGrid envb.cpp:258-468  (UpdateSurfaceNormalGridCell :162-187)
// The distance transform is an exact Euclidean one (separable lower-envelope passes on integer
// squared distances) instead of sdf_tools' bucket propagation; see DESIGN.md "restatement choices".
//
// This is product host code (the workloads and the C++ adapter use it); it is NOT part of oracle/.

#include "fksgpu.h"
#include "../fks_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Iso3 {
    double m[12];
};

inline void iso_apply(const double* T, const double* p, double* out) {
    for (int r = 0; r < 3; r++) out[r] = T[4 * r + 0] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3];
}
inline void iso_rotate(const double* T, const double* v, double* out) {
    for (int r = 0; r < 3; r++) out[r] = T[4 * r + 0] * v[0] + T[4 * r + 1] * v[1] + T[4 * r + 2] * v[2];
}

// EigenHelpers::SafeNormal (spcs.hpp:59,62): v / |v| when |v| > DBL_EPSILON, else v.
inline void safe_normal(const double* v, int n, double* out) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += v[i] * v[i];
    const double nrm = std::sqrt(s);
    if (nrm > std::numeric_limits<double>::epsilon()) {
        for (int i = 0; i < n; i++) out[i] = v[i] / nrm;
    } else {
        for (int i = 0; i < n; i++) out[i] = v[i];
    }
}

struct Grid {
    double origin[12];
    double inv_origin[12];
    double res;
    int64_t nx, ny, nz;
    // VoxelGrid::LocationToGridIndex: grid-frame point * (1/cell) then C-cast (trunc toward zero)
    bool location_to_index(const double* p, int64_t* ix, int64_t* iy, int64_t* iz) const {
        double g[3];
        iso_apply(inv_origin, p, g);
        const double inv = 1.0 / res;
        *ix = (int64_t)(g[0] * inv);
        *iy = (int64_t)(g[1] * inv);
        *iz = (int64_t)(g[2] * inv);
        return *ix >= 0 && *iy >= 0 && *iz >= 0 && *ix < nx && *iy < ny && *iz < nz;
    }
    int64_t lin(int64_t x, int64_t y, int64_t z) const { return (x * ny + y) * nz + z; }
};

const int32_t kInf = 1 << 29;

// 1-D squared distance transform of f along a line (lower envelope of parabolas).
// All inputs are integers (< 2^29), intersections are computed in double: exact, see DESIGN.md.
void dt_1d(const int32_t* f, int32_t* d, int n, int* v, double* z) {
    int k = -1;
    for (int q = 0; q < n; q++) {
        if (f[q] >= kInf) continue;
        if (k < 0) {
            k = 0;
            v[0] = q;
            z[0] = -1e300;
            z[1] = 1e300;
            continue;
        }
        double s;
        while (true) {
            const int p = v[k];
            s = (((double)f[q] + (double)q * q) - ((double)f[p] + (double)p * p)) / (2.0 * q - 2.0 * p);
            if (s <= z[k] && k > 0) {
                k--;
            } else if (s <= z[k]) {  // k == 0: q dominates everywhere
                k = -1;
                break;
            } else {
                break;
            }
        }
        if (k < 0) {
            k = 0;
            v[0] = q;
            z[0] = -1e300;
            z[1] = 1e300;
        } else {
            k++;
            v[k] = q;
            z[k] = s;
            z[k + 1] = 1e300;
        }
    }
    if (k < 0) {
        for (int q = 0; q < n; q++) d[q] = kInf;
        return;
    }
    int j = 0;
    for (int q = 0; q < n; q++) {
        while (z[j + 1] < (double)q) j++;
        const int64_t dq = (int64_t)q - v[j];
        const int64_t val = dq * dq + f[v[j]];
        d[q] = val >= kInf ? kInf : (int32_t)val;
    }
}

// Exact squared Euclidean distance (in cells) from every cell to the nearest cell with seed != 0.
void edt_squared(const std::vector<uint8_t>& seed, int64_t nx, int64_t ny, int64_t nz, std::vector<int32_t>& d2) {
    const int64_t n = nx * ny * nz;
    d2.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) d2[(size_t)i] = seed[(size_t)i] ? 0 : kInf;
    const int maxn = (int)std::max(nx, std::max(ny, nz));
    // z lines (contiguous)
#pragma omp parallel
    {
        std::vector<int32_t> f((size_t)maxn), d((size_t)maxn);
        std::vector<int> v((size_t)maxn + 1);
        std::vector<double> z((size_t)maxn + 2);
#pragma omp for schedule(static)
        for (int64_t xy = 0; xy < nx * ny; xy++) {
            int32_t* line = &d2[(size_t)(xy * nz)];
            for (int64_t k = 0; k < nz; k++) f[(size_t)k] = line[k];
            dt_1d(f.data(), d.data(), (int)nz, v.data(), z.data());
            for (int64_t k = 0; k < nz; k++) line[k] = d[(size_t)k];
        }
#pragma omp for schedule(static)
        for (int64_t xz = 0; xz < nx * nz; xz++) {
            const int64_t x = xz / nz, zz = xz % nz;
            for (int64_t y = 0; y < ny; y++) f[(size_t)y] = d2[(size_t)((x * ny + y) * nz + zz)];
            dt_1d(f.data(), d.data(), (int)ny, v.data(), z.data());
            for (int64_t y = 0; y < ny; y++) d2[(size_t)((x * ny + y) * nz + zz)] = d[(size_t)y];
        }
#pragma omp for schedule(static)
        for (int64_t yz = 0; yz < ny * nz; yz++) {
            const int64_t y = yz / nz, zz = yz % nz;
            for (int64_t x = 0; x < nx; x++) f[(size_t)x] = d2[(size_t)((x * ny + y) * nz + zz)];
            dt_1d(f.data(), d.data(), (int)nx, v.data(), z.data());
            for (int64_t x = 0; x < nx; x++) d2[(size_t)((x * ny + y) * nz + zz)] = d[(size_t)x];
        }
    }
}

struct NormalEntry {
    double e[4];  // entry direction (w = 0)
    double n[3];
};

// StoredSurfaceNormal(normal3, direction3) (spcs.hpp:59-63)
NormalEntry make_entry(const double* normal, const double* direction) {
    NormalEntry out;
    safe_normal(normal, 3, out.n);
    const double d4[4] = {direction[0], direction[1], direction[2], 0.0};
    safe_normal(d4, 4, out.e);
    return out;
}

}  // namespace

namespace {

// sdf_tools SignedDistanceField::GetGradient(x,y,z, enable_edge_gradients=true) (call site envb.cpp:272):
// interior cells: central difference with the subtraction in float and the 1/(2 res) scale in double;
// boundary cells: clamped neighbours, values widened to double first, divided by the index span.
void sdf_gradient(const float* sdf, const Grid& g, int64_t x, int64_t y, int64_t z, double* out) {
    auto at = [&](int64_t a, int64_t b, int64_t c) -> float { return sdf[(size_t)g.lin(a, b, c)]; };
    if (x > 0 && y > 0 && z > 0 && x < g.nx - 1 && y < g.ny - 1 && z < g.nz - 1) {
        const double inv_twice_res = 1.0 / (2.0 * g.res);
        out[0] = (double)(at(x + 1, y, z) - at(x - 1, y, z)) * inv_twice_res;
        out[1] = (double)(at(x, y + 1, z) - at(x, y - 1, z)) * inv_twice_res;
        out[2] = (double)(at(x, y, z + 1) - at(x, y, z - 1)) * inv_twice_res;
        return;
    }
    const int64_t lx = std::max<int64_t>(0, x - 1), hx = std::min<int64_t>(g.nx - 1, x + 1);
    const int64_t ly = std::max<int64_t>(0, y - 1), hy = std::min<int64_t>(g.ny - 1, y + 1);
    const int64_t lz = std::max<int64_t>(0, z - 1), hz = std::min<int64_t>(g.nz - 1, z + 1);
    const double ix = (double)(hx - lx) * g.res, iy = (double)(hy - ly) * g.res, iz = (double)(hz - lz) * g.res;
    out[0] = out[1] = out[2] = 0.0;
    if (ix > 0.0) out[0] = ((double)at(hx, y, z) - (double)at(lx, y, z)) * (1.0 / ix);
    if (iy > 0.0) out[1] = ((double)at(x, hy, z) - (double)at(x, ly, z)) * (1.0 / iy);
    if (iz > 0.0) out[2] = ((double)at(x, y, hz) - (double)at(x, y, lz)) * (1.0 / iz);
}

}  // namespace

// Grid bounds and cell counts of BuildEnvironment (envb.cpp:49-160).
int fks_host::compute_grid_geometry(const fks_obstacle* obstacles, size_t n_obstacles, double resolution,
                                    fks_host::GridGeometry* g) {
    const double res = resolution;
    const double eff = resolution * 0.5;  // envb.cpp:23
    g->res = resolution;
    double x_min = 0, y_min = 0, z_min = 0, x_max = 0, y_max = 0, z_max = 0;
    double x_size = 10.0, y_size = 10.0, z_size = 10.0;
    if (n_obstacles == 0) {
        x_min = y_min = z_min = 0.0;  // default 10x10x10 grid at the origin (envb.cpp:51-65)
    } else {
        bool init = false;
        for (size_t o = 0; o < n_obstacles; o++) {
            const fks_obstacle& ob = obstacles[o];
            const int32_t xc = (int32_t)(ob.extents[0] * 2.0 * (1.0 / eff));
            const int32_t yc = (int32_t)(ob.extents[1] * 2.0 * (1.0 / eff));
            const int32_t zc = (int32_t)(ob.extents[2] * 2.0 * (1.0 / eff));
            // the extreme discretised locations are at the index-range corners; every location is
            // visited in the reference, but min/max over an affine image of a box is attained at
            // its corners, so only those are evaluated here (same comparisons, same values).
            for (int cx = 0; cx < 2; cx++)
                for (int cy = 0; cy < 2; cy++)
                    for (int cz = 0; cz < 2; cz++) {
                        if (xc <= 0 || yc <= 0 || zc <= 0) continue;
                        const int32_t xi = cx ? xc - 1 : 0, yi = cy ? yc - 1 : 0, zi = cz ? zc - 1 : 0;
                        const double loc[3] = {-(ob.extents[0] - (res * 0.5)) + (eff * xi),
                                               -(ob.extents[1] - (res * 0.5)) + (eff * yi),
                                               -(ob.extents[2] - (res * 0.5)) + (eff * zi)};
                        double w[3];
                        iso_apply(ob.pose, loc, w);
                        if (!init) {
                            x_min = x_max = w[0];
                            y_min = y_max = w[1];
                            z_min = z_max = w[2];
                            init = true;
                        } else {
                            x_min = std::min(x_min, w[0]);
                            x_max = std::max(x_max, w[0]);
                            y_min = std::min(y_min, w[1]);
                            y_max = std::max(y_max, w[1]);
                            z_min = std::min(z_min, w[2]);
                            z_max = std::max(z_max, w[2]);
                        }
                    }
        }
        x_min -= (res * 0.5);
        y_min -= (res * 0.5);
        z_min -= (res * 0.5);
        x_min -= (res * 3.0);
        y_min -= (res * 3.0);
        z_min -= (res * 3.0);
        x_max += (res * 3.0);
        y_max += (res * 3.0);
        z_max += (res * 3.0);
        x_size = x_max - x_min;
        y_size = y_max - y_min;
        z_size = z_max - z_min;
    }
    const double ident[12] = {1, 0, 0, x_min, 0, 1, 0, y_min, 0, 0, 1, z_min};
    const double ident_inv[12] = {1, 0, 0, -x_min, 0, 1, 0, -y_min, 0, 0, 1, -z_min};
    std::memcpy(g->origin, ident, sizeof(ident));
    std::memcpy(g->inv_origin, ident_inv, sizeof(ident_inv));
    g->nx = (int64_t)std::ceil(x_size / res);  // VoxelGrid ctor: ceil(size / cell_size)
    g->ny = (int64_t)std::ceil(y_size / res);
    g->nz = (int64_t)std::ceil(z_size / res);
    if (g->nx <= 0 || g->ny <= 0 || g->nz <= 0 || g->nx > 4096 || g->ny > 4096 || g->nz > 4096) {
        fks_host::set_last_error("fks_build_environment: grid dimensions out of range");
        return FKS_ERR_INVALID_ARGUMENT;
    }
    return FKS_OK;
}

extern "C" int fks_build_environment(const fks_obstacle* obstacles, size_t n_obstacles, double resolution,
                                     fks_built_env** out) {
    if (!out || resolution <= 0.0 || (n_obstacles > 0 && !obstacles)) {
        fks_host::set_last_error("fks_build_environment: invalid argument");
        return FKS_ERR_INVALID_ARGUMENT;
    }
    for (size_t i = 0; i < n_obstacles; i++) {
        if (obstacles[i].object_id == 0) {  // envb.hpp:35,41 assert(in_object_id > 0)
            fks_host::set_last_error("fks_build_environment: obstacle object_id must be > 0");
            return FKS_ERR_INVALID_ARGUMENT;
        }
    }
    fks_built_env* env = new (std::nothrow) fks_built_env();
    if (!env) return FKS_ERR_OUT_OF_MEMORY;
    Grid g;
    g.res = resolution;
    const double res = resolution;
    const double eff = resolution * 0.5;  // envb.cpp:23

    // ---- BuildEnvironment (envb.cpp:49-160) ---------------------------------------------------
    {
        fks_host::GridGeometry gg;
        const int rc = fks_host::compute_grid_geometry(obstacles, n_obstacles, resolution, &gg);
        if (rc != FKS_OK) {
            delete env;
            return rc;
        }
        std::memcpy(g.origin, gg.origin, sizeof(g.origin));
        std::memcpy(g.inv_origin, gg.inv_origin, sizeof(g.inv_origin));
        g.nx = gg.nx;
        g.ny = gg.ny;
        g.nz = gg.nz;
    }
    const int64_t ncells = g.nx * g.ny * g.nz;
    env->occupancy.assign((size_t)ncells, 0);
    for (size_t o = 0; o < n_obstacles; o++) {
        const fks_obstacle& ob = obstacles[o];
        const int32_t xc = (int32_t)(ob.extents[0] * 2.0 * (1.0 / eff));
        const int32_t yc = (int32_t)(ob.extents[1] * 2.0 * (1.0 / eff));
        const int32_t zc = (int32_t)(ob.extents[2] * 2.0 * (1.0 / eff));
        for (int32_t xi = 0; xi < xc; xi++)
            for (int32_t yi = 0; yi < yc; yi++)
                for (int32_t zi = 0; zi < zc; zi++) {
                    const double loc[3] = {-(ob.extents[0] - (res * 0.5)) + (eff * xi),
                                           -(ob.extents[1] - (res * 0.5)) + (eff * yi),
                                           -(ob.extents[2] - (res * 0.5)) + (eff * zi)};
                    double w[3];
                    iso_apply(ob.pose, loc, w);
                    int64_t ix, iy, iz;
                    if (g.location_to_index(w, &ix, &iy, &iz)) env->occupancy[(size_t)g.lin(ix, iy, iz)] = 1;
                }
    }

    // ---- ExtractSignedDistanceField(+inf, {}, true, false) (envb.cpp:473) -----------------------
    {
        std::vector<int32_t> d_filled, d_free;
        edt_squared(env->occupancy, g.nx, g.ny, g.nz, d_filled);
        std::vector<uint8_t> freec((size_t)ncells);
        for (int64_t i = 0; i < ncells; i++) freec[(size_t)i] = env->occupancy[(size_t)i] ? 0 : 1;
        edt_squared(freec, g.nx, g.ny, g.nz, d_free);
        env->sdf.resize((size_t)ncells);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < ncells; i++) {
            const double d1 = std::sqrt((double)d_filled[(size_t)i]) * res;
            const double d2 = std::sqrt((double)d_free[(size_t)i]) * res;
            env->sdf[(size_t)i] = (float)(d1 - d2);
        }
    }
    const float* sdf = env->sdf.data();

    // ---- BuildSurfaceNormalsGrid (envb.cpp:258-468) --------------------------------------------
    // pass 2 first into a map (it clears and rewrites whole cells, later writers win), then one
    // ordered sweep emits either the pass-2 cell or the pass-1 gradient entry.
    std::unordered_map<int64_t, std::vector<NormalEntry>> exact_cells;
    for (size_t o = 0; o < n_obstacles; o++) {
        const fks_obstacle& ob = obstacles[o];
        const int32_t xc = (int32_t)(ob.extents[0] * 2.0 * (1.0 / eff));
        const int32_t yc = (int32_t)(ob.extents[1] * 2.0 * (1.0 / eff));
        const int32_t zc = (int32_t)(ob.extents[2] * 2.0 * (1.0 / eff));
        for (int32_t xi = 0; xi < xc; xi++)
            for (int32_t yi = 0; yi < yc; yi++)
                for (int32_t zi = 0; zi < zc; zi++) {
                    const bool x0 = xi == 0, x1 = xi == xc - 1, y0 = yi == 0, y1 = yi == yc - 1, z0 = zi == 0,
                               z1 = zi == zc - 1;
                    if (!(x0 || x1 || y0 || y1 || z0 || z1)) continue;
                    // The 26-way if/else chain of envb.cpp:302-461 reduces to: test each axis with the
                    // low face first (x0 before x1 etc.); corners get x,y,z normals, edges the two
                    // axes involved in x,y,z order, faces one.  The chain tests index==0 before
                    // index==n-1 on every axis, so for an axis with a single cell the low face wins.
                    double normals[3][3];
                    int nn = 0;
                    const int sx = x0 ? -1 : (x1 ? 1 : 0);
                    const int sy = y0 ? -1 : (y1 ? 1 : 0);
                    const int sz = z0 ? -1 : (z1 ? 1 : 0);
                    if (sx) { normals[nn][0] = sx; normals[nn][1] = 0; normals[nn][2] = 0; nn++; }
                    if (sy) { normals[nn][0] = 0; normals[nn][1] = sy; normals[nn][2] = 0; nn++; }
                    if (sz) { normals[nn][0] = 0; normals[nn][1] = 0; normals[nn][2] = sz; nn++; }
                    const double loc[3] = {-(ob.extents[0] - eff) + (eff * xi), -(ob.extents[1] - eff) + (eff * yi),
                                           -(ob.extents[2] - eff) + (eff * zi)};
                    double w[3];
                    iso_apply(ob.pose, loc, w);
                    int64_t ix, iy, iz;
                    const bool inb = g.location_to_index(w, &ix, &iy, &iz);
                    // UpdateSurfaceNormalGridCell (envb.cpp:162-187)
                    const float distance = inb ? sdf[(size_t)g.lin(ix, iy, iz)] : std::numeric_limits<float>::infinity();
                    if ((double)distance > -(res * 1.5)) {
                        if (!inb) continue;  // Clear/Insert on an out-of-bounds location are no-ops
                        std::vector<NormalEntry>& cell = exact_cells[g.lin(ix, iy, iz)];
                        cell.clear();
                        for (int k = 0; k < nn; k++) {
                            double rn[3], re[3];
                            const double raw_e[3] = {-normals[k][0], -normals[k][1], -normals[k][2]};
                            iso_rotate(ob.pose, normals[k], rn);
                            iso_rotate(ob.pose, raw_e, re);
                            cell.push_back(make_entry(rn, re));
                        }
                    }
                }
    }
    env->normal_cell_start.push_back(0);
    for (int64_t x = 0; x < g.nx; x++)
        for (int64_t y = 0; y < g.ny; y++)
            for (int64_t z = 0; z < g.nz; z++) {
                const int64_t li = g.lin(x, y, z);
                auto it = exact_cells.find(li);
                if (it != exact_cells.end()) {
                    if (it->second.empty()) continue;
                    env->normal_cell_index.push_back(li);
                    for (const NormalEntry& e : it->second) {
                        for (int k = 0; k < 4; k++) env->normal_entries.push_back(e.e[k]);
                        for (int k = 0; k < 3; k++) env->normal_entries.push_back(e.n[k]);
                    }
                    env->normal_cell_start.push_back((uint32_t)(env->normal_entries.size() / 7));
                } else if (sdf[(size_t)li] < 0.0f) {  // envb.cpp:269-274
                    double grad[3];
                    sdf_gradient(sdf, g, x, y, z, grad);
                    const double zero[3] = {0.0, 0.0, 0.0};
                    const NormalEntry e = make_entry(grad, zero);
                    env->normal_cell_index.push_back(li);
                    for (int k = 0; k < 4; k++) env->normal_entries.push_back(e.e[k]);
                    for (int k = 0; k < 3; k++) env->normal_entries.push_back(e.n[k]);
                    env->normal_cell_start.push_back((uint32_t)(env->normal_entries.size() / 7));
                }
            }

    fks_env_desc& d = env->desc;
    std::memset(&d, 0, sizeof(d));
    std::memcpy(d.origin, g.origin, sizeof(d.origin));
    std::memcpy(d.inverse_origin, g.inv_origin, sizeof(d.inverse_origin));
    d.map_resolution = res;
    d.sdf_resolution = res;
    d.nx = g.nx;
    d.ny = g.ny;
    d.nz = g.nz;
    d.sdf = env->sdf.data();
    d.oob_value = std::numeric_limits<float>::infinity();
    d.n_normal_cells = (int64_t)env->normal_cell_index.size();
    d.normal_cell_index = env->normal_cell_index.data();
    d.normal_cell_start = env->normal_cell_start.data();
    d.normal_entries = env->normal_entries.data();
    *out = env;
    return FKS_OK;
}

extern "C" const fks_env_desc* fks_built_env_desc(const fks_built_env* env) { return env ? &env->desc : nullptr; }
extern "C" const uint8_t* fks_built_env_occupancy(const fks_built_env* env) {
    return (env && !env->occupancy.empty()) ? env->occupancy.data() : nullptr;
}
extern "C" void fks_built_env_destroy(fks_built_env* env) { delete env; }
