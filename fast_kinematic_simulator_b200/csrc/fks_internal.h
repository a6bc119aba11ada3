// Types shared between the translation units of libfksgpu.so (not part of the C ABI).
#ifndef FKS_INTERNAL_H
#define FKS_INTERNAL_H

#include <cstdint>
#include <string>
#include <vector>

#include "fks_device_types.h"

namespace fks_host {
void set_last_error(const std::string& msg);

// Grid geometry chosen by BuildEnvironment (simulator_environment_builder.cpp:49-160): origin at the
// minimum corner of the obstacle bounding box minus 3.5 cells, identity rotation, ceil(size / res) cells.
struct GridGeometry {
    double origin[12];
    double inv_origin[12];
    double res;
    int64_t nx, ny, nz;
};
// Returns FKS_OK or FKS_ERR_INVALID_ARGUMENT (message set).  Used by the host and the device builder.
int compute_grid_geometry(const fks_obstacle* obstacles, size_t n_obstacles, double resolution, GridGeometry* out);
}  // namespace fks_host

// Device environment (fks_env_create / fks_env_build_device).
struct fks_env {
    int device;
    fksdev::DevEnv dev;
    float* d_sdf;
    unsigned long long* d_keys;   // normal hash, 16-byte slots
    double* d_entries;            // 6 doubles per stored normal
    // CSR view of the normal table in ascending cell order (kept for fks_env_download)
    long long* d_cell_index;      // [n_normal_cells]
    unsigned int* d_cell_start;   // [n_normal_cells + 1]
    unsigned char* d_occupancy;   // 1 byte per cell, only for device-built environments (else null)
    long long n_normal_cells;
    size_t n_entries;
    size_t sdf_bytes;
    size_t l2_window_bytes;
    double build_ms[9];           // device builder phase timings (fks_env_build_timings)
};

// Host environment (fks_build_environment / fks_env_download).
struct fks_built_env {
    fks_env_desc desc;
    std::vector<float> sdf;
    std::vector<uint8_t> occupancy;
    std::vector<int64_t> normal_cell_index;
    std::vector<uint32_t> normal_cell_start;
    std::vector<double> normal_entries;
};

#endif
