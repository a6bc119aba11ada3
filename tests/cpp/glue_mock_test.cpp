// Compile-and-run check of the always-on templates of include/fksgpu_glue.hpp against MOCK types that have exactly the members
// the reference calls on its own types (sdf_tools grids, PointSphereGeometry, Eigen vectors / transforms): no GPU, no Eigen.
#include <cstdio>
#include <memory>

#include "fksgpu_glue.hpp"

struct Vec {  // Eigen::Vector3d / Vector4d: operator()(i)
    double v[4];
    double operator()(int i) const { return v[i]; }
};
struct Iso {  // Eigen::Isometry3d: operator()(row, col)
    double m[4][4];
    double operator()(int r, int c) const { return m[r][c]; }
};
struct Stored {  // SurfaceNormalGrid::StoredSurfaceNormal (spcs.hpp:48-83)
    Vec dir, n;
    const Vec& EntryDirection4d() const { return dir; }
    const Vec& Normal() const { return n; }
};
struct Map {  // sdf_tools::TaggedObjectCollisionMapGrid metadata
    Iso o, inv;
    double GetResolution() const { return 0.25; }
    const Iso& GetOriginTransform() const { return o; }
    const Iso& GetInverseOriginTransform() const { return inv; }
};
struct Sdf {  // sdf_tools::SignedDistanceField
    int64_t GetNumXCells() const { return 2; }
    int64_t GetNumYCells() const { return 3; }
    int64_t GetNumZCells() const { return 4; }
    double GetResolution() const { return 0.25; }
    std::pair<float, bool> GetImmutable(int64_t x, int64_t y, int64_t z) const { return std::make_pair((float)(100 * x + 10 * y + z), true); }
};
struct PointGeometry {  // simple_robot_models::PointSphereGeometry: Geometry() -> shared_ptr to the points
    std::shared_ptr<std::vector<Vec>> pts;
    std::shared_ptr<std::vector<Vec>> Geometry() const { return pts; }
};
struct RigidConfig {  // SE2_ROBOT_CONFIG / SE3_ROBOT_CONFIG fields read by the glue
    double kp = 1, ki = 2, kd = 3, integral_clamp = 4, velocity_limit = 5, max_actuator_proportional_noise = 6, max_actuator_minimum_noise = 7;
    double r_kp = 11, r_ki = 12, r_kd = 13, r_integral_clamp = 14, r_velocity_limit = 15, r_max_actuator_proportional_noise = 16,
           r_max_actuator_minimum_noise = 17;
};

int main() {
    int bad = 0;
    Map map;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { map.o.m[r][c] = 10 * r + c; map.inv.m[r][c] = -(10 * r + c); }
    Sdf sdf;
    std::vector<Stored> empty, two(2);
    two[0].dir = Vec{{1, 0, 0, 0}}; two[0].n = Vec{{0, 0, 1, 0}};
    two[1].dir = Vec{{0, 1, 0, 0}}; two[1].n = Vec{{0, 1, 0, 0}};
    auto normals_of = [&](int64_t x, int64_t y, int64_t z) -> const std::vector<Stored>& { return (x == 1 && y == 2 && z == 3) || (x == 0 && y == 1 && z == 0) ? two : empty; };
    fksgpu::FlatEnvironment env;
    fksgpu::FlattenEnvironment(map, sdf, normals_of, 1e30f, env);
    const fks_env_desc& d = env.desc;
    if (d.nx != 2 || d.ny != 3 || d.nz != 4 || d.origin[3] != 3.0 || d.origin[4] != 10.0 || d.inverse_origin[11] != -23.0) bad++;
    if (d.sdf[(1 * 3 + 2) * 4 + 3] != 123.0f || d.sdf[0] != 0.0f) bad++;
    if (d.n_normal_cells != 2 || d.normal_cell_index[0] != 4 || d.normal_cell_index[1] != 23 || d.normal_cell_start[2] != 4) bad++;
    if (d.normal_entries[7 * 1 + 1] != 1.0 || d.normal_entries[7 * 0 + 6] != 1.0) bad++;

    std::vector<std::pair<std::string, PointGeometry>> links(2);
    links[0].first = "a"; links[0].second.pts.reset(new std::vector<Vec>(2, Vec{{1, 2, 3, 1}}));
    links[1].first = "b"; links[1].second.pts.reset(new std::vector<Vec>(1, Vec{{4, 5, 6, 1}}));
    const fksgpu::FlatGeometry g = fksgpu::FlattenLinkGeometries(links);
    if (g.point_link != std::vector<int32_t>{0, 0, 1} || g.points_xyz.size() != 9 || g.points_xyz[8] != 6.0 || g.link_names[1] != "b") bad++;

    const std::vector<fks_axis_params> axes = fksgpu::AxesOfRigidBodyConfig(RigidConfig(), 2, 1);
    if (axes.size() != 3 || axes[1].kd != 3 || axes[2].velocity_limit != 15 || axes[2].minimum_noise != 17 || axes[0].noise_sigma != 0.5) bad++;

    std::vector<fksgpu::SimulationResult<std::vector<double>>> results(3);
    results[0].did_contact = false; results[1].did_contact = true; results[2].did_contact = false;
    const auto split = fksgpu::SelectByContact(results);
    if (split.first != std::vector<size_t>{0, 2} || split.second != std::vector<size_t>{1}) bad++;
    std::printf("%s\n", bad == 0 ? "ok" : "FAILED");
    return bad;
}
