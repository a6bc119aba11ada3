// Glue between the reference's own C++ types and the flat layouts of the C ABI (include/fksgpu.h), for the maintainer who
// wires the B200 path into calderpg/fast_kinematic_simulator.
//
// Three layers, each behind its own switch because none of the reference's dependencies exists in this repository's image:
//   (always)                       templates over "anything with the reference's call-site API": Transform12, AxisParams,
//                                  FlattenLinkGeometries, FlattenEnvironment, SelectByContact.  tests/cpp/glue_mock_test.cpp
//                                  instantiates them with mock types that have exactly the members the reference calls.
//   FKSGPU_WITH_EIGEN              ConfigTraits for the reference's configuration types: Eigen::Matrix<double, 3, 1>
//                                  (SE2, tnuva_robot_models.hpp:26) and Eigen::Isometry3d (SE3, :201).
//   FKSGPU_WITH_REFERENCE_HEADERS  (needs sdf_tools, arc_utilities, uncertainty_planning_core, ROS message headers and the
//                                  reference's include/) FlattenReferenceEnvironment for the three grids the factories take
//                                  (fast_kinematic_simulator.hpp:18-22) and GpuBatchSimulator, the subclass of
//                                  SimpleParticleContactSimulator that overrides ForwardSimulateRobots (spcs.hpp:788).
// Nothing here is on the hot path: it runs once per environment / robot, and per batch only to copy configurations.
#ifndef FKSGPU_GLUE_HPP
#define FKSGPU_GLUE_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "fksgpu_simulator.hpp"

namespace fksgpu {

// Eigen::Isometry3d (or anything with operator()(row, col)) -> row-major 3x4 [R | t]; never through quaternions, so the
// device sees the reference's bits
template <typename Transform>
inline void Transform12(const Transform& T, double out[12]) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) out[4 * r + c] = T(r, c);
}

// One actuated axis from a *_ROBOT_CONFIG block (tnuva_robot_models.hpp:43-62, :227-246, :420-431): the PID gains go to
// SimplePIDController (:118-120), the limits and noise bounds to TruncatedNormalUncertainVelocityActuator with
// percent_variance 0.5 (:128-130, :469).  Sensor noise is not on this path (the simulator reads the true configuration).
inline fks_axis_params AxisParams(double kp, double ki, double kd, double integral_clamp, double velocity_limit,
                                  double max_actuator_proportional_noise, double max_actuator_minimum_noise) {
    fks_axis_params a;
    a.kp = kp;
    a.ki = ki;
    a.kd = kd;
    a.integral_clamp = integral_clamp;
    a.velocity_limit = velocity_limit;
    a.proportional_noise = max_actuator_proportional_noise;
    a.minimum_noise = max_actuator_minimum_noise;
    a.noise_sigma = 0.5;
    return a;
}
// SE2: x, y share the linear block, zr takes the r_ block (tnuva.hpp:118-130); SE3: x, y, z / xr, yr, zr (:302-323)
template <typename RobotConfig>
inline std::vector<fks_axis_params> AxesOfRigidBodyConfig(const RobotConfig& c, int linear_axes, int rotary_axes) {
    std::vector<fks_axis_params> axes;
    for (int i = 0; i < linear_axes; i++)
        axes.push_back(AxisParams(c.kp, c.ki, c.kd, c.integral_clamp, c.velocity_limit, c.max_actuator_proportional_noise, c.max_actuator_minimum_noise));
    for (int i = 0; i < rotary_axes; i++)
        axes.push_back(AxisParams(c.r_kp, c.r_ki, c.r_kd, c.r_integral_clamp, c.r_velocity_limit, c.r_max_actuator_proportional_noise,
                                  c.r_max_actuator_minimum_noise));
    return axes;
}

// robot->GetLinkGeometries() (std::vector<std::pair<std::string, PointSphereGeometry>>, walked at spcs.hpp:925-936) ->
// link-major point list of fks_robot_desc.  `Geometry()` returns a pointer to a vector of 4-vectors (w = 1).
struct FlatGeometry {
    std::vector<double> points_xyz;
    std::vector<int32_t> point_link;
    std::vector<std::string> link_names;
};
template <typename LinkGeometries>
inline FlatGeometry FlattenLinkGeometries(const LinkGeometries& robot_link_geometries) {
    FlatGeometry g;
    for (size_t link_idx = 0; link_idx < robot_link_geometries.size(); link_idx++) {
        g.link_names.push_back(robot_link_geometries[link_idx].first);
        const auto& link_points = *(robot_link_geometries[link_idx].second.Geometry());
        for (size_t point_idx = 0; point_idx < link_points.size(); point_idx++) {
            for (int k = 0; k < 3; k++) g.points_xyz.push_back(link_points[point_idx](k));
            g.point_link.push_back((int32_t)link_idx);
        }
    }
    return g;
}

// The three grids the reference simulator copies at construction (spcs.hpp:420) -> fks_env_desc + the storage it points to.
//   collision_map : sdf_tools::TaggedObjectCollisionMapGrid   GetResolution / GetOriginTransform / GetInverseOriginTransform
//   sdf           : sdf_tools::SignedDistanceField             GetNumX/Y/ZCells, GetResolution, GetImmutable(x, y, z).first
//   normals_of    : callable (x, y, z) -> const reference to the cell's std::vector<StoredSurfaceNormal>
//                   (EntryDirection4d() / Normal(), spcs.hpp:48-83; both already SafeNormal'd by the constructor)
struct FlatEnvironment {
    fks_env_desc desc;
    std::vector<float> sdf;
    std::vector<int64_t> normal_cell_index;
    std::vector<uint32_t> normal_cell_start;
    std::vector<double> normal_entries;
    FlatEnvironment() {}
    FlatEnvironment(const FlatEnvironment&) = delete;  // desc points into the vectors
    FlatEnvironment& operator=(const FlatEnvironment&) = delete;
};
template <typename CollisionMap, typename SignedDistanceField, typename NormalsOf>
inline void FlattenEnvironment(const CollisionMap& collision_map, const SignedDistanceField& sdf, const NormalsOf& normals_of,
                               float oob_value, FlatEnvironment& out) {
    fks_env_desc& d = out.desc;
    Transform12(collision_map.GetOriginTransform(), d.origin);
    Transform12(collision_map.GetInverseOriginTransform(), d.inverse_origin);
    d.map_resolution = collision_map.GetResolution();  // microstep sizing, self-collision cells (spcs.hpp:524-527, 1219, 1560)
    d.sdf_resolution = sdf.GetResolution();            // collision thresholds (spcs.hpp:923, 957)
    d.nx = sdf.GetNumXCells();
    d.ny = sdf.GetNumYCells();
    d.nz = sdf.GetNumZCells();
    d.oob_value = oob_value;
    d._pad = 0;
    out.sdf.resize((size_t)(d.nx * d.ny * d.nz));
    out.normal_cell_index.clear();
    out.normal_cell_start.assign(1, 0u);
    out.normal_entries.clear();
    for (int64_t x = 0; x < d.nx; x++)
        for (int64_t y = 0; y < d.ny; y++)
            for (int64_t z = 0; z < d.nz; z++) {
                const int64_t linear = (x * d.ny + y) * d.nz + z;
                out.sdf[(size_t)linear] = sdf.GetImmutable(x, y, z).first;
                const auto& stored = normals_of(x, y, z);
                if (stored.size() == 0) continue;
                out.normal_cell_index.push_back(linear);  // ascending by construction
                for (size_t i = 0; i < stored.size(); i++) {
                    for (int k = 0; k < 4; k++) out.normal_entries.push_back(stored[i].EntryDirection4d()(k));
                    for (int k = 0; k < 3; k++) out.normal_entries.push_back(stored[i].Normal()(k));
                }
                out.normal_cell_start.push_back((uint32_t)(out.normal_entries.size() / 7));
            }
    d.sdf = out.sdf.data();
    d.n_normal_cells = (int64_t)out.normal_cell_index.size();
    d.normal_cell_index = out.normal_cell_index.data();
    d.normal_cell_start = out.normal_cell_start.data();
    d.normal_entries = out.normal_entries.data();
}

// The planner's first use of a batch (uncertainty_planning_core.cpp:97-99): split the particles by did_contact
template <typename Result>
inline std::pair<std::vector<size_t>, std::vector<size_t>> SelectByContact(const std::vector<Result>& results) {
    std::pair<std::vector<size_t>, std::vector<size_t>> out;  // (no contact, contact)
    for (size_t i = 0; i < results.size(); i++) (results[i].did_contact ? out.second : out.first).push_back(i);
    return out;
}

}  // namespace fksgpu

#ifdef FKSGPU_WITH_EIGEN
#include <Eigen/Geometry>
namespace fksgpu {
// SE2 configuration (x, y, theta), tnuva_robot_models.hpp:26
template <>
struct ConfigTraits<Eigen::Matrix<double, 3, 1>> {
    static void Flatten(const Eigen::Matrix<double, 3, 1>& c, double* out, int stride) {
        if (stride != 3) throw std::invalid_argument("fksgpu: an SE2 configuration has 3 values");
        for (int i = 0; i < 3; i++) out[i] = c(i);
    }
    static Eigen::Matrix<double, 3, 1> Unflatten(const double* in, int) { return Eigen::Matrix<double, 3, 1>(in[0], in[1], in[2]); }
};
// SE3 configuration, tnuva_robot_models.hpp:201
template <>
struct ConfigTraits<Eigen::Isometry3d> {
    static void Flatten(const Eigen::Isometry3d& c, double* out, int stride) {
        if (stride != 12) throw std::invalid_argument("fksgpu: an SE3 configuration has 12 values");
        Transform12(c, out);
    }
    static Eigen::Isometry3d Unflatten(const double* in, int) {
        Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) T(r, c) = in[4 * r + c];
        return T;
    }
};
}  // namespace fksgpu
#endif  // FKSGPU_WITH_EIGEN

#ifdef FKSGPU_WITH_REFERENCE_HEADERS
#include <fast_kinematic_simulator/simple_particle_contact_simulator.hpp>
namespace fksgpu {
// SurfaceNormalGrid keeps its cells protected (spcs.hpp:46, :134): read them through a derived view
class SurfaceNormalCells : public simple_particle_contact_simulator::SurfaceNormalGrid {
public:
    typedef std::vector<StoredSurfaceNormal> Cell;
    const Cell& operator()(int64_t x, int64_t y, int64_t z) const { return surface_normal_grid_.GetImmutable(x, y, z).first; }
};
inline void FlattenReferenceEnvironment(const sdf_tools::TaggedObjectCollisionMapGrid& environment,
                                        const sdf_tools::SignedDistanceField& environment_sdf,
                                        const simple_particle_contact_simulator::SurfaceNormalGrid& surface_normals_grid,
                                        FlatEnvironment& out) {
    // out-of-bounds reads of the SDF return its default value: +inf in BuildCompleteEnvironment (envb.cpp:473)
    FlattenEnvironment(environment, environment_sdf, static_cast<const SurfaceNormalCells&>(surface_normals_grid),
                       std::numeric_limits<float>::infinity(), out);
}

// The GPU sibling the factories return: SimpleParticleContactSimulator with ONE override.
template <typename Robot, typename Config, typename RNG, typename Alloc>
class GpuBatchSimulator : public simple_particle_contact_simulator::SimpleParticleContactSimulator<Robot, Config, RNG, Alloc> {
    typedef simple_particle_contact_simulator::SimpleParticleContactSimulator<Robot, Config, RNG, Alloc> Base;
    typedef simple_simulator_interface::SimulationResult<Config> ReferenceResult;
    std::shared_ptr<GpuParticleContactSimulator<Config>> gpu_;  // set by the factory (environment + robot description + seed)

public:
    using Base::Base;
    void SetGpuSimulator(const std::shared_ptr<GpuParticleContactSimulator<Config>>& gpu) { gpu_ = gpu; }
    std::vector<ReferenceResult> ForwardSimulateRobots(const std::shared_ptr<typename Base::BaseRobotType>& immutable_robot,
                                                       const std::vector<Config, Alloc>& start_positions,
                                                       const std::vector<Config, Alloc>& target_positions, const bool allow_contacts,
                                                       const std::function<void(const visualization_msgs::MarkerArray&)>& display_fn) override {
        // markers are only produced at debug_level >= 2 (spcs.hpp:1725-1736): that case stays on the CPU path
        if (!gpu_ || this->GetDebugLevel() >= 2) return Base::ForwardSimulateRobots(immutable_robot, start_positions, target_positions, allow_contacts, display_fn);
        const std::vector<Config> starts(start_positions.begin(), start_positions.end()), targets(target_positions.begin(), target_positions.end());
        const auto results = gpu_->ForwardSimulateRobots(starts, targets, allow_contacts);
        std::vector<ReferenceResult> out;
        out.reserve(results.size());
        for (size_t i = 0; i < results.size(); i++)
            out.emplace_back(results[i].result_config, results[i].actual_target, results[i].did_contact, results[i].outcome_is_nominal);  // spcs.hpp:918
        return out;
    }
};
}  // namespace fksgpu
#endif  // FKSGPU_WITH_REFERENCE_HEADERS

#endif  // FKSGPU_GLUE_HPP
