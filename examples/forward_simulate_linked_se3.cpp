// C++ users of the drop-in surface for the other two robot kinds of the reference (tnuva_robot_models.hpp:201, :415):
//   1. a 3-joint serial arm (FKS_ROBOT_LINKED) whose last link is driven into a table slab;
//   2. an SE(3) peg (FKS_ROBOT_SE3) pushed against a block face.
// Each scenario runs through fksgpu::MakeGpu{Linked,SE3}Simulator / ForwardSimulateRobots, then again through the several-GPU
// class on every device of the box (records must be identical: Philox noise is keyed by the particle id), once more with
// FKS_NOISE_NONE (all particles of one start must end in the same configuration), and through the SimulatorInterface base.
//   g++ -std=c++11 -Iinclude examples/forward_simulate_linked_se3.cpp -Lfast_kinematic_simulator_b200 -lfksgpu -Wl,-rpath,... -o /tmp/x
// Exit code 0 and "ok"; 3 when no GPU is present (no CPU fallback).
#include <cmath>
#include <cstdio>

#include "fksgpu_glue.hpp"

typedef std::vector<double> Config;

static fks_obstacle Box(double x, double y, double z, double hx, double hy, double hz, uint32_t id) {
    fks_obstacle o;
    const double pose[12] = {1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z};
    for (int i = 0; i < 12; i++) o.pose[i] = pose[i];
    o.extents[0] = hx; o.extents[1] = hy; o.extents[2] = hz;
    o.object_id = id; o._pad = 0;
    return o;
}
static void Identity12(double* t, double x, double y, double z) {
    const double v[12] = {1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z};
    for (int i = 0; i < 12; i++) t[i] = v[i];
}

template <typename Sim>
static int CountDifferent(Sim& a, fksgpu::SimulatorInterface<Config>& b, const std::vector<Config>& starts, const std::vector<Config>& target) {
    a.SetNextParticleId(0);
    const auto ra = a.ForwardSimulateRobots(starts, target, true);
    const auto rb = b.ForwardSimulateRobots(starts, target, true);  // through the interface the planner holds
    int differ = 0;
    for (size_t i = 0; i < ra.size(); i++)
        if (ra[i].result_config != rb[i].result_config || ra[i].n_microsteps != rb[i].n_microsteps || ra[i].flags != rb[i].flags) differ++;
    return differ;
}

static int RunScenario(const char* name, int kind, const fks_env_desc& env, const fks_robot_desc& robot, const Config& start,
                       const Config& target, double contact_fraction_at_least) {
    const fks_solver_params params = fksgpu::GetDefaultSolverParameters();
    fksgpu::GpuSimulatorPtr sim = fksgpu::MakeGpuSimulator(kind, env, robot, params, 25.0, 42, 0);
    const std::vector<Config> starts(96, start), targets(1, target);
    const auto results = sim->ForwardSimulateRobots(starts, targets, true);
    const auto split = fksgpu::SelectByContact(results);
    int bad = 0;
    for (const auto& r : results)
        for (double v : r.result_config)
            if (!std::isfinite(v)) bad++;
    if ((double)split.second.size() < contact_fraction_at_least * (double)results.size()) bad++;
    // every GPU of the box behind one call: the same records
    int n_devices = 1;
    {
        fksgpu::GpuSimulatorPtr single = fksgpu::MakeGpuSimulator(kind, env, robot, params, 25.0, 42, 0);
        fksgpu::SimulatorInterface<Config>& as_interface = *single;
        int count = 0;
        if (fks_device_count(&count) == FKS_OK && count > 1) n_devices = count;
        fksgpu::MultiGpuParticleContactSimulator<Config> multi(std::vector<int32_t>(), n_devices, env, robot, params, 25.0, 42, 0);
        bad += CountDifferent(multi, as_interface, starts, targets);
    }
    // no noise: identical starts end identically, and a second call repeats the first
    sim->SetNoiseMode(FKS_NOISE_NONE);
    const auto quiet = sim->ForwardSimulateRobots(starts, targets, true);
    const auto again = sim->ForwardSimulateRobots(starts, targets, true);
    for (size_t i = 1; i < quiet.size(); i++)
        if (quiet[i].result_config != quiet[0].result_config || again[i].result_config != quiet[0].result_config) bad++;
    const auto stats = sim->GetStatistics();
    std::printf("%s: %zu particles, %zu in contact, %d device(s), collision_resolves %.0f, %s\n", name, results.size(), split.second.size(),
                n_devices, stats.at("collision_resolves"), bad == 0 ? "ok" : "FAILED");
    return bad;
}

int main() {
    try {
        int bad = 0;
        // ---- 1. three-joint arm over a table ------------------------------------------------------------------------
        {
            std::vector<fks_obstacle> obstacles;
            obstacles.push_back(Box(0.0, 0.0, -0.05, 1.5, 1.5, 0.05, 1));   // floor
            obstacles.push_back(Box(0.55, 0.0, 0.30, 0.35, 0.5, 0.05, 2));  // table slab, top at z = 0.35
            fksgpu::Environment env(obstacles, 0.05);
            const int L = 4, J = 3;
            std::vector<double> pts;
            std::vector<int32_t> plink;
            for (int l = 0; l < L; l++)
                for (int k = 0; k < 6; k++) {
                    pts.push_back(0.0); pts.push_back(0.0); pts.push_back(0.05 + 0.05 * k);
                    plink.push_back(l);
                }
            std::vector<fks_joint_desc> joints(J);
            for (int j = 0; j < J; j++) {
                joints[j].parent_link = j; joints[j].child_link = j + 1; joints[j].type = FKS_JOINT_REVOLUTE; joints[j]._pad = 0;
                Identity12(joints[j].transform, 0.0, 0.0, 0.3);
                joints[j].axis[0] = 0.0; joints[j].axis[1] = 1.0; joints[j].axis[2] = 0.0;
                joints[j].lower_limit = -2.5; joints[j].upper_limit = 2.5; joints[j].distance_weight = 1.0;
            }
            std::vector<fks_axis_params> axes(J, fksgpu::AxisParams(3.0, 0.0, 0.0, 0.0, 1.0, 0.1, 0.005));
            std::vector<uint8_t> allowed((size_t)L * L, 1);  // self collisions are not the subject here
            fks_robot_desc robot;
            std::memset(&robot, 0, sizeof(robot));
            robot.kind = FKS_ROBOT_LINKED; robot.n_links = L; robot.n_joints = J; robot.n_dof = J; robot.n_points = (int64_t)plink.size();
            robot.points_xyz = pts.data(); robot.point_link = plink.data(); robot.axes = axes.data();
            Identity12(robot.base_transform, 0.013, 0.007, 0.021);
            robot.joints = joints.data(); robot.allowed_self_collision = allowed.data();
            robot.position_distance_weight = 1.0; robot.rotation_distance_weight = 1.0;
            bad += RunScenario("linked arm", FKS_ROBOT_LINKED, env.Description(), robot, Config{0.2, 0.5, 0.4}, Config{0.6, 1.3, 1.0}, 0.9);
        }
        // ---- 2. SE(3) peg against a block -----------------------------------------------------------------------------
        {
            std::vector<fks_obstacle> obstacles(1, Box(1.0, 0.0, 0.0, 0.5, 1.0, 1.0, 1));  // face at x = 0.5
            fksgpu::Environment env(obstacles, 0.05);
            std::vector<double> pts;
            std::vector<int32_t> plink;
            for (int i = -2; i <= 2; i++)
                for (int j = -2; j <= 2; j++)
                    for (int k = -2; k <= 2; k++) {
                        pts.push_back(0.05 * i); pts.push_back(0.05 * j); pts.push_back(0.05 * k);
                        plink.push_back(0);
                    }
            std::vector<fks_axis_params> axes;
            for (int i = 0; i < 6; i++) axes.push_back(fksgpu::AxisParams(3.0, 0.0, 0.0, 0.0, i < 3 ? 1.0 : 0.5, 0.1, 0.01));
            fks_robot_desc robot;
            std::memset(&robot, 0, sizeof(robot));
            robot.kind = FKS_ROBOT_SE3; robot.n_links = 1; robot.n_joints = 0; robot.n_dof = 6; robot.n_points = (int64_t)plink.size();
            robot.points_xyz = pts.data(); robot.point_link = plink.data(); robot.axes = axes.data();
            robot.position_distance_weight = 1.0; robot.rotation_distance_weight = 1.0;
            Config start(12), target(12);
            Identity12(start.data(), -0.2, 0.011, 0.007);
            Identity12(target.data(), 0.6, 0.011, 0.007);
            bad += RunScenario("SE3 peg", FKS_ROBOT_SE3, env.Description(), robot, start, target, 0.9);
        }
        std::printf("%s\n", bad == 0 ? "ok" : "FAILED");
        return bad == 0 ? 0 : 1;
    } catch (const std::exception& e) {
        std::printf("no device path: %s\n", e.what());
        return 3;
    }
}
