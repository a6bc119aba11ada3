// Minimal C++ user of the drop-in surface: builds a one-wall environment, an SE(2) point robot, and pushes
// 64 particles into the wall through fksgpu::MakeGpuSE2Simulator / ForwardSimulateRobots.
//   g++ -std=c++11 -Iinclude examples/forward_simulate_se2.cpp -Lfast_kinematic_simulator_b200 -lfksgpu -Wl,-rpath,... -o /tmp/se2
// Exit code 0 and "ok" when every particle ends within one voxel of the wall face; 3 when no GPU is present
// (the library has no CPU fallback and says so).
#include <cmath>
#include <cstdio>

#include "fksgpu_simulator.hpp"

int main() {
    fks_obstacle wall;
    const double pose[12] = {1, 0, 0, 1.25, 0, 1, 0, 0.0, 0, 0, 1, 0.05};
    for (int i = 0; i < 12; i++) wall.pose[i] = pose[i];
    wall.extents[0] = 0.25; wall.extents[1] = 2.0; wall.extents[2] = 0.5;
    wall.object_id = 1; wall._pad = 0;
    fksgpu::Environment env(std::vector<fks_obstacle>(1, wall), 0.125);

    const double point[3] = {0.0, 0.0, 0.0};
    const int32_t link0 = 0;
    fks_axis_params axes[3];
    for (int i = 0; i < 3; i++) {
        axes[i].kp = 1.0; axes[i].ki = 0.0; axes[i].kd = 0.0; axes[i].integral_clamp = 0.0;
        axes[i].velocity_limit = (i < 2) ? 1.0 : 0.5;
        axes[i].proportional_noise = 0.1; axes[i].minimum_noise = 0.01; axes[i].noise_sigma = 0.5;
    }
    fks_robot_desc robot;
    std::memset(&robot, 0, sizeof(robot));
    robot.kind = FKS_ROBOT_SE2; robot.n_links = 1; robot.n_joints = 0; robot.n_dof = 3; robot.n_points = 1;
    robot.points_xyz = point; robot.point_link = &link0; robot.axes = axes;
    robot.position_distance_weight = 1.0; robot.rotation_distance_weight = 1.0;

    try {
        fksgpu::GpuSimulatorPtr sim = fksgpu::MakeGpuSE2Simulator(env.Description(), robot, fksgpu::GetDefaultSolverParameters(), 25.0, 42, 0);
        std::vector<std::vector<double>> starts(64, std::vector<double>{0.6, 0.0, 0.0});
        std::vector<std::vector<double>> target(1, std::vector<double>{1.4, 0.0, 0.0});
        auto results = sim->ForwardSimulateRobots(starts, target, true);
        int bad = 0;
        for (const auto& r : results)
            if (!r.did_contact || r.result_config[0] < 1.0 - 0.125 || r.result_config[0] > 1.0 + 0.02) bad++;
        auto stats = sim->GetStatistics();
        // the same environment built on the device (fks_env_build_device): identical end states (Philox noise is keyed by
        // particle id, and a fresh simulator starts at id 0 again)
        std::shared_ptr<fksgpu::DeviceEnvironment> denv(new fksgpu::DeviceEnvironment(std::vector<fks_obstacle>(1, wall), 0.125));
        fksgpu::GpuSimulatorPtr sim2 = fksgpu::MakeGpuSimulator(FKS_ROBOT_SE2, denv, robot, fksgpu::GetDefaultSolverParameters(), 25.0, 42, 0);
        auto results2 = sim2->ForwardSimulateRobots(starts, target, true);
        int differ = 0;
        for (size_t i = 0; i < results.size(); i++)
            if (results[i].result_config != results2[i].result_config || results[i].n_microsteps != results2[i].n_microsteps) differ++;
        // one particle with the step trace of the reference (ForwardSimulateRobot, enable_tracing = true)
        fksgpu::ForwardSimulationStepTrace<std::vector<double>> trace;
        auto traced = sim2->ForwardSimulateRobot(starts[0], target[0], true, trace, true);
        size_t micro = 0, configs = 0;
        for (const auto& rs : trace.resolver_steps) {
            micro += rs.contact_resolver_steps.size();
            for (const auto& cs : rs.contact_resolver_steps) configs += cs.contact_resolution_steps.size();
        }
        if (trace.resolver_steps.size() != traced.n_steps || micro != traced.n_microsteps ||
            configs < (size_t)traced.n_microsteps + traced.n_resolver_iterations)
            differ++;
        std::printf("trace: %zu controller steps, %zu microsteps, %zu configurations\n", trace.resolver_steps.size(), micro, configs);
        std::printf("%zu particles, %d outside the expected band, collision_resolves %.0f, %d differ in the device-built environment (built in %.3f ms)\n%s\n",
                    results.size(), bad, stats["collision_resolves"], differ, denv->BuildTimingsMs()[0], (bad == 0 && differ == 0) ? "ok" : "FAILED");
        return (bad == 0 && differ == 0) ? 0 : 1;
    } catch (const std::exception& e) {
        std::printf("no device path: %s\n", e.what());
        return 3;
    }
}
