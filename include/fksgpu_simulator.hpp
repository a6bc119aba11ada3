// C++ host adapter above the C ABI (include/fksgpu.h): the GPU sibling of
// simple_particle_contact_simulator::SimpleParticleContactSimulator for the batch calls.
//
// It mirrors the reference's interface for this path -- same names, argument meaning and error behaviour:
//   ForwardSimulateRobots / ReverseSimulateRobots   simple_particle_contact_simulator.hpp:788-804 / :806-822
//   GetStatistics / ResetStatistics                 :488-512
//   MakeGpu{SE2,SE3,Linked}Simulator                fast_kinematic_simulator.hpp:18-22 (same argument order)
//   GetDefaultSolverParameters                      fast_kinematic_simulator.hpp:13-16
// but carries no Eigen / ROS / arc_utilities types: configurations are the flat layouts of the C ABI
// (SE2: x, y, theta; SE3: row-major 3x4 [R|t]; linked: one value per active joint).  INTEGRATION.md shows the
// ~30-line glue that converts the reference's own types (Eigen::Matrix<double,3,1>, Eigen::Isometry3d,
// std::vector<SimpleJointModel>) and derives this class from simple_simulator_interface::SimulatorInterface.
//
// Errors: the reference asserts (asserts are live in its build, CMakeLists.txt:66).  Here argument errors and
// device errors throw std::runtime_error with the library's message; conditions the reference would abort on
// inside the simulation are reported per particle in SimulationResult::flags (FKS_FLAG_WOULD_ASSERT_*).
#ifndef FKSGPU_SIMULATOR_HPP
#define FKSGPU_SIMULATOR_HPP

#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fksgpu.h"

namespace fksgpu {

inline void Check(int code) {
    if (code != FKS_OK) throw std::runtime_error(std::string("fksgpu: ") + fks_last_error_string());
}

inline fks_solver_params GetDefaultSolverParameters() {
    fks_solver_params p;
    fks_default_solver_params(&p);
    return p;
}

// simple_simulator_interface::SimulationResult<Configuration> (spcs.hpp:918) + the per-particle counters
template <typename Configuration>
struct SimulationResult {
    Configuration result_config;
    Configuration actual_target;
    bool did_contact;
    bool outcome_is_nominal;
    uint32_t flags, n_microsteps, n_resolver_iterations, n_steps;
};

// simple_simulator_interface::ForwardSimulationStepTrace (filled at spcs.hpp:1583-1617, :1703, :1714, :1778): one entry per
// controller step holding the control input, its per-microstep share, and per microstep the configurations the resolver
// went through (post-action configuration, every correction step, the returned previous configuration on a stop).
template <typename Configuration>
struct ForwardSimulationContactResolverStepTrace {
    std::vector<Configuration> contact_resolution_steps;
};
template <typename Configuration>
struct ForwardSimulationResolverTrace {
    std::vector<double> control_input;
    std::vector<double> control_input_step;
    std::vector<ForwardSimulationContactResolverStepTrace<Configuration>> contact_resolver_steps;
};
template <typename Configuration>
struct ForwardSimulationStepTrace {
    std::vector<ForwardSimulationResolverTrace<Configuration>> resolver_steps;
};

// Environment = what the reference simulator copies at construction (spcs.hpp:420)
class Environment {
public:
    // BuildCompleteEnvironment (simulator_environment_builder.cpp:470-476)
    Environment(const std::vector<fks_obstacle>& obstacles, double resolution) : built_(nullptr) {
        Check(fks_build_environment(obstacles.data(), obstacles.size(), resolution, &built_));
    }
    ~Environment() { fks_built_env_destroy(built_); }
    Environment(const Environment&) = delete;
    Environment& operator=(const Environment&) = delete;
    const fks_env_desc& Description() const { return *fks_built_env_desc(built_); }

private:
    fks_built_env* built_;
};

// The same environment built by CUDA kernels directly in device memory (fks_env_build_device): no host grids.
class DeviceEnvironment {
public:
    // BuildCompleteEnvironment (simulator_environment_builder.cpp:470-476) on `device`
    DeviceEnvironment(const std::vector<fks_obstacle>& obstacles, double resolution, int device = 0) : env_(nullptr) {
        Check(fks_env_build_device(device, obstacles.data(), obstacles.size(), resolution, &env_));
    }
    ~DeviceEnvironment() { fks_env_destroy(env_); }
    DeviceEnvironment(const DeviceEnvironment&) = delete;
    DeviceEnvironment& operator=(const DeviceEnvironment&) = delete;
    fks_env* Handle() const { return env_; }
    // milliseconds: total, rasterise, z / y / x passes, surface marking, normal emit, distance-field check, table allocation
    std::vector<double> BuildTimingsMs() const {
        std::vector<double> ms(9, 0.0);
        Check(fks_env_build_timings(env_, ms.data(), (int)ms.size()));
        return ms;
    }

private:
    fks_env* env_;
};

template <typename Configuration>
struct ConfigTraits;  // Flatten(config, double*) / Unflatten(const double*, stride) -> config

template <>
struct ConfigTraits<std::vector<double>> {
    static void Flatten(const std::vector<double>& c, double* out, int stride) {
        if ((int)c.size() != stride) throw std::invalid_argument("fksgpu: configuration has the wrong size");
        std::memcpy(out, c.data(), sizeof(double) * (size_t)stride);
    }
    static std::vector<double> Unflatten(const double* in, int stride) { return std::vector<double>(in, in + stride); }
};

template <typename Configuration>
class GpuParticleContactSimulator {
public:
    typedef SimulationResult<Configuration> Result;

    GpuParticleContactSimulator(const fks_env_desc& environment, const fks_robot_desc& robot, const fks_solver_params& solver_config,
                                double simulation_controller_frequency, uint64_t prng_seed, int32_t debug_level, int device = 0)
        : env_(nullptr), robot_(nullptr), sim_(nullptr) {
        try {
            Check(fks_env_create(device, &environment, &env_));
            Check(fks_robot_create(device, &robot, &robot_));
            Check(fks_sim_create(env_, robot_, &solver_config, simulation_controller_frequency, prng_seed, debug_level, &sim_));
        } catch (...) {
            Release();
            throw;
        }
        stride_ = fks_robot_config_stride(robot_);
        record_ = fks_sim_result_stride(sim_);
        dof_ = robot.n_dof;
    }
    // same, in an environment that already lives on the device (shared, must outlive the simulator)
    GpuParticleContactSimulator(const std::shared_ptr<DeviceEnvironment>& environment, const fks_robot_desc& robot,
                                const fks_solver_params& solver_config, double simulation_controller_frequency, uint64_t prng_seed,
                                int32_t debug_level, int device = 0)
        : shared_env_(environment), env_(nullptr), robot_(nullptr), sim_(nullptr) {
        if (!environment) throw std::invalid_argument("fksgpu: null device environment");
        try {
            Check(fks_robot_create(device, &robot, &robot_));
            Check(fks_sim_create(environment->Handle(), robot_, &solver_config, simulation_controller_frequency, prng_seed, debug_level, &sim_));
        } catch (...) {
            Release();
            throw;
        }
        stride_ = fks_robot_config_stride(robot_);
        record_ = fks_sim_result_stride(sim_);
        dof_ = robot.n_dof;
    }
    ~GpuParticleContactSimulator() { Release(); }
    GpuParticleContactSimulator(const GpuParticleContactSimulator&) = delete;
    GpuParticleContactSimulator& operator=(const GpuParticleContactSimulator&) = delete;

    // spcs.hpp:788: one result per start; target_positions.size() must be 1 or start_positions.size() (assert :790-793)
    std::vector<Result> ForwardSimulateRobots(const std::vector<Configuration>& start_positions,
                                              const std::vector<Configuration>& target_positions, bool allow_contacts,
                                              const std::function<void(const void*)>& display_fn = nullptr) {
        (void)display_fn;  // RViz markers: not on the hot path (spcs.hpp:1725-1736 only fires at debug_level >= 2)
        return Simulate(start_positions, target_positions, allow_contacts, false);
    }
    // spcs.hpp:806 (ReverseSimulateMutableRobot forwards to the forward path, :838-841)
    std::vector<Result> ReverseSimulateRobots(const std::vector<Configuration>& start_positions,
                                              const std::vector<Configuration>& target_positions, bool allow_contacts,
                                              const std::function<void(const void*)>& display_fn = nullptr) {
        (void)display_fn;
        return Simulate(start_positions, target_positions, allow_contacts, true);
    }
    // spcs.hpp:824: single particle
    Result ForwardSimulateRobot(const Configuration& start, const Configuration& target, bool allow_contacts) {
        return ForwardSimulateRobots(std::vector<Configuration>(1, start), std::vector<Configuration>(1, target), allow_contacts)[0];
    }

    // spcs.hpp:824 with the trace arguments: single particle, trace filled when enable_tracing is set
    Result ForwardSimulateRobot(const Configuration& start, const Configuration& target, bool allow_contacts,
                                ForwardSimulationStepTrace<Configuration>& trace, bool enable_tracing) {
        if (!enable_tracing) return ForwardSimulateRobot(start, target, allow_contacts);
        std::vector<double> s0((size_t)stride_), t0((size_t)stride_);
        ConfigTraits<Configuration>::Flatten(start, s0.data(), stride_);
        ConfigTraits<Configuration>::Flatten(target, t0.data(), stride_);
        const size_t rec = fks_sim_trace_stride(sim_);
        std::vector<unsigned char> result(record_), records;
        size_t capacity = 4096, n = 0;
        for (;;) {  // grow until the whole trace fits (the simulation is deterministic in the particle id)
            records.resize(capacity * rec);
            Check(fks_forward_simulate_traced(sim_, s0.data(), t0.data(), allow_contacts ? 1 : 0, FKS_NOISE_PHILOX, nullptr,
                                              next_particle_id_, result.data(), records.data(), capacity, &n));
            if (n <= capacity) break;
            capacity = n;
        }
        next_particle_id_ += 1;
        for (size_t i = 0; i < n; i++) {
            fks_trace_header h;
            std::memcpy(&h, records.data() + i * rec, sizeof(h));
            const double* v = reinterpret_cast<const double*>(records.data() + i * rec + sizeof(h));
            if (h.kind == FKS_TRACE_CONTROL_INPUT) {
                trace.resolver_steps.emplace_back();
                trace.resolver_steps.back().control_input.assign(v, v + dof_);
            } else if (h.kind == FKS_TRACE_CONTROL_INPUT_STEP) {
                trace.resolver_steps.back().control_input_step.assign(v, v + dof_);
            } else {
                auto& steps = trace.resolver_steps.back().contact_resolver_steps;
                if (h.kind == FKS_TRACE_POST_ACTION) steps.emplace_back();
                steps.back().contact_resolution_steps.push_back(ConfigTraits<Configuration>::Unflatten(v, stride_));
            }
        }
        return Unpack(result.data(), target);
    }

    // spcs.hpp:1398: planner-side static query (environment inflated by inflation_ratio cells, self collisions)
    bool CheckConfigCollision(const Configuration& config, double inflation_ratio) {
        std::vector<double> flat((size_t)stride_);
        ConfigTraits<Configuration>::Flatten(config, flat.data(), stride_);
        uint8_t out = 0;
        Check(fks_check_config_collision(sim_, flat.data(), 1, inflation_ratio, &out));
        return out != 0;
    }
    std::vector<uint8_t> CheckConfigCollisions(const std::vector<Configuration>& configs, double inflation_ratio) {
        std::vector<double> flat(configs.size() * (size_t)stride_);
        for (size_t i = 0; i < configs.size(); i++) ConfigTraits<Configuration>::Flatten(configs[i], flat.data() + i * (size_t)stride_, stride_);
        std::vector<uint8_t> out(configs.size(), 0);
        Check(fks_check_config_collision(sim_, flat.data(), configs.size(), inflation_ratio, out.data()));
        return out;
    }

    // spcs.hpp:488-500, same keys
    std::map<std::string, double> GetStatistics() {
        uint64_t s[FKS_NUM_STATS];
        Check(fks_get_statistics(sim_, s));
        std::map<std::string, double> out;
        out["successful_resolves"] = (double)s[FKS_STAT_SUCCESSFUL_RESOLVES];
        out["unsuccessful_resolves"] = (double)s[FKS_STAT_UNSUCCESSFUL_RESOLVES];
        out["free_resolves"] = (double)s[FKS_STAT_FREE_RESOLVES];
        out["collision_resolves"] = (double)s[FKS_STAT_COLLISION_RESOLVES];
        out["fallback_resolves"] = (double)s[FKS_STAT_FALLBACK_RESOLVES];
        out["unsuccessful_self_collision_resolves"] = (double)s[FKS_STAT_UNSUCCESSFUL_SELF_COLLISION_RESOLVES];
        out["unsuccessful_env_collision_resolves"] = (double)s[FKS_STAT_UNSUCCESSFUL_ENV_COLLISION_RESOLVES];
        out["recovered_unsuccessful_resolves"] = (double)s[FKS_STAT_RECOVERED_UNSUCCESSFUL_RESOLVES];
        return out;
    }
    void ResetStatistics() { Check(fks_reset_statistics(sim_)); }  // spcs.hpp:502-512

    // noise: Philox keyed by (prng_seed, first_particle_id + index); the id offset advances per call so that
    // successive calls draw fresh noise, as the reference's per-thread generators do
    void SetNextParticleId(uint64_t id) { next_particle_id_ = id; }
    int ConfigStride() const { return stride_; }
    const char* KernelInfo() { return fks_sim_kernel_info(sim_); }

private:
    std::vector<Result> Simulate(const std::vector<Configuration>& starts, const std::vector<Configuration>& targets,
                                 bool allow_contacts, bool reverse) {
        const size_t n = starts.size();
        if (!(targets.size() == 1 || targets.size() == n)) throw std::invalid_argument("fksgpu: need 1 target or one per start");
        std::vector<double> hs(n * (size_t)stride_), ht(targets.size() * (size_t)stride_);
        for (size_t i = 0; i < n; i++) ConfigTraits<Configuration>::Flatten(starts[i], hs.data() + i * (size_t)stride_, stride_);
        for (size_t i = 0; i < targets.size(); i++) ConfigTraits<Configuration>::Flatten(targets[i], ht.data() + i * (size_t)stride_, stride_);
        std::vector<unsigned char> rec(n * record_);
        Check((reverse ? fks_reverse_simulate : fks_forward_simulate)(sim_, hs.data(), ht.data(), n, targets.size(), allow_contacts ? 1 : 0,
                                                                     FKS_NOISE_PHILOX, nullptr, next_particle_id_, rec.data()));
        next_particle_id_ += n;
        std::vector<Result> out(n);
        for (size_t i = 0; i < n; i++) out[i] = Unpack(rec.data() + i * record_, targets.size() == 1 ? targets[0] : targets[i]);
        return out;
    }
    Result Unpack(const unsigned char* r, const Configuration& target) const {
        Result out;
        fks_result_tail tail;
        std::memcpy(&tail, r + sizeof(double) * (size_t)stride_, sizeof(tail));
        out.result_config = ConfigTraits<Configuration>::Unflatten(reinterpret_cast<const double*>(r), stride_);
        out.actual_target = target;
        out.did_contact = (tail.flags & FKS_FLAG_DID_CONTACT) != 0;  // spcs.hpp:918
        out.outcome_is_nominal = true;                               // always true in the reference (:918)
        out.flags = tail.flags;
        out.n_microsteps = tail.n_microsteps;
        out.n_resolver_iterations = tail.n_resolver_iters;
        out.n_steps = tail.n_steps;
        return out;
    }
    void Release() {
        fks_sim_destroy(sim_);
        fks_robot_destroy(robot_);
        fks_env_destroy(env_);
        sim_ = nullptr;
        robot_ = nullptr;
        env_ = nullptr;
    }

    std::shared_ptr<DeviceEnvironment> shared_env_;  // set when the environment is not owned (env_ stays null)
    fks_env* env_;
    fks_robot* robot_;
    fks_sim* sim_;
    int stride_ = 0, dof_ = 0;
    size_t record_ = 0;
    uint64_t next_particle_id_ = 0;
};

typedef GpuParticleContactSimulator<std::vector<double>> GpuSimulator;
typedef std::shared_ptr<GpuSimulator> GpuSimulatorPtr;

// fast_kinematic_simulator.hpp:18-22, same argument order (grid + SDF + surface normals = one fks_env_desc)
inline GpuSimulatorPtr MakeGpuSimulator(int kind, const fks_env_desc& environment, const fks_robot_desc& robot,
                                        const fks_solver_params& solver_config, double simulation_controller_frequency,
                                        uint64_t prng_seed, int32_t debug_level) {
    if (robot.kind != kind) throw std::invalid_argument("fksgpu: robot description does not match the factory");
    return GpuSimulatorPtr(new GpuSimulator(environment, robot, solver_config, simulation_controller_frequency, prng_seed, debug_level));
}
inline GpuSimulatorPtr MakeGpuSimulator(int kind, const std::shared_ptr<DeviceEnvironment>& environment, const fks_robot_desc& robot,
                                        const fks_solver_params& solver_config, double simulation_controller_frequency,
                                        uint64_t prng_seed, int32_t debug_level) {
    if (robot.kind != kind) throw std::invalid_argument("fksgpu: robot description does not match the factory");
    return GpuSimulatorPtr(new GpuSimulator(environment, robot, solver_config, simulation_controller_frequency, prng_seed, debug_level));
}
inline GpuSimulatorPtr MakeGpuSE2Simulator(const fks_env_desc& e, const fks_robot_desc& r, const fks_solver_params& p, double f, uint64_t seed, int32_t dbg) {
    return MakeGpuSimulator(FKS_ROBOT_SE2, e, r, p, f, seed, dbg);
}
inline GpuSimulatorPtr MakeGpuSE3Simulator(const fks_env_desc& e, const fks_robot_desc& r, const fks_solver_params& p, double f, uint64_t seed, int32_t dbg) {
    return MakeGpuSimulator(FKS_ROBOT_SE3, e, r, p, f, seed, dbg);
}
inline GpuSimulatorPtr MakeGpuLinkedSimulator(const fks_env_desc& e, const fks_robot_desc& r, const fks_solver_params& p, double f, uint64_t seed, int32_t dbg) {
    return MakeGpuSimulator(FKS_ROBOT_LINKED, e, r, p, f, seed, dbg);
}

}  // namespace fksgpu

#endif  // FKSGPU_SIMULATOR_HPP
