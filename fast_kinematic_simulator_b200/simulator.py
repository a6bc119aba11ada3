"""Host-side mirror of the reference's simulator interface for the batched forward-simulate path.

Names follow the reference:

* ``build_complete_environment``  <- simulator_environment_builder::BuildCompleteEnvironment
  (src/fast_kinematic_simulator/simulator_environment_builder.cpp:470-476)
* ``make_{se2,se3,linked}_simulator`` <- fast_kinematic_simulator::Make{SE2,SE3,Linked}Simulator
  (include/fast_kinematic_simulator/fast_kinematic_simulator.hpp:18-22)
* ``GpuParticleContactSimulator.forward_simulate_robots`` <- SimpleParticleContactSimulator::
  ForwardSimulateRobots (simple_particle_contact_simulator.hpp:788-804); ``reverse_simulate_robots``
  (:806-822); ``get_statistics`` / ``reset_statistics`` (:488-512).

Everything computes in libfksgpu.so; this module only marshals numpy arrays across the C ABI.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import lib, check


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


IDENTITY12 = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float64)


def make_transform(translation=(0.0, 0.0, 0.0), rotation=None):
    """Row-major 3x4 [R|t] as 12 doubles."""
    T = np.zeros((3, 4))
    T[:, :3] = np.eye(3) if rotation is None else np.asarray(rotation, dtype=np.float64)
    T[:, 3] = translation
    return T.reshape(12).copy()


class BuiltEnvironment:
    """Result of BuildCompleteEnvironment: collision-map metadata + SDF + surface-normal table,
    owned by the C library; numpy views are exposed for tests and the oracle."""

    def __init__(self, handle):
        self._h = handle
        self.desc = lib.fks_built_env_desc(handle).contents
        d = self.desc
        self.shape = (int(d.nx), int(d.ny), int(d.nz))
        n = self.shape[0] * self.shape[1] * self.shape[2]
        self.resolution = float(d.sdf_resolution)
        self.sdf = np.ctypeslib.as_array(d.sdf, shape=(n,)).reshape(self.shape)
        occ = lib.fks_built_env_occupancy(handle)
        self.occupancy = np.ctypeslib.as_array(occ, shape=(n,)).reshape(self.shape) if occ else None
        self.origin = np.array(list(d.origin))
        self.inverse_origin = np.array(list(d.inverse_origin))
        nc = int(d.n_normal_cells)
        self.n_normal_cells = nc
        if nc:
            self.normal_cell_index = np.ctypeslib.as_array(d.normal_cell_index, shape=(nc,))
            self.normal_cell_start = np.ctypeslib.as_array(d.normal_cell_start, shape=(nc + 1,))
            ne = int(self.normal_cell_start[-1])
            self.normal_entries = np.ctypeslib.as_array(d.normal_entries, shape=(ne * 7,)).reshape(ne, 7)
        else:
            self.normal_cell_index = np.zeros(0, np.int64)
            self.normal_cell_start = np.zeros(1, np.uint32)
            self.normal_entries = np.zeros((0, 7))

    def close(self):
        if self._h:
            lib.fks_built_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_complete_environment(obstacles, resolution):
    """obstacles: iterable of (pose12, half_extents3, object_id) -- OBSTACLE_CONFIG
    (simulator_environment_builder.hpp:25-49)."""
    arr, n = _obstacle_array(obstacles)
    h = C.c_void_p()
    check(lib.fks_build_environment(arr, n, float(resolution), C.byref(h)))
    return BuiltEnvironment(h)


def _obstacle_array(obstacles):
    obstacles = list(obstacles)
    arr = (capi.Obstacle * max(len(obstacles), 1))()
    for i, (pose, ext, oid) in enumerate(obstacles):
        arr[i].pose = (C.c_double * 12)(*[float(v) for v in pose])
        arr[i].extents = (C.c_double * 3)(*[float(v) for v in ext])
        arr[i].object_id = int(oid)
    return arr, len(obstacles)


BUILD_PHASES = ("total", "rasterize", "edt_z", "edt_y", "edt_x_sdf", "normals_mark", "normals_emit", "distance_field_check",
                "table_allocation")


def build_complete_environment_on_device(obstacles, resolution, device=0):
    """BuildCompleteEnvironment (simulator_environment_builder.cpp:470-476) computed by CUDA kernels straight into a
    device environment (fks_env_build_device): no host copy of the occupancy grid, SDF or normal table."""
    arr, n = _obstacle_array(obstacles)
    h = C.c_void_p()
    check(lib.fks_env_build_device(int(device), arr, n, float(resolution), C.byref(h)))
    return GpuEnvironment(None, device, _handle=h)


class RobotDescription:
    """Flat description of a Tnuva{SE2,SE3,Linked}Robot (tnuva_robot_models.hpp:26,201,415): collision
    points per link (link-major), one PID + velocity actuator per axis, joints for the linked robot."""

    def __init__(self, kind, points_xyz, point_link, axes, n_links=1, joints=(), base_transform=None,
                 allowed_self_collision=None, position_distance_weight=1.0, rotation_distance_weight=1.0):
        self.kind = int(kind)
        self.points = _as_f64(points_xyz).reshape(-1, 3)
        self.point_link = np.ascontiguousarray(point_link, dtype=np.int32)
        self.axes = [dict(a) for a in axes]
        self.n_links = int(n_links)
        self.joints = [dict(j) for j in joints]
        self.base = _as_f64(IDENTITY12 if base_transform is None else base_transform)
        self.allowed = None if allowed_self_collision is None else np.ascontiguousarray(allowed_self_collision, dtype=np.uint8)
        self.pos_w = float(position_distance_weight)
        self.rot_w = float(rotation_distance_weight)
        self.n_dof = len(self.axes)
        self._keep = None

    @property
    def config_stride(self):
        return 3 if self.kind == capi.ROBOT_SE2 else (12 if self.kind == capi.ROBOT_SE3 else self.n_dof)

    def to_c(self):
        d = capi.RobotDesc()
        d.kind = self.kind
        d.n_links = self.n_links
        d.n_joints = len(self.joints)
        d.n_dof = self.n_dof
        d.n_points = self.points.shape[0]
        d.points_xyz = _dptr(self.points)
        d.point_link = self.point_link.ctypes.data_as(C.POINTER(C.c_int32))
        axes = (capi.AxisParams * self.n_dof)()
        for i, a in enumerate(self.axes):
            axes[i].kp = a.get("kp", 1.0)
            axes[i].ki = a.get("ki", 0.0)
            axes[i].kd = a.get("kd", 0.0)
            axes[i].integral_clamp = a.get("integral_clamp", 0.0)
            axes[i].velocity_limit = a["velocity_limit"]
            axes[i].proportional_noise = a.get("proportional_noise", 0.0)
            axes[i].minimum_noise = a.get("minimum_noise", 0.0)
            axes[i].noise_sigma = a.get("noise_sigma", 0.5)  # tnuva.hpp:128-130
        d.axes = axes
        d.base_transform = (C.c_double * 12)(*self.base.tolist())
        joints = (capi.JointDesc * max(len(self.joints), 1))()
        for i, j in enumerate(self.joints):
            joints[i].parent_link = j["parent"]
            joints[i].child_link = j["child"]
            joints[i].type = j["type"]
            joints[i].transform = (C.c_double * 12)(*_as_f64(j["transform"]).tolist())
            joints[i].axis = (C.c_double * 3)(*[float(v) for v in j["axis"]])
            joints[i].lower_limit = j.get("lower", -np.pi)
            joints[i].upper_limit = j.get("upper", np.pi)
            joints[i].distance_weight = j.get("weight", 1.0)
        d.joints = joints
        if self.allowed is not None:
            d.allowed_self_collision = self.allowed.ctypes.data_as(C.POINTER(C.c_uint8))
        d.position_distance_weight = self.pos_w
        d.rotation_distance_weight = self.rot_w
        self._keep = (axes, joints)
        return d


class GpuEnvironment:
    """Device copy of what the simulator copies at construction (spcs.hpp:420): SDF + normals."""

    def __init__(self, built_env, device=0, _handle=None):
        self.built = built_env
        self.device = device
        if _handle is not None:  # built on the device (build_complete_environment_on_device)
            self._h = _handle
            return
        self._h = C.c_void_p()
        desc = built_env.desc if isinstance(built_env, BuiltEnvironment) else built_env
        check(lib.fks_env_create(device, C.byref(desc), C.byref(self._h)))

    def download(self):
        """Host copy of the device environment (fks_env_download) as a BuiltEnvironment."""
        h = C.c_void_p()
        check(lib.fks_env_download(self._h, C.byref(h)))
        return BuiltEnvironment(h)

    @property
    def build_timings_ms(self):
        """Device time of each phase of the device builder (zeros for uploaded environments)."""
        out = (C.c_double * len(BUILD_PHASES))()
        check(lib.fks_env_build_timings(self._h, out, len(BUILD_PHASES)))
        return {k: float(out[i]) for i, k in enumerate(BUILD_PHASES)}

    def close(self):
        if self._h:
            lib.fks_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuRobot:
    def __init__(self, description, device=0):
        self.description = description
        self.device = device
        self._h = C.c_void_p()
        d = description.to_c()
        check(lib.fks_robot_create(device, C.byref(d), C.byref(self._h)))
        self.config_stride = lib.fks_robot_config_stride(self._h)

    def close(self):
        if self._h:
            lib.fks_robot_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def result_dtype(stride):
    """fks_result record: cfg_stride doubles followed by fks_result_tail."""
    return np.dtype([("cfg", np.float64, (stride,)), ("flags", np.uint32), ("n_microsteps", np.uint32),
                     ("n_resolver_iters", np.uint32), ("n_steps", np.uint32)])


class SimulationResults:
    """vector<SimulationResult> (spcs.hpp:918): result_config, did_contact, plus the per-particle
    counters the parity tests and the metric need."""

    def __init__(self, records):
        self.records = records
        self.configs = records["cfg"]
        self.flags = records["flags"]
        self.n_microsteps = records["n_microsteps"]
        self.n_resolver_iters = records["n_resolver_iters"]
        self.n_steps = records["n_steps"]

    @property
    def did_contact(self):
        return (self.flags & capi.FLAG_DID_CONTACT) != 0

    @property
    def resolve_failed(self):
        return (self.flags & capi.FLAG_RESOLVE_FAILED) != 0

    def __len__(self):
        return len(self.records)


def make_tape(draws, offsets, decisions=None, decision_offsets=None):
    """fks_noise_tape: the injected truncated-normal draws and, optionally, the decision tape (include/fksgpu.h)."""
    draws = _as_f64(draws)
    if not draws.size:
        draws = np.zeros(1)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    t = capi.NoiseTape()
    t.draws = _dptr(draws)
    t.offsets = offsets.ctypes.data_as(C.POINTER(C.c_uint64))
    keep = [draws, offsets]
    if decisions is not None:
        decisions = np.ascontiguousarray(decisions, dtype=np.uint64)
        if not decisions.size:
            decisions = np.zeros(1, dtype=np.uint64)
        decision_offsets = np.ascontiguousarray(decision_offsets, dtype=np.uint64)
        t.decisions = decisions.ctypes.data_as(C.POINTER(C.c_uint64))
        t.decision_offsets = decision_offsets.ctypes.data_as(C.POINTER(C.c_uint64))
        keep += [decisions, decision_offsets]
    return t, keep


class GpuParticleContactSimulator:
    """The GPU sibling of SimpleParticleContactSimulator for the batch calls."""

    def __init__(self, env, robot, solver_params=None, simulation_controller_frequency=25.0, prng_seed=42, debug_level=0):
        self.env = env
        self.robot = robot
        self.params = solver_params if solver_params is not None else capi.default_solver_params()
        self._h = C.c_void_p()
        check(lib.fks_sim_create(env._h, robot._h, C.byref(self.params), float(simulation_controller_frequency),
                                 int(prng_seed), int(debug_level), C.byref(self._h)))
        self.config_stride = robot.config_stride
        self.result_stride = lib.fks_sim_result_stride(self._h)
        self.dtype = result_dtype(self.config_stride)
        assert self.dtype.itemsize == self.result_stride

    # -- ForwardSimulateRobots / ReverseSimulateRobots (host buffers; H2D + kernel + D2H inside) --
    def _simulate(self, fn, starts, targets, allow_contacts, noise_mode, tape, first_particle_id, out):
        starts = _as_f64(starts).reshape(-1, self.config_stride)
        targets = _as_f64(targets).reshape(-1, self.config_stride)
        n = starts.shape[0]
        if out is None:
            out = np.empty(n, dtype=self.dtype)
        ctape, keep = (None, None)
        if tape is not None:
            ctape, keep = make_tape(*tape)
        check(fn(self._h, starts.ctypes.data, targets.ctypes.data, n, targets.shape[0], int(bool(allow_contacts)),
                 int(noise_mode), C.byref(ctape) if ctape is not None else None, int(first_particle_id), out.ctypes.data))
        return SimulationResults(out)

    def forward_simulate_robots(self, starts, targets, allow_contacts=True, noise_mode=capi.NOISE_PHILOX, tape=None,
                                first_particle_id=0, out=None):
        return self._simulate(lib.fks_forward_simulate, starts, targets, allow_contacts, noise_mode, tape, first_particle_id, out)

    def reverse_simulate_robots(self, starts, targets, allow_contacts=True, noise_mode=capi.NOISE_PHILOX, tape=None,
                                first_particle_id=0, out=None):
        return self._simulate(lib.fks_reverse_simulate, starts, targets, allow_contacts, noise_mode, tape, first_particle_id, out)

    # -- device-resident variant (torch tensors or raw pointers), asynchronous on `stream` --------
    def forward_simulate_device(self, d_starts, d_targets, n, n_targets, d_results, allow_contacts=True,
                                noise_mode=capi.NOISE_PHILOX, d_tape=None, d_tape_offsets=None, first_particle_id=0, stream=0):
        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else int(x)

        check(lib.fks_forward_simulate_device(self._h, ptr(d_starts), ptr(d_targets), int(n), int(n_targets),
                                              int(bool(allow_contacts)), int(noise_mode), ptr(d_tape), ptr(d_tape_offsets),
                                              int(first_particle_id), ptr(d_results), int(stream) if stream else None))

    def trace_dtype(self):
        width = (lib.fks_sim_trace_stride(self._h) - 16) // 8
        return np.dtype([("kind", np.uint32), ("step", np.uint32), ("microstep", np.uint32), ("iteration", np.uint32),
                         ("values", np.float64, (width,))])

    def forward_simulate_robot_traced(self, start, target, allow_contacts=True, noise_mode=capi.NOISE_PHILOX, tape=None,
                                      particle_id=0, capacity=65536):
        """ForwardSimulateRobot with enable_tracing (spcs.hpp:824-829): (SimulationResults of one particle, trace records).
        The flat records follow the order in which the reference fills ForwardSimulationStepTrace (spcs.hpp:1583-1617,
        :1703, :1714, :1778); `values` holds n_dof control values or cfg_stride configuration values."""
        start = _as_f64(start).reshape(1, self.config_stride)
        target = _as_f64(target).reshape(1, self.config_stride)
        out = np.empty(1, dtype=self.dtype)
        rec = np.zeros(capacity, dtype=self.trace_dtype())
        ctape, keep = (None, None)
        if tape is not None:
            ctape, keep = make_tape(*tape)
        n = C.c_size_t(0)
        check(lib.fks_forward_simulate_traced(self._h, start.ctypes.data, target.ctypes.data, int(bool(allow_contacts)), int(noise_mode),
                                              C.byref(ctape) if ctape is not None else None, int(particle_id), out.ctypes.data,
                                              rec.ctypes.data, capacity, C.byref(n)))
        if n.value > capacity:
            raise capi.FksError(capi.ERR_INVALID_ARGUMENT, "trace capacity %d too small for %d records" % (capacity, n.value))
        return SimulationResults(out), rec[: n.value]

    def check_config_collision(self, configs, inflation_ratio=0.0):
        """CheckConfigCollision (spcs.hpp:1398-1416) for a batch of configurations -> bool array."""
        configs = _as_f64(configs).reshape(-1, self.config_stride)
        out = np.zeros(configs.shape[0], dtype=np.uint8)
        check(lib.fks_check_config_collision(self._h, configs.ctypes.data, configs.shape[0], float(inflation_ratio), out.ctypes.data))
        return out.astype(bool)

    def get_statistics(self):
        out = (C.c_uint64 * capi.NUM_STATS)()
        check(lib.fks_get_statistics(self._h, out))
        return {k: int(out[i]) for i, k in enumerate(capi.STAT_NAMES)}

    def reset_statistics(self):
        check(lib.fks_reset_statistics(self._h))

    @property
    def launch_count(self):
        return int(lib.fks_sim_launch_count(self._h))

    @property
    def kernel_info(self):
        return lib.fks_sim_kernel_info(self._h).decode()

    def close(self):
        if self._h:
            lib.fks_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _make(kind, built_env, robot_description, solver_params, simulation_controller_frequency, prng_seed, debug_level, device):
    if robot_description.kind != kind:
        raise ValueError("robot description kind does not match the factory")
    env = built_env if isinstance(built_env, GpuEnvironment) else GpuEnvironment(built_env, device)
    robot = GpuRobot(robot_description, device)
    return GpuParticleContactSimulator(env, robot, solver_params, simulation_controller_frequency, prng_seed, debug_level)


def make_se2_simulator(built_env, robot_description, solver_params=None, simulation_controller_frequency=25.0, prng_seed=42,
                       debug_level=0, device=0):
    """MakeSE2Simulator (fast_kinematic_simulator.cpp:4-25)."""
    return _make(capi.ROBOT_SE2, built_env, robot_description, solver_params, simulation_controller_frequency, prng_seed, debug_level, device)


def make_se3_simulator(built_env, robot_description, solver_params=None, simulation_controller_frequency=25.0, prng_seed=42,
                       debug_level=0, device=0):
    """MakeSE3Simulator (fast_kinematic_simulator.cpp:27-48)."""
    return _make(capi.ROBOT_SE3, built_env, robot_description, solver_params, simulation_controller_frequency, prng_seed, debug_level, device)


def make_linked_simulator(built_env, robot_description, solver_params=None, simulation_controller_frequency=25.0, prng_seed=42,
                          debug_level=0, device=0):
    """MakeLinkedSimulator (fast_kinematic_simulator.cpp:50-71)."""
    return _make(capi.ROBOT_LINKED, built_env, robot_description, solver_params, simulation_controller_frequency, prng_seed, debug_level, device)


def debug_qr_solve(systems, device=0):
    """fks_debug_qr_solve: the device's stacked-Jacobian solver on a list of (A rows x cols, b) systems sharing `cols`.
    Returns (solutions n x cols, flags n)."""
    n = len(systems)
    cols = int(systems[0][0].shape[1])
    rows = np.array([A.shape[0] for A, _ in systems], dtype=np.int32)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(rows.astype(np.uint64) * np.uint64(cols + 1))
    flat = np.empty(int(offsets[-1]), dtype=np.float64)
    for i, (A, b) in enumerate(systems):
        o = int(offsets[i])
        r = int(rows[i])
        flat[o:o + r * cols] = np.asarray(A, dtype=np.float64).T.reshape(-1)  # column major
        flat[o + r * cols:o + r * (cols + 1)] = b
    x = np.zeros((n, cols))
    flags = np.zeros(n, dtype=np.uint32)
    check(lib.fks_debug_qr_solve(int(device), flat.ctypes.data, offsets.ctypes.data, rows.ctypes.data, cols, n, x.ctypes.data,
                                 flags.ctypes.data))
    return x, flags
