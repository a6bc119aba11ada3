"""TEST INFRASTRUCTURE -- ctypes binding of oracle/libfks_oracle.so (the CPU restatement of the
reference's forward-simulate path).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfks_oracle.so")
PID_REF = os.path.join(_HERE, "_ref", "pid_ref")
UNC_REF = os.path.join(_HERE, "_ref", "unc_ref")

ORACLE_NOISE_MT19937 = 3
SENS_NAMES = ("cell_boundary", "est_threshold", "est_zero", "rank_cut", "pivot_tie", "nmicro", "normal_tie",
              "angle_wrap", "self_collision", "raw_threshold", "step_fraction", "ill_conditioned")
SENS_RANK_CUT = 1 << 3
SENS_PIVOT_TIE = 1 << 4
SENS_SELF_COLLISION = 1 << 8
SENS_ILL_CONDITIONED = 1 << 11
# Sensitivities that stop mattering once the GPU consumes the oracle's DECISION TAPE (QR rank / pivot order injected), plus the
# informational self-collision bit (the impulse solve is compared like everything else).
SENS_COVERED_BY_DECISION_TAPE = SENS_RANK_CUT | SENS_PIVOT_TIE | SENS_SELF_COLLISION
# decision record word 0 (include/fksgpu.h)
DECISION_OVERRIDE_SOLUTION, DECISION_ROUNDOFF_PIVOT, DECISION_PIVOT_TIE = 1 << 8, 1 << 9, 1 << 10


def build():
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def load():
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.c_int]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_destroy.restype = None
    lib.oracle_num_threads.argtypes = [C.c_void_p]
    lib.oracle_config_stride.argtypes = [C.c_void_p]
    lib.oracle_forward_simulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                            C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    lib.oracle_check_config_collision.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
    lib.oracle_check_config_collision.restype = None
    lib.oracle_actuate_run.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_actuate_run.restype = None
    lib.oracle_end_states_partition.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.oracle_end_states_partition.restype = None
    lib.oracle_pairwise_config_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.oracle_pairwise_config_distance.restype = None
    lib.oracle_tape_total.argtypes = [C.c_void_p]
    lib.oracle_tape_total.restype = C.c_uint64
    lib.oracle_copy_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_copy_tape.restype = None
    lib.oracle_copy_sensitivity.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_copy_sensitivity.restype = None
    lib.oracle_decision_words.argtypes = [C.c_void_p]
    lib.oracle_decision_words.restype = C.c_uint64
    lib.oracle_copy_decisions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_copy_decisions.restype = None
    lib.oracle_copy_max_condition.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_copy_max_condition.restype = None
    lib.oracle_set_qr_model.argtypes = [C.c_void_p, C.c_int]
    lib.oracle_set_qr_model.restype = None
    lib.oracle_set_decision_cond_limit.argtypes = [C.c_void_p, C.c_double]
    lib.oracle_set_decision_cond_limit.restype = None
    lib.oracle_debug_capture_systems.argtypes = [C.c_size_t]
    lib.oracle_debug_capture_systems.restype = None
    lib.oracle_debug_captured_count.argtypes = []
    lib.oracle_debug_captured_count.restype = C.c_size_t
    lib.oracle_debug_captured_system.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_debug_captured_system.restype = None
    lib.oracle_colpiv_qr_solve_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_colpiv_qr_solve_info.restype = None
    lib.oracle_qr_device_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_qr_device_model.restype = None
    lib.oracle_get_statistics.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_get_statistics.restype = None
    lib.oracle_reset_statistics.argtypes = [C.c_void_p]
    lib.oracle_reset_statistics.restype = None
    lib.oracle_colpiv_qr_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_colpiv_qr_solve.restype = None
    lib.oracle_pid_run.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_pid_run.restype = None
    lib.oracle_exp_twist.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_exp_twist.restype = None
    lib.oracle_twist_between.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_twist_between.restype = None
    lib.oracle_wrap_angle.argtypes = [C.c_double]
    lib.oracle_wrap_angle.restype = C.c_double
    lib.oracle_truncated_normal.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_void_p]
    lib.oracle_truncated_normal.restype = None
    lib.oracle_philox_truncated_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
    lib.oracle_philox_truncated_normal.restype = C.c_double
    lib.oracle_philox_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_philox_raw.restype = None
    lib.oracle_env_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_robot_kinematics.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_trace_stride.argtypes = [C.c_void_p]
    lib.oracle_trace_stride.restype = C.c_size_t
    lib.oracle_forward_simulate_traced.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                                   C.c_void_p, C.c_void_p, C.c_size_t]
    lib.oracle_forward_simulate_traced.restype = C.c_size_t
    lib.oracle_build_environment.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
    lib.oracle_build_environment.restype = C.c_void_p
    lib.oracle_env_desc.argtypes = [C.c_void_p]
    lib.oracle_env_desc.restype = C.POINTER(_EnvDesc)
    lib.oracle_env_occupancy.argtypes = [C.c_void_p]
    lib.oracle_env_occupancy.restype = C.POINTER(C.c_uint8)
    lib.oracle_env_destroy.argtypes = [C.c_void_p]
    lib.oracle_env_destroy.restype = None
    return lib


class _Obstacle(C.Structure):  # fks_obstacle (include/fksgpu.h)
    _fields_ = [("pose", C.c_double * 12), ("extents", C.c_double * 3), ("object_id", C.c_uint32), ("_pad", C.c_uint32)]


class _EnvDesc(C.Structure):  # fks_env_desc (include/fksgpu.h)
    _fields_ = [("origin", C.c_double * 12), ("inverse_origin", C.c_double * 12), ("map_resolution", C.c_double),
                ("sdf_resolution", C.c_double), ("nx", C.c_int64), ("ny", C.c_int64), ("nz", C.c_int64),
                ("sdf", C.POINTER(C.c_float)), ("oob_value", C.c_float), ("_pad", C.c_int32), ("n_normal_cells", C.c_int64),
                ("normal_cell_index", C.POINTER(C.c_int64)), ("normal_cell_start", C.POINTER(C.c_uint32)),
                ("normal_entries", C.POINTER(C.c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def build_environment(obstacles, resolution):
    """oracle_build_environment: scalar restatement of BuildCompleteEnvironment (simulator_environment_builder.cpp:470-476).
    obstacles: iterable of (pose12, half_extents3, object_id).  Returns numpy copies of every array."""
    obstacles = list(obstacles)
    arr = (_Obstacle * max(len(obstacles), 1))()
    for i, (pose, ext, oid) in enumerate(obstacles):
        arr[i].pose = (C.c_double * 12)(*[float(v) for v in pose])
        arr[i].extents = (C.c_double * 3)(*[float(v) for v in ext])
        arr[i].object_id = int(oid)
    h = lib().oracle_build_environment(arr, len(obstacles), float(resolution))
    try:
        d = lib().oracle_env_desc(h).contents
        shape = (int(d.nx), int(d.ny), int(d.nz))
        n = shape[0] * shape[1] * shape[2]
        nc = int(d.n_normal_cells)
        out = {
            "shape": shape,
            "origin": np.array(list(d.origin)),
            "inverse_origin": np.array(list(d.inverse_origin)),
            "resolution": float(d.sdf_resolution),
            "sdf": np.ctypeslib.as_array(d.sdf, shape=(n,)).reshape(shape).copy(),
            "occupancy": np.ctypeslib.as_array(lib().oracle_env_occupancy(h), shape=(n,)).reshape(shape).copy(),
            "normal_cell_index": np.ctypeslib.as_array(d.normal_cell_index, shape=(nc,)).copy() if nc else np.zeros(0, np.int64),
            "normal_cell_start": np.ctypeslib.as_array(d.normal_cell_start, shape=(nc + 1,)).copy(),
        }
        ne = int(out["normal_cell_start"][-1])
        out["normal_entries"] = (np.ctypeslib.as_array(d.normal_entries, shape=(ne * 7,)).reshape(ne, 7).copy() if ne
                                 else np.zeros((0, 7)))
        return out
    finally:
        lib().oracle_env_destroy(h)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleSimulator:
    """CPU oracle of SimpleParticleContactSimulator::ForwardSimulateRobots.  `env_desc` / `robot_desc` are
    the same ctypes structures (include/fksgpu.h) the product consumes."""

    def __init__(self, env_desc, robot_desc, solver_params, frequency=25.0, seed=42, num_threads=0):
        self._keep = (env_desc, robot_desc, solver_params)
        self._h = lib().oracle_create(C.addressof(env_desc), C.addressof(robot_desc), C.addressof(solver_params),
                                      float(frequency), int(seed), int(num_threads))
        if not self._h:
            raise ValueError("oracle_create rejected the robot description")
        self.stride = lib().oracle_config_stride(self._h)
        self.num_threads = lib().oracle_num_threads(self._h)
        self.dtype = np.dtype([("cfg", np.float64, (self.stride,)), ("flags", np.uint32), ("n_microsteps", np.uint32),
                               ("n_resolver_iters", np.uint32), ("n_steps", np.uint32)])

    def forward_simulate(self, starts, targets, allow_contacts=True, noise_mode=ORACLE_NOISE_MT19937, tape=None,
                         first_particle_id=0, record_tape=False):
        starts = _f64(starts).reshape(-1, self.stride)
        targets = _f64(targets).reshape(-1, self.stride)
        n = starts.shape[0]
        out = np.empty(n, dtype=self.dtype)
        ctape = None
        keep = None
        if tape is not None:
            draws = _f64(tape[0])
            offs = np.ascontiguousarray(tape[1], dtype=np.uint64)
            if draws.size == 0:
                draws = np.zeros(1)

            class T(C.Structure):
                _fields_ = [("draws", C.c_void_p), ("offsets", C.c_void_p), ("decisions", C.c_void_p), ("decision_offsets", C.c_void_p)]

            ctape = T(draws.ctypes.data, offs.ctypes.data, None, None)
            keep = (draws, offs)
            if len(tape) >= 4:
                dec = np.ascontiguousarray(tape[2], dtype=np.uint64)
                if dec.size == 0:
                    dec = np.zeros(1, dtype=np.uint64)
                dec_offs = np.ascontiguousarray(tape[3], dtype=np.uint64)
                ctape = T(draws.ctypes.data, offs.ctypes.data, dec.ctypes.data, dec_offs.ctypes.data)
                keep = (draws, offs, dec, dec_offs)
        rc = lib().oracle_forward_simulate(self._h, starts.ctypes.data, targets.ctypes.data, n, targets.shape[0],
                                           int(bool(allow_contacts)), int(noise_mode),
                                           C.addressof(ctape) if ctape is not None else None, int(first_particle_id),
                                           int(bool(record_tape)), out.ctypes.data)
        del keep
        if rc != 0:
            raise ValueError("oracle_forward_simulate failed with code %d" % rc)
        return out

    def forward_simulate_traced(self, start, target, allow_contacts=True, noise_mode=0, particle_id=0, capacity=65536):
        """ForwardSimulateRobot with enable_tracing (spcs.hpp:824-829) for one particle -> (result record, trace records)."""
        start = _f64(start).reshape(-1)
        target = _f64(target).reshape(-1)
        width = (lib().oracle_trace_stride(self._h) - 16) // 8
        dt = np.dtype([("kind", np.uint32), ("step", np.uint32), ("microstep", np.uint32), ("iteration", np.uint32),
                       ("values", np.float64, (width,))])
        rec = np.zeros(capacity, dtype=dt)
        out = np.zeros(1, dtype=self.dtype)
        n = lib().oracle_forward_simulate_traced(self._h, start.ctypes.data, target.ctypes.data, int(bool(allow_contacts)), int(noise_mode),
                                                 None, int(particle_id), out.ctypes.data, rec.ctypes.data, capacity)
        assert n <= capacity
        return out, rec[:n]

    def check_config_collision(self, configs, inflation_ratio=0.0):
        configs = _f64(configs).reshape(-1, self.stride)
        out = np.zeros(configs.shape[0], dtype=np.uint8)
        lib().oracle_check_config_collision(self._h, configs.ctypes.data, configs.shape[0], float(inflation_ratio), out.ctypes.data)
        return out.astype(bool)

    def pairwise_config_distance(self, configs):
        """ComputeConfigurationDistanceTo (spcs.hpp:898) between every pair of configurations -> (n, n)."""
        configs = _f64(configs).reshape(-1, self.stride)
        out = np.zeros((configs.shape[0], configs.shape[0]))
        lib().oracle_pairwise_config_distance(self._h, configs.ctypes.data, configs.shape[0], out.ctypes.data)
        return out

    def statistics(self):
        out = np.zeros(11, dtype=np.uint64)
        lib().oracle_get_statistics(self._h, out.ctypes.data)
        return out

    def reset_statistics(self):
        lib().oracle_reset_statistics(self._h)

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_with_tape(oracle, starts, targets, allow_contacts=True, noise_mode=ORACLE_NOISE_MT19937, first_particle_id=0):
    """Run the oracle recording its truncated-normal draws and its solver decisions; returns
    (records, (draws, offsets, decision words, decision offsets), sensitivity)."""
    starts = _f64(starts).reshape(-1, oracle.stride)
    n = starts.shape[0]
    rec = oracle.forward_simulate(starts, targets, allow_contacts, noise_mode, None, first_particle_id, record_tape=True)
    total = int(lib().oracle_tape_total(oracle._h))
    draws = np.zeros(max(total, 1))
    offs = np.zeros(n + 1, dtype=np.uint64)
    lib().oracle_copy_tape(oracle._h, draws.ctypes.data, offs.ctypes.data)
    sens = np.zeros(n, dtype=np.uint32)
    lib().oracle_copy_sensitivity(oracle._h, sens.ctypes.data)
    nwords = int(lib().oracle_decision_words(oracle._h))
    dec = np.zeros(max(nwords, 1), dtype=np.uint64)
    dec_offs = np.zeros(n + 1, dtype=np.uint64)
    lib().oracle_copy_decisions(oracle._h, dec.ctypes.data, dec_offs.ctypes.data)
    return rec, (draws[:total], offs, dec[:nwords], dec_offs), sens


def decision_records(tape, n_dof):
    """Structured view of a decision tape: rank, flags, rows, order, solution per solve (include/fksgpu.h)."""
    words = np.asarray(tape[2], dtype=np.uint64).reshape(-1, 2 + n_dof)
    w0 = words[:, 0]
    return dict(rank=(w0 & np.uint64(0xFF)).astype(np.int64), flags=(w0 & np.uint64(0xFF00)).astype(np.int64),
                rows=((w0 >> np.uint64(16)) & np.uint64(0xFFFFFFFF)).astype(np.int64), order=words[:, 1],
                solution=words[:, 2:].view(np.float64), offsets=np.asarray(tape[3], dtype=np.int64))


def captured_systems(limit=None):
    """The stacked systems kept by oracle_debug_capture_systems: list of (A rows x cols, b)."""
    n = int(lib().oracle_debug_captured_count())
    out = []
    for i in range(n if limit is None else min(n, limit)):
        r, c = C.c_int(), C.c_int()
        lib().oracle_debug_captured_system(i, C.byref(r), C.byref(c), None, None)
        A = np.zeros((c.value, r.value))
        b = np.zeros(r.value)
        lib().oracle_debug_captured_system(i, C.byref(r), C.byref(c), A.ctypes.data, b.ctypes.data)
        out.append((A.T.copy(), b))
    return out


def qr_solve_info(A, b):
    """Eigen-semantics solver of the oracle on A (rows x cols), b -> (x, rank, order, flags)."""
    A = np.asarray(A, dtype=np.float64)
    Acm = np.ascontiguousarray(A.T)
    b = _f64(b)
    x = np.zeros(A.shape[1])
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_colpiv_qr_solve_info(Acm.ctypes.data, b.ctypes.data, A.shape[0], A.shape[1], x.ctypes.data, out.ctypes.data)
    return x, int(out[0]), int(out[2]), int(out[3])


def qr_device_model(A, b, opt_bits=3):
    """CPU model of the DEVICE solver (oracle/fks_qr_model.cpp); opt_bits: 1 = fma in norms / back substitution, 2 = fma in
    the reflector dots / updates.  -> (x, rank, order, min pivot^2 / cut)."""
    A = np.asarray(A, dtype=np.float64)
    Acm = np.ascontiguousarray(A.T)
    b = _f64(b)
    x = np.zeros(A.shape[1])
    out = np.zeros(4)
    lib().oracle_qr_device_model(Acm.ctypes.data, b.ctypes.data, A.shape[0], A.shape[1], int(opt_bits), x.ctypes.data, out.ctypes.data)
    return x, int(out[0]), int(out[2]), float(out[3])


def end_states_partition(flags):
    """Particles split by did_contact: (order, n_without_contact, n_with_contact); ascending ids inside each part."""
    flags = np.ascontiguousarray(flags, dtype=np.uint32)
    order = np.zeros(flags.shape[0], dtype=np.uint32)
    counts = np.zeros(2, dtype=np.uint64)
    lib().oracle_end_states_partition(flags.ctypes.data, flags.shape[0], order.ctypes.data, counts.ctypes.data)
    return order, int(counts[0]), int(counts[1])


def max_condition_of_last_call(oracle, n):
    """Per particle: the largest condition estimate (|R00| / min |Rkk|) among the stacked systems it solved."""
    out = np.ones(n)
    lib().oracle_copy_max_condition(oracle._h, out.ctypes.data)
    return out


def sensitivity_of_last_call(oracle, n):
    sens = np.zeros(n, dtype=np.uint32)
    lib().oracle_copy_sensitivity(oracle._h, sens.ctypes.data)
    return sens


def actuator_reference(velocity_limit, acceleration_limit, proportional_noise, minimum_noise, percent_variance, controls, draws):
    """Run the REFERENCE's own TruncatedNormalUncertainVelocityActuator (oracle/_ref/unc_ref, simple_uncertainty_models.hpp
    compiled against the stand-in of arc_utilities in oracle/shim) with injected draws.
    -> (noiseless values, noisy values, (mean, stddev, lower, upper) handed to its noise distribution)."""
    lines = ["%r %r %r %r %r %d" % (float(velocity_limit), float(acceleration_limit), float(proportional_noise), float(minimum_noise),
                                    float(percent_variance), len(controls))]
    lines += ["%r %r" % (float(c), float(d)) for c, d in zip(controls, draws)]
    out = subprocess.run([UNC_REF], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split()
    vals = np.array([float(x) for x in out[:2 * len(controls)]]).reshape(-1, 2)
    return vals[:, 0], vals[:, 1], tuple(float(x) for x in out[2 * len(controls):])


def actuator_oracle(velocity_limit, proportional_noise, minimum_noise, controls, draws):
    """The oracle's actuator arithmetic (Robot::actuate_axis) on the same pairs -> (noiseless, noisy)."""
    controls, draws = _f64(controls), _f64(draws)
    quiet, noisy = np.zeros(len(controls)), np.zeros(len(controls))
    lib().oracle_actuate_run(float(velocity_limit), float(proportional_noise), float(minimum_noise), controls.ctypes.data,
                             draws.ctypes.data, len(controls), quiet.ctypes.data, noisy.ctypes.data)
    return quiet, noisy


def pid_reference(kp, ki, kd, iclamp, errors, timesteps):
    """Run the REFERENCE's own simple_pid_controller.hpp (oracle/_ref/pid_ref)."""
    lines = ["%r %r %r %r %d" % (float(kp), float(ki), float(kd), float(iclamp), len(errors))]
    lines += ["%r %r" % (float(e), float(t)) for e, t in zip(errors, timesteps)]
    out = subprocess.run([PID_REF], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    return np.array([float(x) for x in out.split()])
