"""ctypes binding of include/fksgpu.h (the C ABI of libfksgpu.so).  Structures mirror the header
field for field; nothing here computes."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # FKSGPU_LIBRARY: developer override used to A/B kernel builds (always an in-tree libfksgpu*.so)
    return os.environ.get("FKSGPU_LIBRARY", os.path.join(_HERE, "libfksgpu.so"))


class FksError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("fksgpu error %d: %s" % (code, message))
        self.code = code


if not os.path.exists(library_path()):
    raise ImportError(
        "libfksgpu.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C fast_kinematic_simulator_b200/csrc` (there is no CPU fallback)"
    )
lib = C.CDLL(library_path())

OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_OUT_OF_MEMORY = range(6)
ROBOT_SE2, ROBOT_SE3, ROBOT_LINKED = 0, 1, 2
JOINT_PRISMATIC, JOINT_REVOLUTE, JOINT_CONTINUOUS, JOINT_FIXED = 0, 1, 2, 3
NOISE_PHILOX, NOISE_INJECTED, NOISE_NONE = 0, 1, 2
FLAG_DID_CONTACT = 1 << 0
FLAG_RESOLVE_FAILED = 1 << 1
FLAG_ENDED_BY_FAILURE = 1 << 2
FLAG_ENDED_BY_NOCONTACT = 1 << 3
FLAG_ENDED_BY_SHORTCUT = 1 << 4
FLAG_WOULD_ASSERT_MICROSTEP = 1 << 8
FLAG_WOULD_ASSERT_NORMAL = 1 << 9
FLAG_WOULD_ASSERT_NAN = 1 << 10
FLAG_EMPTY_JACOBIAN = 1 << 11
FLAG_TAPE_EXHAUSTED = 1 << 12
FLAG_NEAR_RANK_CUT = 1 << 13
FLAG_J_SPILLED = 1 << 14
FLAG_DECISION_OVERRIDDEN = 1 << 15
FLAG_DECISION_DESYNC = 1 << 16
TRACE_CONTROL_INPUT, TRACE_CONTROL_INPUT_STEP, TRACE_POST_ACTION, TRACE_RESOLUTION_STEP, TRACE_RETURNED_PREVIOUS = range(5)
NUM_STATS = 11
STAT_NAMES = (
    "successful_resolves",
    "unsuccessful_resolves",
    "free_resolves",
    "collision_resolves",
    "fallback_resolves",
    "unsuccessful_self_collision_resolves",
    "unsuccessful_env_collision_resolves",
    "recovered_unsuccessful_resolves",
    "total_microsteps",
    "total_resolver_iterations",
    "total_corrected_points",
)


class SolverParams(C.Structure):
    _fields_ = [
        ("forward_simulation_time", C.c_double),
        ("simulation_shortcut_distance", C.c_double),
        ("environment_collision_check_tolerance", C.c_double),
        ("resolve_correction_step_scaling_decay_rate", C.c_double),
        ("resolve_correction_initial_step_size", C.c_double),
        ("resolve_correction_min_step_scaling", C.c_double),
        ("max_resolver_iterations", C.c_uint32),
        ("resolve_correction_step_scaling_decay_iterations", C.c_uint32),
        ("failed_resolves_end_motion", C.c_int32),
        ("_pad", C.c_int32),
    ]


class EnvDesc(C.Structure):
    _fields_ = [
        ("origin", C.c_double * 12),
        ("inverse_origin", C.c_double * 12),
        ("map_resolution", C.c_double),
        ("sdf_resolution", C.c_double),
        ("nx", C.c_int64),
        ("ny", C.c_int64),
        ("nz", C.c_int64),
        ("sdf", C.POINTER(C.c_float)),
        ("oob_value", C.c_float),
        ("_pad", C.c_int32),
        ("n_normal_cells", C.c_int64),
        ("normal_cell_index", C.POINTER(C.c_int64)),
        ("normal_cell_start", C.POINTER(C.c_uint32)),
        ("normal_entries", C.POINTER(C.c_double)),
    ]


class AxisParams(C.Structure):
    _fields_ = [
        ("kp", C.c_double),
        ("ki", C.c_double),
        ("kd", C.c_double),
        ("integral_clamp", C.c_double),
        ("velocity_limit", C.c_double),
        ("proportional_noise", C.c_double),
        ("minimum_noise", C.c_double),
        ("noise_sigma", C.c_double),
    ]


class JointDesc(C.Structure):
    _fields_ = [
        ("parent_link", C.c_int32),
        ("child_link", C.c_int32),
        ("type", C.c_int32),
        ("_pad", C.c_int32),
        ("transform", C.c_double * 12),
        ("axis", C.c_double * 3),
        ("lower_limit", C.c_double),
        ("upper_limit", C.c_double),
        ("distance_weight", C.c_double),
    ]


class RobotDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_links", C.c_int32),
        ("n_joints", C.c_int32),
        ("n_dof", C.c_int32),
        ("n_points", C.c_int64),
        ("points_xyz", C.POINTER(C.c_double)),
        ("point_link", C.POINTER(C.c_int32)),
        ("axes", C.POINTER(AxisParams)),
        ("base_transform", C.c_double * 12),
        ("joints", C.POINTER(JointDesc)),
        ("allowed_self_collision", C.POINTER(C.c_uint8)),
        ("position_distance_weight", C.c_double),
        ("rotation_distance_weight", C.c_double),
    ]


class NoiseTape(C.Structure):
    _fields_ = [("draws", C.POINTER(C.c_double)), ("offsets", C.POINTER(C.c_uint64)),
                ("decisions", C.POINTER(C.c_uint64)), ("decision_offsets", C.POINTER(C.c_uint64))]


class Obstacle(C.Structure):
    _fields_ = [
        ("pose", C.c_double * 12),
        ("extents", C.c_double * 3),
        ("object_id", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


# every symbol include/fksgpu.h declares (tests/test_capi_symbols.py checks the list against the header)
EXPORTS = (
    "fks_abi_version",
    "fks_last_error_string",
    "fks_device_count",
    "fks_default_solver_params",
    "fks_env_create",
    "fks_env_destroy",
    "fks_robot_create",
    "fks_robot_destroy",
    "fks_robot_config_stride",
    "fks_sim_create",
    "fks_sim_destroy",
    "fks_sim_result_stride",
    "fks_forward_simulate",
    "fks_reverse_simulate",
    "fks_forward_simulate_device",
    "fks_check_config_collision",
    "fks_sim_trace_stride",
    "fks_forward_simulate_traced",
    "fks_get_statistics",
    "fks_reset_statistics",
    "fks_sim_launch_count",
    "fks_sim_kernel_info",
    "fks_build_environment",
    "fks_built_env_desc",
    "fks_built_env_occupancy",
    "fks_built_env_destroy",
    "fks_env_build_device",
    "fks_env_build_timings",
    "fks_env_download",
    "fks_debug_qr_solve",
    "fks_measure_fp64_peak",
    "fks_measure_gather_rate",
)

P = C.POINTER
lib.fks_abi_version.restype = C.c_int
lib.fks_last_error_string.restype = C.c_char_p
lib.fks_device_count.argtypes = [P(C.c_int)]
lib.fks_default_solver_params.argtypes = [P(SolverParams)]
lib.fks_default_solver_params.restype = None
lib.fks_env_create.argtypes = [C.c_int, P(EnvDesc), P(C.c_void_p)]
lib.fks_env_destroy.argtypes = [C.c_void_p]
lib.fks_env_destroy.restype = None
lib.fks_robot_create.argtypes = [C.c_int, P(RobotDesc), P(C.c_void_p)]
lib.fks_robot_destroy.argtypes = [C.c_void_p]
lib.fks_robot_destroy.restype = None
lib.fks_robot_config_stride.argtypes = [C.c_void_p]
lib.fks_sim_create.argtypes = [C.c_void_p, C.c_void_p, P(SolverParams), C.c_double, C.c_uint64, C.c_int32, P(C.c_void_p)]
lib.fks_sim_destroy.argtypes = [C.c_void_p]
lib.fks_sim_destroy.restype = None
lib.fks_sim_result_stride.argtypes = [C.c_void_p]
lib.fks_sim_result_stride.restype = C.c_size_t
_sim_args = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, P(NoiseTape), C.c_uint64, C.c_void_p]
lib.fks_forward_simulate.argtypes = _sim_args
lib.fks_reverse_simulate.argtypes = _sim_args
lib.fks_forward_simulate_device.argtypes = [
    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
    C.c_uint64, C.c_void_p, C.c_void_p,
]
lib.fks_check_config_collision.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
lib.fks_sim_trace_stride.argtypes = [C.c_void_p]
lib.fks_sim_trace_stride.restype = C.c_size_t
lib.fks_forward_simulate_traced.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, P(NoiseTape), C.c_uint64, C.c_void_p,
                                            C.c_void_p, C.c_size_t, P(C.c_size_t)]
lib.fks_get_statistics.argtypes = [C.c_void_p, P(C.c_uint64)]
lib.fks_reset_statistics.argtypes = [C.c_void_p]
lib.fks_sim_launch_count.argtypes = [C.c_void_p]
lib.fks_sim_launch_count.restype = C.c_uint64
lib.fks_sim_kernel_info.argtypes = [C.c_void_p]
lib.fks_sim_kernel_info.restype = C.c_char_p
lib.fks_build_environment.argtypes = [P(Obstacle), C.c_size_t, C.c_double, P(C.c_void_p)]
lib.fks_built_env_desc.argtypes = [C.c_void_p]
lib.fks_built_env_desc.restype = P(EnvDesc)
lib.fks_built_env_occupancy.argtypes = [C.c_void_p]
lib.fks_built_env_occupancy.restype = P(C.c_uint8)
lib.fks_built_env_destroy.argtypes = [C.c_void_p]
lib.fks_built_env_destroy.restype = None
lib.fks_env_build_device.argtypes = [C.c_int, P(Obstacle), C.c_size_t, C.c_double, P(C.c_void_p)]
lib.fks_env_build_timings.argtypes = [C.c_void_p, P(C.c_double), C.c_int]
lib.fks_env_download.argtypes = [C.c_void_p, P(C.c_void_p)]
lib.fks_debug_qr_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t, C.c_void_p, C.c_void_p]
lib.fks_measure_fp64_peak.argtypes = [C.c_int, P(C.c_double)]
lib.fks_measure_gather_rate.argtypes = [C.c_int, C.c_size_t, P(C.c_double)]


def check(code):
    if code != OK:
        msg = lib.fks_last_error_string()
        raise FksError(code, msg.decode() if msg else "")


def default_solver_params():
    """GetDefaultSolverParameters (fast_kinematic_simulator.hpp:13-16)."""
    p = SolverParams()
    lib.fks_default_solver_params(C.byref(p))
    return p
