"""Flat descriptions of the reference's robots (tnuva_robot_models.hpp:26,201,415) -> fks_robot_desc.  Pure Python + ctypes
structures (fast_kinematic_simulator_b200.abi): no library is loaded here, so the CPU oracle's harness can use them too."""
import ctypes as C

import numpy as np

from . import abi as capi


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
