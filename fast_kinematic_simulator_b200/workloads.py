"""The five BASELINE.json configurations made concrete: environment obstacles, robot
description, seeded starts/targets.  Pure numpy; environments are built by the C library's host-side
restatement of BuildCompleteEnvironment.

Common parameters: SimulatorSolverParameters defaults (spcs.hpp:357-368), controller 25 Hz,
forward_simulation_time 1.0 s (25 steps), allow_contacts = true, actuator noise sigma fraction 0.5
(tnuva.hpp:128-130); seeds: geometry 1002, starts/targets 1003, simulator prng_seed 42.
"""
import numpy as np

from . import abi as capi
from .robots import RobotDescription, make_transform

CONTROLLER_HZ = 25.0
PRNG_SEED = 42


def _axis(vlim, kp=1.0, prop=0.1, minn=0.01):
    return dict(kp=kp, ki=0.0, kd=0.0, integral_clamp=0.0, velocity_limit=vlim, proportional_noise=prop,
                minimum_noise=minn, noise_sigma=0.5)


def _rot(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


class Workload:
    def __init__(self, name, kind, obstacles, resolution, robot, starts, targets, description):
        self.name = name
        self.kind = kind
        self.obstacles = obstacles
        self.resolution = resolution
        self.robot = robot
        self.starts = np.ascontiguousarray(starts, dtype=np.float64)
        self.targets = np.ascontiguousarray(targets, dtype=np.float64)
        self.description = description
        self._env = None

    @property
    def n_particles(self):
        return self.starts.shape[0]

    def environment(self):
        """BuildCompleteEnvironment by the product's host builder (libfksgpu.so)."""
        if self._env is None:
            from . import simulator as S

            self._env = S.build_complete_environment(self.obstacles, self.resolution)
        return self._env

    def device_environment(self, device=0):
        """BuildCompleteEnvironment on the GPU (fks_env_build_device): no host copy of the grids."""
        from . import simulator as S

        return S.build_complete_environment_on_device(self.obstacles, self.resolution, device)

    def make_simulator(self, device=0, solver_params=None, seed=PRNG_SEED, build_on_device=False):
        from . import simulator as S

        fn = {capi.ROBOT_SE2: S.make_se2_simulator, capi.ROBOT_SE3: S.make_se3_simulator,
              capi.ROBOT_LINKED: S.make_linked_simulator}[self.kind]
        env = self.device_environment(device) if build_on_device else self.environment()
        return fn(env, self.robot, solver_params, CONTROLLER_HZ, seed, 0, device)

    def subset(self, n, offset=0):
        t = self.targets if self.targets.shape[0] == 1 else self.targets[offset:offset + n]
        return self.starts[offset:offset + n], t


# ------------------------------------------------------------------------------------------------
# config 1: SE(2) robot, 128 particles x 1 action, builder-generated voxel SDF environment
# ------------------------------------------------------------------------------------------------
def se2_arena(n_particles=128):
    res = 0.125
    obstacles = [
        # z centre 0.05: keeps the robot's z = 0 plane off a voxel boundary
        (make_transform((0.0, -5.25, 0.05)), (5.5, 0.25, 0.5), 1),   # south wall
        (make_transform((0.0, 5.25, 0.05)), (5.5, 0.25, 0.5), 2),    # north wall
        (make_transform((-5.25, 0.0, 0.05)), (0.25, 5.0, 0.5), 3),   # west wall
        (make_transform((5.25, 0.0, 0.05)), (0.25, 5.0, 0.5), 4),    # east wall
        (make_transform((1.0, 0.0, 0.05)), (0.5, 1.0, 0.5), 5),      # interior box on the way to the target
        (make_transform((-2.0, 2.5, 0.05), _rot((0, 0, 1), 0.5)), (0.75, 0.5, 0.5), 6),  # rotated interior box
    ]
    xs = np.arange(-0.5, 0.5 + 1e-9, res)
    ys = np.arange(-0.25, 0.25 + 1e-9, res)
    pts = np.array([(x, y, 0.0) for x in xs for y in ys])
    robot = RobotDescription(capi.ROBOT_SE2, pts, np.zeros(len(pts), np.int32),
                             [_axis(1.0), _axis(1.0), _axis(0.5)])
    start = np.array([[-0.35, 0.3, 0.2]])
    starts = np.repeat(start, n_particles, axis=0)
    targets = np.array([[1.0, 0.1, -0.3]])  # inside the interior box: contact guaranteed
    return Workload("se2_arena", capi.ROBOT_SE2, obstacles, res, robot, starts, targets,
                    "SE(2) 1.0x0.5 lattice robot (P=%d), 10x10 arena, res 0.125, one start, one target inside a box" % len(pts))


# ------------------------------------------------------------------------------------------------
# config 2 / 4: SE(3) peg
# ------------------------------------------------------------------------------------------------
def _peg_points():
    xs = np.linspace(-0.125, 0.125, 6)
    ys = np.linspace(-0.25, 0.25, 11)
    zs = np.linspace(-0.125, 0.125, 6)
    return np.array([(x, y, z) for x in xs for y in ys for z in zs])


def _se3_robot():
    pts = _peg_points()
    axes = [_axis(1.0)] * 3 + [_axis(0.5)] * 3
    return RobotDescription(capi.ROBOT_SE3, pts, np.zeros(len(pts), np.int32), axes)


def _se3_config(t, rotvec):
    ang = np.linalg.norm(rotvec)
    R = np.eye(3) if ang < 1e-12 else _rot(rotvec / ang, ang)
    return make_transform(t, R)


def se3_narrow_passage(n_particles=16384, seed=1003):
    res = 0.05
    gap = 0.35  # peg width 0.25 + 2 cells
    obstacles = [
        (make_transform((-(1.0 + gap / 2), 0.0, 0.0)), (1.0, 1.0, 1.0), 1),
        (make_transform(((1.0 + gap / 2), 0.0, 0.0)), (1.0, 1.0, 1.0), 2),
    ]
    rng = np.random.Generator(np.random.MT19937(seed))
    t0 = np.array([0.0, -1.35, 0.0])
    starts = np.empty((n_particles, 12))
    for i in range(n_particles):
        starts[i] = _se3_config(t0 + rng.normal(0.0, 0.02, 3), rng.normal(0.0, 0.05, 3))
    targets = _se3_config(t0 + np.array([0.0, 0.9, 0.0]), np.zeros(3)).reshape(1, 12)
    return Workload("se3_narrow_passage", capi.ROBOT_SE3, obstacles, res, _se3_robot(), starts, targets,
                    "SE(3) peg 0.25x0.5x0.25 (P=396) through a 0.35 m slot between two 2 m blocks, res 0.05")


def se3_highres(n_particles=65536, seed=1003, n_cuboids=256, cube=10.1, res=0.02):
    """config 4: ~512^3 SDF.  8 small corner cuboids pin the bounding box; seeded random cuboids inside."""
    rng = np.random.Generator(np.random.MT19937(1001))
    h = cube / 2
    obstacles = []
    oid = 1
    for sx in (-1, 1):
        for sy in (-1, 1):
            for sz in (-1, 1):
                obstacles.append((make_transform((sx * (h - 0.05), sy * (h - 0.05), sz * (h - 0.05))), (0.05, 0.05, 0.05), oid))
                oid += 1
    centres, exts = [], []
    for _ in range(n_cuboids):
        ext = rng.uniform(0.1, 0.6, 3) * (cube / 10.1)
        c = rng.uniform(-h + 1.0 * (cube / 10.1), h - 1.0 * (cube / 10.1), 3)
        rv = rng.normal(0.0, 0.6, 3)
        obstacles.append((_se3_config(c, rv), tuple(ext), oid))
        centres.append(c)
        exts.append(np.linalg.norm(ext))
        oid += 1
    centres = np.array(centres)
    exts = np.array(exts)
    boxes = obstacles[8:]
    prng = np.random.Generator(np.random.MT19937(seed))
    starts = np.empty((n_particles, 12))
    targets = np.empty((n_particles, 12))
    i = 0
    while i < n_particles:
        rv = prng.normal(0.0, 0.5, 3)
        if i % 2 == 0:
            # free flight: anywhere clear of every cuboid's bounding sphere + peg radius, 0.3 m in a random direction
            c = prng.uniform(-h + 0.6, h - 0.6, 3)
            if np.any(np.linalg.norm(centres - c, axis=1) < exts + 0.33):
                continue
            d = prng.normal(0.0, 1.0, 3)
            d = 0.3 * d / np.linalg.norm(d)
        else:
            # approach: in front of a face of a random cuboid (clear of all the others), 0.3 m straight at the face --
            # the peg ends up pressed against it, so half of the batch resolves contacts against the HBM-resident SDF
            k = int(prng.integers(0, len(boxes)))
            T, ext = np.asarray(boxes[k][0]).reshape(3, 4), np.asarray(boxes[k][1])
            a, sgn = int(prng.integers(0, 3)), (1.0 if prng.random() < 0.5 else -1.0)
            normal = sgn * T[:, a]
            lateral = sum(prng.uniform(-0.5, 0.5) * ext[b] * T[:, b] for b in range(3) if b != a)
            c = T[:, 3] + normal * (ext[a] + 0.31 + prng.uniform(0.02, 0.12)) + lateral
            others = np.arange(len(boxes)) != k
            if np.any(np.abs(c) > h - 0.6) or np.any(np.linalg.norm(centres[others] - c, axis=1) < exts[others] + 0.33):
                continue
            d = -0.3 * normal
        starts[i] = _se3_config(c, rv)
        targets[i] = _se3_config(c + d, rv)
        i += 1
    return Workload("se3_highres", capi.ROBOT_SE3, obstacles, res, _se3_robot(), starts, targets,
                    "SE(3) peg (P=396) among %d random cuboids, %.2f m cube at res %.3f (HBM-resident SDF); every second particle is "
                    "driven 0.3 m straight at a cuboid face, the others 0.3 m in a random direction through free space" % (n_cuboids, cube, res))


# ------------------------------------------------------------------------------------------------
# config 3 / 5: 7-DoF serial arm
# ------------------------------------------------------------------------------------------------
ARM_LINK_LENGTH = 0.3
ARM_BASE_Z = 0.62


def _arm_link_points():
    pts = []
    for z in np.linspace(0.03, 0.27, 6):
        for k in range(8):
            a = 2 * np.pi * k / 8
            pts.append((0.04 * np.cos(a), 0.04 * np.sin(a), z))
    return np.array(pts)


def arm_robot():
    L = 8
    link_pts = _arm_link_points()
    pts = np.concatenate([link_pts for _ in range(L)])
    plink = np.repeat(np.arange(L, dtype=np.int32), len(link_pts))
    joints = []
    for j in range(7):
        joints.append(dict(parent=j, child=j + 1,
                           type=capi.JOINT_CONTINUOUS if j in (0, 6) else capi.JOINT_REVOLUTE,
                           transform=make_transform((0.0, 0.0, ARM_LINK_LENGTH)),
                           axis=(0.0, 0.0, 1.0) if j % 2 == 0 else (0.0, 1.0, 0.0),
                           lower=-2.9, upper=2.9, weight=1.0))
    allowed = np.zeros((L, L), np.uint8)
    for a in range(L):
        for b in range(L):
            if abs(a - b) <= 2:
                allowed[a, b] = 1
    axes = [_axis(1.0, kp=2.0, prop=0.1, minn=0.005) for _ in range(7)]
    base = make_transform((0.0, 0.0, ARM_BASE_Z))
    return RobotDescription(capi.ROBOT_LINKED, pts, plink, axes, n_links=L, joints=joints, base_transform=base,
                            allowed_self_collision=allowed)


def arm_fk_tip(q):
    """numpy FK of the arm's last-link tip (only used to place the target; not the oracle)."""
    T = np.eye(4)
    T[2, 3] = ARM_BASE_Z
    for j in range(7):
        A = np.eye(4)
        A[2, 3] = ARM_LINK_LENGTH
        R = np.eye(4)
        R[:3, :3] = _rot((0, 0, 1) if j % 2 == 0 else (0, 1, 0), q[j])
        T = T @ A @ R
    return (T @ np.array([0, 0, ARM_LINK_LENGTH, 1.0]))[:3]


def arm_room_obstacles():
    return [
        (make_transform((0.0, 0.0, -0.04)), (2.5, 2.5, 0.04), 1),      # floor
        (make_transform((0.0, 0.0, 2.46)), (2.5, 2.5, 0.04), 2),       # ceiling
        (make_transform((-2.46, 0.0, 1.21)), (0.04, 2.5, 1.21), 3),    # walls
        (make_transform((2.46, 0.0, 1.21)), (0.04, 2.5, 1.21), 4),
        (make_transform((0.0, -2.46, 1.21)), (2.42, 0.04, 1.21), 5),
        (make_transform((0.0, 2.46, 1.21)), (2.42, 0.04, 1.21), 6),
        (make_transform((0.9, 0.0, 0.72)), (0.5, 0.6, 0.04), 7),       # table slab, top at z = 0.76
    ]


# start: last link hovering ~7 cm above the table; target: its tip ~12 cm below the table top
ARM_START = np.array([0.0, 0.9, 0.0, 1.1, 0.0, 0.6, 0.0])
ARM_TARGET = np.array([0.2, 1.0, 0.1, 1.1, -0.1, 1.0, 0.3])


def arm_table(n_particles=65536, seed=1003):
    res = 0.04
    rng = np.random.Generator(np.random.MT19937(seed))
    starts = ARM_START[None, :] + rng.normal(0.0, 0.02, (n_particles, 7))
    targets = ARM_TARGET.reshape(1, 7)
    return Workload("arm_table", capi.ROBOT_LINKED, arm_room_obstacles(), res, arm_robot(), starts, targets,
                    "7-DoF serial arm (8 links x 48 points = 384), 5x5x2.5 m room + table, res 0.04; "
                    "target drives the last link into the table")


def arm_free(n_particles=256, seed=1003):
    """Raised arm moving through free space (no contact anywhere): FK, env check, self-collision broad phase."""
    res = 0.04
    rng = np.random.Generator(np.random.MT19937(seed))
    start = np.array([0.0, 0.5, 0.0, 0.9, 0.0, 0.5, 0.0])
    target = np.array([0.4, 0.6, 0.3, 1.0, 0.2, 0.6, 0.5])
    starts = start[None, :] + rng.normal(0.0, 0.02, (n_particles, 7))
    return Workload("arm_free", capi.ROBOT_LINKED, arm_room_obstacles(), res, arm_robot(), starts, target.reshape(1, 7),
                    "7-DoF arm moving through free space")


def arm_selfcollision(n_particles=256, seed=1003):
    """A folded arm whose distal links sweep through the proximal ones: exercises the self-collision path."""
    res = 0.04
    rng = np.random.Generator(np.random.MT19937(seed))
    start = np.array([0.0, 1.2, 0.0, 2.2, 0.0, 2.0, 0.0])
    target = np.array([0.0, 1.3, 0.0, 2.6, 0.0, 2.6, 0.0])
    starts = start[None, :] + rng.normal(0.0, 0.02, (n_particles, 7))
    return Workload("arm_selfcollision", capi.ROBOT_LINKED, arm_room_obstacles(), res, arm_robot(), starts,
                    target.reshape(1, 7), "7-DoF arm folding onto itself (self-collision resolver path)")


def arm_elbow(n_particles=256, seed=1003):
    """The arm straightens up under an overhead beam: links 3/4 (around the elbow) make the contact, so the
    stacked Jacobian has structurally-zero columns for the distal joints and full column rank on the rest --
    unlike `arm_table`, where a single distal link touching gives a rank-6 system in 7 unknowns and the
    reference's QR pivot is decided by round-off."""
    res = 0.04
    rng = np.random.Generator(np.random.MT19937(seed))
    obstacles = arm_room_obstacles() + [(make_transform((0.35, 0.0, 1.78)), (0.12, 0.3, 0.06), 8)]  # beam, bottom at z = 1.72
    start = np.array([0.0, 0.9, 0.0, 2.0, 0.0, -1.5, 0.0])  # forearm folded down, hand forward above the table
    starts = start[None, :] + rng.normal(0.0, 0.02, (n_particles, 7))
    target = np.array([0.1, 0.3, 0.0, 2.0, 0.0, -1.5, 0.0])
    return Workload("arm_elbow", capi.ROBOT_LINKED, obstacles, res, arm_robot(), starts, target.reshape(1, 7),
                    "7-DoF arm whose elbow (links 3/4) rises into an overhead beam")


WORKLOADS = {
    "se2_arena": se2_arena,
    "se3_narrow_passage": se3_narrow_passage,
    "arm_table": arm_table,
    "se3_highres": se3_highres,
    "arm_selfcollision": arm_selfcollision,
    "arm_elbow": arm_elbow,
    "arm_free": arm_free,
}


def make(name, **kw):
    return WORKLOADS[name](**kw)


def gantry(n_particles=128, seed=1003):
    """A small linked robot with every joint type the reference's linked model supports: two PRISMATIC axes (x, y), a FIXED
    mounting joint, and a REVOLUTE + CONTINUOUS wrist carrying a bar of points.  Exercises the prismatic / fixed FK and Jacobian
    paths that the 7-DoF arm does not."""
    res = 0.05
    rng = np.random.Generator(np.random.MT19937(seed))
    obstacles = [
        (make_transform((0.0, 0.0, -0.05)), (1.5, 1.5, 0.05), 1),        # floor
        (make_transform((0.9, 0.0, 0.3)), (0.1, 0.6, 0.3), 2),           # a wall the bar is pushed into
    ]
    L = 6
    bar = np.array([(0.0, 0.0, z) for z in np.linspace(0.02, 0.3, 8)] + [(x, 0.0, 0.3) for x in np.linspace(-0.15, 0.15, 7)])
    cube = np.array([(x, y, z) for x in (-0.04, 0.04) for y in (-0.04, 0.04) for z in (0.0, 0.08)])
    pts, plink = [], []
    for l in range(L):
        p = bar if l == L - 1 else cube
        pts.append(p)
        plink += [l] * len(p)
    pts = np.concatenate(pts)
    joints = [
        dict(parent=0, child=1, type=capi.JOINT_PRISMATIC, transform=make_transform((0.0, 0.0, 0.25)), axis=(1.0, 0.0, 0.0), lower=-1.0, upper=1.0),
        dict(parent=1, child=2, type=capi.JOINT_PRISMATIC, transform=make_transform((0.0, 0.0, 0.1)), axis=(0.0, 1.0, 0.0), lower=-0.2, upper=0.2),
        dict(parent=2, child=3, type=capi.JOINT_FIXED, transform=make_transform((0.0, 0.0, 0.1), _rot((1, 0, 0), 0.1)), axis=(0.0, 0.0, 1.0), lower=0.0, upper=0.0),
        dict(parent=3, child=4, type=capi.JOINT_REVOLUTE, transform=make_transform((0.0, 0.0, 0.1)), axis=(0.0, 1.0, 0.0), lower=-1.0, upper=1.0),
        dict(parent=4, child=5, type=capi.JOINT_CONTINUOUS, transform=make_transform((0.0, 0.0, 0.1)), axis=(0.0, 0.0, 1.0), lower=-np.pi, upper=np.pi),
    ]
    allowed = np.ones((L, L), np.uint8)
    axes = [_axis(1.0, kp=2.0), _axis(1.0, kp=2.0), _axis(1.0, kp=2.0), _axis(2.0, kp=2.0)]
    robot = RobotDescription(capi.ROBOT_LINKED, pts, np.array(plink, np.int32), axes, n_links=L, joints=joints,
                             base_transform=make_transform((0.011, 0.007, 0.063)), allowed_self_collision=allowed)  # off the voxel lattice
    start = np.array([0.0, 0.0, 0.2, 3.0])
    target = np.array([0.9, 0.3, 0.9, -2.9])   # x drives the bar into the wall, y runs into its upper limit, the wrist wraps through pi
    starts = start[None, :] + rng.normal(0.0, 0.02, (n_particles, 4))
    return Workload("gantry", capi.ROBOT_LINKED, obstacles, res, robot, starts, target.reshape(1, 4),
                    "linked robot with prismatic, fixed, revolute and continuous joints pushed into a wall")


WORKLOADS["gantry"] = gantry
