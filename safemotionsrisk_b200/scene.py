"""Scene builder: env_config + compiled URDF/mesh assets -> ``SmScene`` (include/smenv.h).

Restates the scene-assembly logic of the reference on the host, once per env construction:
  * limits: robot_scene_base.py:20-22, :347-371, :441-455
  * kinematic chain: description/urdf/robot*.urdf via tools/compile_assets.py
  * which link is checked against which obstacle: robot_scene_base.py:226-230, :288-292,
    collision_torque_limit_prevention.py (ctlp.py) :577-589, :1409-1459, :2416-2434
  * planet orbit tables: ctlp.py:4403-4439 (pure NumPy in the reference, reproduced with the same NumPy calls)
  * distance caps / reward weights: ctlp.py:354-366, rewards.py:81-93, :95-162
  * ball observation ranges: ctlp.py:303-348
Bullet collision-shape conventions (margins, box core shrink, contact-breaking threshold) follow SURVEY.md
Appendix B and are marked as such.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .config import EnvConfig

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "scene_assets.npz")

JOINT_LIMITS_SAFETY_BUFFER_IIWA = 0.035                       # robot_scene_base.py:20
MAX_ACCELERATION_IIWA = [15.0, 7.5, 10.0, 12.5, 15.0, 20.0, 20.0]   # robot_scene_base.py:21
MAX_JERK_IIWA = [7500, 3750, 5000, 6250, 7500, 10000, 10000]        # robot_scene_base.py:22
URDF_MARGIN = 0.001           # Bullet gUrdfDefaultCollisionMargin
CONTACT_BREAKING_FACTOR = 0.02  # Bullet gContactBreakingThreshold
SIM_TIME_STEP = 1.0 / 240.0   # safe_motions_base.py:34


def rpy_to_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def matrix_to_quat(m):
    """xyzw quaternion of a rotation matrix."""
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    return np.array(q)


class _Body:
    """One compiled URDF: link tree + convex parts in link frames."""

    def __init__(self, assets, prefix, scale=1.0):
        g = lambda k: assets[prefix + "/" + k]
        self.link_names = [str(n) for n in g("link_names")]
        self.parent = g("parent")
        self.joint_type = g("joint_type")
        self.joint_axis = g("joint_axis")
        self.joint_xyz = g("joint_xyz") * scale
        self.joint_rpy = g("joint_rpy")
        self.joint_limit = g("joint_limit")
        self.inertial_xyz = g("inertial_xyz") * scale
        self.part_link = g("part_link")
        self.part_kind = [str(k) for k in g("part_kind")]
        start = g("part_start")
        verts = g("part_verts") * scale  # globalScaling scales vertices and origins, not margins (Appendix B.1)
        self.parts = [verts[start[i]:start[i + 1]] for i in range(len(start) - 1)]

    def link_index(self, name):
        return self.link_names.index(name)


def planet_tables(center, radius_xy, euler_angles, period, update_time_step, default_euler, rotations_per_period,
                  orbit_interpolation_steps=1000):
    """Orbit pose table of one planet, as Planet.__init__ builds it (ctlp.py:4397-4433)."""
    center = np.asarray(center, dtype=np.float64)
    direction_sign = 1
    if period < 0:
        period, direction_sign = -period, -1
    theta = np.linspace(0, direction_sign * 2 * np.pi, orbit_interpolation_steps)
    local = np.array([radius_xy[0] * np.cos(theta), radius_xy[1] * np.sin(theta), np.zeros_like(theta)]).T
    r_shift = rpy_to_matrix(euler_angles)
    glob = center + local @ r_shift.T  # p.multiplyTransforms(center, orn_shift, local_pos, identity)
    seg = np.linalg.norm(np.diff(local, axis=0), axis=1)
    cum = np.concatenate(([0.0], np.cumsum(seg)))
    total = cum[-1]
    t_steps = np.arange(0, period, update_time_step)
    idx = np.searchsorted(cum, t_steps / period * total)
    pos, local_pos = glob[idx], local[idx]
    r_default = rpy_to_matrix(default_euler)
    quat = np.zeros((len(t_steps), 4))
    for i, t in enumerate(t_steps):
        ang = direction_sign * t / period * 2 * np.pi * rotations_per_period
        rot = r_shift @ r_default @ rpy_to_matrix([0, 0, ang]) if rotations_per_period else r_default
        quat[i] = matrix_to_quat(rot)
    return np.ascontiguousarray(pos), np.ascontiguousarray(quat), np.ascontiguousarray(local_pos[:, :2]), total


def ball_target_height_time(initial_height, initial_z_speed, target_height):
    """Ball.get_target_height_time (ctlp.py:4346-4361), time only."""
    g = 9.81
    sqrt_value = initial_z_speed ** 2 + 2 * g * (initial_height - target_height)
    if sqrt_value >= 0:
        hit_time = (initial_z_speed + np.sqrt(sqrt_value)) / g
        if hit_time > 0:
            return hit_time
    return np.nan


def merge_chain(body):
    """Merges the fixed joints of a compiled URDF into the frame of the nearest revolute ancestor.  Returns the revolute
    joints (parent frame, fixed transform in front of the rotation, axis, limits), per link the frame that moves it
    (0 = the body's root, 1 + j = child link of revolute joint j) and its fixed pose in that frame."""
    n_links = len(body.link_names)
    link_frame = np.zeros(n_links, dtype=np.int32)
    link_X = [(np.eye(3), np.zeros(3)) for _ in range(n_links)]
    joints = []
    for i in range(1, n_links):
        p = body.parent[i]
        r_loc, t_loc = rpy_to_matrix(body.joint_rpy[i]), body.joint_xyz[i]
        rp, tp = link_X[p]
        r_abs, t_abs = rp @ r_loc, rp @ t_loc + tp
        if body.joint_type[i] == 0:
            link_frame[i] = link_frame[p]
            link_X[i] = (r_abs, t_abs)
        elif body.joint_type[i] == 1:
            joints.append(dict(link=i, parent_frame=int(link_frame[p]), R=r_abs, t=t_abs,
                               axis=body.joint_axis[i] / np.linalg.norm(body.joint_axis[i]),
                               limit=body.joint_limit[i]))
            link_frame[i] = len(joints)
        else:
            raise NotImplementedError("joint type of link " + body.link_names[i])
    return joints, link_frame, link_X


# human_network/params.json (the nested env's env_config; the keys that shape the path) and robot_scene_base.py:24-25
HUMAN_ENV = dict(closest_point_safety_distance=0.01, collision_check_time=0.033, check_braking_trajectory_collisions=True,
                 obstacle_scene=1, target_point_radius=0.065, target_point_sequence=1, target_link_offset=[0.0, 0.0, -0.185],
                 trajectory_time_step=0.1, pos_limit_factor=1.0, vel_limit_factor=1.0, acc_limit_factor=1.0,
                 jerk_limit_factor=1.0, log_std_range=[-1.375, 0.0])
MAX_ACCELERATION_HUMAN_ARM = [15.0, 15.0, 15.0, 15.0]     # robot_scene_base.py:24
MAX_JERK_HUMAN_ARM = [7500, 7500, 7500, 7500]             # robot_scene_base.py:25
HUMAN_BASE_POSITION = [0.7, 0, -0.94]                     # robot_scene_base.py:177
HUMAN_BASE_EULER = [0, 0, np.pi]                          # robot_scene_base.py:178
BRAKING_TIMEOUT = 2.0                                     # ctlp.py:3467-3468


class Scene:
    """Holds the ctypes ``SmScene`` plus the NumPy arrays it points into."""

    def __init__(self, config=None, **kwargs):
        self.config = config if isinstance(config, EnvConfig) else EnvConfig(**(config or {}), **kwargs)
        cfg = self.config
        assets = np.load(ASSETS)
        self._keep = []  # arrays referenced by pointers inside the struct
        sc = abi.SmScene()
        self.struct = sc
        robot = _Body(assets, "robot_ball_machine" if cfg.ball_machine_mode else "robot")
        self.robot = robot

        # ---------------- kinematic chain: merge fixed joints into the frame of the nearest revolute ancestor
        n_links = len(robot.link_names)
        joints, link_frame, link_X = merge_chain(robot)
        nj = len(joints)
        assert nj <= abi.SM_MAX_JOINTS
        self.n_joints = nj
        self.link_frame, self.link_X = link_frame, link_X
        sc.n_joints = nj
        for j, jt in enumerate(joints):
            sc.joint_parent[j] = jt["parent_frame"]
            sc.joint_R[j][:] = list(jt["R"].reshape(-1))
            sc.joint_t[j][:] = list(jt["t"])
            sc.joint_axis[j][:] = list(jt["axis"])

        # ---------------- limits (robot_scene_base.py:347-371, :441-455)
        ts = float(cfg.trajectory_time_step)
        lower = np.array([jt["limit"][0] + JOINT_LIMITS_SAFETY_BUFFER_IIWA for jt in joints]) * cfg.pos_limit_factor
        upper = np.array([jt["limit"][1] - JOINT_LIMITS_SAFETY_BUFFER_IIWA for jt in joints]) * cfg.pos_limit_factor
        vel = np.array([jt["limit"][3] for jt in joints]) * cfg.vel_limit_factor
        acc = np.array(MAX_ACCELERATION_IIWA) * cfg.acc_limit_factor
        jerk = np.array([min(2 * acc[i] / ts, MAX_JERK_IIWA[i]) * cfg.jerk_limit_factor for i in range(nj)])
        self.pos_lo, self.pos_hi, self.vel_max, self.acc_max, self.jerk_max = lower, upper, vel, acc, jerk
        for j in range(nj):
            sc.pos_lo[j], sc.pos_hi[j] = lower[j], upper[j]
            sc.vel_max[j], sc.acc_max[j], sc.jerk_max[j] = vel[j], acc[j], jerk[j]
        sc.ts = ts
        sc.substeps = int(round(ts / SIM_TIME_STEP))  # safe_motions_base.py:754-756
        sc.limit_velocity, sc.limit_position = int(bool(cfg.limit_velocity)), int(bool(cfg.limit_position))
        sc.action_mapping_factor = float(cfg.action_mapping_factor)
        sc.track_kp = 0.1                                                     # PyBullet default positionGain
        sc.track_vel = 0.87 if cfg.use_controller_target_velocities else 0.0  # robot_scene_base.py:792, :802-805
        sc.contact_stride = int(cfg.contact_check_stride)

        # ---------------- shapes
        shapes, verts = [], []

        def add_shape(v, frame, link, margin=URDF_MARGIN):
            v = np.asarray(v, dtype=np.float64)
            centre = 0.5 * (v.min(0) + v.max(0))
            sh = dict(frame=frame, off=sum(len(x) for x in verts), cnt=len(v), link=link, margin=margin,
                      center=centre, radius=float(np.linalg.norm(v - centre, axis=1).max()), verts=v)
            shapes.append(sh)
            verts.append(v)
            return len(shapes) - 1

        link_shapes = {}  # link name -> shape ids
        link_disc = {}    # link name -> Bullet angular motion disc of the link's collision object
        for pi, v in enumerate(robot.parts):
            li = int(robot.part_link[pi])
            if robot.part_kind[pi] == "point":
                continue
            r_x, t_x = link_X[li]
            sid = add_shape(v @ r_x.T + t_x, int(link_frame[li]), li)
            link_shapes.setdefault(robot.link_names[li], []).append(sid)
        for name, sids in link_shapes.items():
            li = robot.link_index(name)
            link_disc[name] = self._angular_motion_disc(
                [robot.parts[pi] for pi in range(len(robot.parts)) if robot.part_link[pi] == li
                 and robot.part_kind[pi] != "point"], robot.inertial_xyz[li])
        self.link_shapes = link_shapes
        target_link = cfg.target_link_name or ("ball_machine" if cfg.ball_machine_mode else "iiwa_link_7")
        self.target_link = target_link

        # static obstacle: table (ctlp.py:2426-2434); only its observed-link variant obstacle_scene == 5 is built
        if cfg.obstacle_scene not in (0, 5):
            raise NotImplementedError("obstacle_scene {} (virtual walls / planes) is not implemented".format(
                cfg.obstacle_scene))
        closest_point_active = ["iiwa_link_2", "iiwa_link_3", "iiwa_link_4", "iiwa_link_5", "iiwa_link_6",
                                "iiwa_link_7"] + ([target_link] if cfg.ball_machine_mode else [])  # ctlp.py:577-589
        static_pairs = []
        if cfg.obstacle_scene == 5:
            table = _Body(assets, "obstacle_table")
            corners = table.parts[0]
            centre = corners.mean(0)
            half = np.abs(corners - centre).max(0)
            core = centre + np.sign(corners - centre) * (half - URDF_MARGIN)  # btBoxShape core (Appendix B.1)
            table_sid = add_shape(core, 0, -1)
            observed = ["iiwa_link_1", "iiwa_link_2", "iiwa_link_3", "iiwa_link_4", "iiwa_link_5", "iiwa_link_6",
                        "iiwa_link_7"] + ([target_link] if cfg.ball_machine_mode else [])   # ctlp.py:2419-2424
            for name in observed:
                if name in closest_point_active:                                            # ctlp.py:3297-3298
                    static_pairs += [(sid, table_sid) for sid in link_shapes[name]]
        # self collision (ctlp.py:1438-1446): only the ball machine against the lower links
        self_pairs = []
        if cfg.ball_machine_mode:
            for name in ["iiwa_base_adapter", "iiwa_link_0", "iiwa_link_1", "iiwa_link_2", "iiwa_link_3",
                         "iiwa_link_4", "iiwa_link_5"]:
                self_pairs += [(a, b) for a in link_shapes[target_link] for b in link_shapes[name]]

        compute_static_self = (cfg.collision_avoidance_self_collision_max_reward != 0 or
                               cfg.collision_avoidance_static_obstacles_max_reward != 0 or
                               cfg.terminate_on_self_collision or cfg.terminate_on_collision_with_static_obstacle)
        if not compute_static_self:  # rewards.py:105-107
            static_pairs, self_pairs = [], []

        # ---------------- moving obstacles
        obstacles = []
        upd = ts / sc.substeps  # ctlp.py:97
        human = None
        if cfg.human_network_checkpoint is not None:
            if cfg.planet_mode or cfg.use_moving_objects:
                raise NotImplementedError("a human together with planets or balls is not implemented")
            if os.path.basename(os.path.dirname(os.path.dirname(str(cfg.human_network_checkpoint)))) != "human_network":
                raise NotImplementedError("human_network_checkpoint: only the shipped human_network/checkpoint/checkpoint "
                                          "(its weights are packaged in assets/networks_human.npz) is supported")
            human = _Body(assets, "human")
            observed = ["iiwa_link_2", "iiwa_link_3", "iiwa_link_4", "iiwa_link_5", "iiwa_link_6"] + \
                ([target_link] if cfg.ball_machine_mode else ["iiwa_link_7"])     # ctlp.py:788-794
            obstacles.append(dict(kind=abi.SM_OBST_HUMAN, body=human, observed=observed))
        if cfg.planet_mode:
            iss = _Body(assets, "obstacle_ISS", scale=0.6)       # ctlp.py:736-737
            ast = _Body(assets, "obstacle_asteroid", scale=1.0)  # ctlp.py:762-763
            p1 = dict(center=cfg.planet_one_center,
                      radius_xy=cfg.planet_one_radius_xy if cfg.planet_one_radius_xy is not None else [0.65, 0.8],
                      euler=cfg.planet_one_euler_angles if cfg.planet_one_euler_angles is not None else [0.35, 0, 0],
                      period=cfg.planet_one_period if cfg.planet_one_period is not None else 5.0)
            if p1["center"] is None or cfg.planet_two_center is None:
                raise NotImplementedError("planet_mode needs planet_one_center and planet_two_center")
            time_shift = cfg.planet_two_time_shift
            p2 = dict(center=cfg.planet_two_center,
                      radius_xy=cfg.planet_two_radius_xy if cfg.planet_two_radius_xy is not None else [0.75, 0.8],
                      euler=cfg.planet_two_euler_angles if cfg.planet_two_euler_angles is not None
                      else [-0.35, 0, 0],
                      period=p1["period"] if time_shift is not None else
                      (cfg.planet_two_period if cfg.planet_two_period is not None else 5.0))  # rsb.py:276-281
            if time_shift is None:
                raise NotImplementedError("independent planets (planet_two_time_shift=None) are not implemented")
            pos1, quat1, loc1, _ = planet_tables(p1["center"], p1["radius_xy"], p1["euler"], p1["period"], upd,
                                                 [0, 0, -np.pi / 2], 1)   # ctlp.py:727-737
            pos2, quat2, _, _ = planet_tables(p2["center"], p2["radius_xy"], p2["euler"], p2["period"], upd,
                                              [0, 0, 0], 4)                # ctlp.py:752-763
            assert len(pos1) == len(pos2)
            self.planet_pos, self.planet_quat, self.planet_local_xy = [pos1, pos2], [quat1, quat2], loc1
            sc.planet_steps = len(pos1)
            sc.planet_shift = int(len(pos2) * time_shift / p2["period"])  # ctlp.py:4435-4437
            for o, (p, q) in enumerate(((pos1, quat1), (pos2, quat2))):
                sc.planet_pos[o] = p.ctypes.data_as(abi.dp)
                sc.planet_quat[o] = q.ctypes.data_as(abi.dp)
            sc.planet_local_xy = loc1.ctypes.data_as(abi.dp)
            sc.planet_obs_half[0], sc.planet_obs_half[1] = 1.05 * p1["radius_xy"][0], 1.05 * p1["radius_xy"][1]
            sc.obs_planet_size = int(cfg.obs_planet_size_per_planet)
            if sc.obs_planet_size not in (1, 2):
                raise NotImplementedError("obs_planet_size_per_planet must be 1 or 2")
            observed = ["iiwa_link_3", "iiwa_link_4", "iiwa_link_5", "iiwa_link_6", "iiwa_link_7"] + \
                ([target_link] if cfg.ball_machine_mode else [])          # robot_scene_base.py:288-292
            for body in (iss, ast):
                obstacles.append(dict(kind=abi.SM_OBST_PLANET, body=body, observed=observed))
        if cfg.use_moving_objects:
            if cfg.planet_mode:
                raise NotImplementedError("planets and balls in one scene are not implemented")
            if cfg.moving_object_sphere_center is None:
                raise NotImplementedError("balls released from a plane (moving_object_area_*) are not implemented; "
                                          "set moving_object_sphere_center as in README.md:81")
            radius = 0.038                                                 # robot_scene_base.py:599-601
            ball = _Body(assets, "obstacle_basketball_red", scale=radius / 0.12)  # ctlp.py:4043-4044
            observed = ["iiwa_link_2", "iiwa_link_3", "iiwa_link_4", "iiwa_link_5", "iiwa_link_6", "iiwa_link_7"] + \
                ([target_link] if cfg.ball_machine_mode else [])          # robot_scene_base.py:226-230
            obstacles.append(dict(kind=abi.SM_OBST_BALL, body=ball, observed=observed))
            self._fill_ball(sc, cfg, radius)
        assert len(obstacles) <= abi.SM_MAX_OBSTACLES
        # robot shapes that may touch a moving obstacle in the simulation client: every link
        contact_names = [n for n in robot.link_names if n in link_shapes]
        mov_contact = [sid for n in contact_names for sid in link_shapes[n]]
        mov_reward = []
        sc.n_obstacles = len(obstacles)
        for o, ob in enumerate(obstacles):
            body = ob["body"]
            if ob["kind"] == abi.SM_OBST_HUMAN:
                self._add_human(sc, cfg, body, add_shape, shapes, robot, link_disc, mov_contact,
                                table_sid if cfg.obstacle_scene == 5 else None, assets)
                mov_reward = [sid for n in ob["observed"] for sid in link_shapes[n]]
                continue
            sids = [add_shape(v, 100 + o, -1) for v in body.parts]
            sc.obst_kind[o] = ob["kind"]
            sc.obst_shape_off[o], sc.obst_shape_cnt[o] = sids[0], len(sids)
            allv = np.concatenate(body.parts)
            centre = 0.5 * (allv.min(0) + allv.max(0))
            sc.obst_center[o][:] = list(centre)
            sc.obst_radius[o] = float(np.linalg.norm(allv - centre, axis=1).max()) + URDF_MARGIN
            disc_o = self._angular_motion_disc(body.parts, body.inertial_xyz[0])
            for slot, sid in enumerate(mov_contact):
                name = robot.link_names[shapes[sid]["link"]]
                sc.contact_thresh[o][slot] = CONTACT_BREAKING_FACTOR * min(disc_o, link_disc[name])
            if not mov_reward:
                mov_reward = [sid for n in ob["observed"] for sid in link_shapes[n]]
        compute_moving = cfg.collision_avoidance_moving_obstacles_max_reward != 0 or \
            cfg.terminate_on_collision_with_moving_obstacle   # rewards.py:133-134
        if not compute_moving:
            mov_reward = []

        # ---------------- write shapes / pairs
        assert len(shapes) <= abi.SM_MAX_SHAPES, len(shapes)
        assert len(static_pairs) <= abi.SM_MAX_PAIRS and len(self_pairs) <= abi.SM_MAX_PAIRS
        assert len(mov_reward) <= abi.SM_MAX_MOV_ROBOT and len(mov_contact) <= abi.SM_MAX_MOV_ROBOT
        self.verts = np.ascontiguousarray(np.concatenate(verts))
        self.shapes = shapes
        sc.n_shapes, sc.n_verts = len(shapes), len(self.verts)
        sc.verts = self.verts.ctypes.data_as(abi.dp)
        for i, sh in enumerate(shapes):
            s = sc.shapes[i]
            s.frame, s.vert_off, s.vert_cnt, s.link, s.margin = sh["frame"], sh["off"], sh["cnt"], sh["link"], \
                sh["margin"]
            s.center[:] = list(sh["center"])
            s.radius = sh["radius"]
        sc.n_static_pairs = len(static_pairs)
        for i, (a, b) in enumerate(static_pairs):
            sc.static_pairs[i][0], sc.static_pairs[i][1] = a, b
        sc.n_self_pairs = len(self_pairs)
        for i, (a, b) in enumerate(self_pairs):
            sc.self_pairs[i][0], sc.self_pairs[i][1] = a, b
        sc.n_mov_reward = len(mov_reward)
        for i, sid in enumerate(mov_reward):
            sc.mov_reward[i] = sid
        sc.n_mov_contact = len(mov_contact)
        for i, sid in enumerate(mov_contact):
            sc.mov_contact[i] = sid
        self.static_pairs, self.self_pairs, self.mov_reward, self.mov_contact = static_pairs, self_pairs, \
            mov_reward, mov_contact

        # ---------------- distances and reward (rewards.py:81-93, ctlp.py:354-366)
        rmax = -1.0
        if cfg.collision_avoidance_self_collision_max_reward != 0:
            rmax = cfg.collision_avoidance_self_collision_max_reward_distance
        if cfg.collision_avoidance_static_obstacles_max_reward != 0 and \
                cfg.collision_avoidance_static_obstacles_max_reward_distance > rmax:
            rmax = cfg.collision_avoidance_static_obstacles_max_reward_distance
        if rmax == -1.0:
            sc.static_cap = cfg.closest_point_safety_distance + 0.002
        else:
            if rmax <= cfg.closest_point_safety_distance:
                raise ValueError("reward_maximum_relevant_distance {} needs to be greater than "
                                 "closest_point_safety_distance {}".format(rmax, cfg.closest_point_safety_distance))
            sc.static_cap = rmax + 0.002
        sc.moving_query = 0.001 if cfg.collision_avoidance_moving_obstacles_max_reward == 0 else \
            cfg.collision_avoidance_moving_obstacles_max_reward_distance   # rewards.py:142-145
        sc.collision_dist = 0.001
        sc.w_self = cfg.collision_avoidance_self_collision_max_reward
        sc.w_static = cfg.collision_avoidance_static_obstacles_max_reward
        sc.w_moving = cfg.collision_avoidance_moving_obstacles_max_reward if compute_moving else 0.0
        sc.d_self = cfg.collision_avoidance_self_collision_max_reward_distance
        sc.d_static = cfg.collision_avoidance_static_obstacles_max_reward_distance
        sc.d_moving = cfg.collision_avoidance_moving_obstacles_max_reward_distance
        sc.w_low_acc = cfg.collision_avoidance_low_acceleration_max_reward
        sc.thr_low_acc = cfg.collision_avoidance_low_acceleration_threshold
        sc.w_low_vel = cfg.collision_avoidance_low_velocity_max_reward
        sc.thr_low_vel = cfg.collision_avoidance_low_velocity_threshold
        sc.punish_action = int(bool(cfg.punish_action))
        sc.terminate_self = int(bool(cfg.terminate_on_self_collision))
        sc.terminate_static = int(bool(cfg.terminate_on_collision_with_static_obstacle))
        sc.terminate_moving = int(bool(cfg.terminate_on_collision_with_moving_obstacle))
        sc.action_thresh = cfg.action_punishment_min_threshold
        sc.action_max_punishment = cfg.action_max_punishment
        sc.termination_bonus = cfg.collision_avoidance_episode_termination_bonus
        sc.early_termination_punishment = cfg.collision_avoidance_episode_early_termination_punishment
        sc.episode_steps = int(round(cfg.trajectory_duration / ts))   # trajectory_manager.py:87, :187-192
        sc.reward_scale = ts / 0.1 if cfg.normalize_reward_to_frequency else 1.0   # rewards.py:172-176
        obs_size = 3 * nj                                            # observations.py:54-110
        if cfg.use_moving_objects:
            obs_size += 6
        if cfg.planet_mode:
            obs_size += sc.obs_planet_size  # planet two is phase-coupled (observations.py:94-98)
        if human is not None:
            obs_size += 3 * sc.human.n_joints   # Human.kinematic_observation (observations.py:100-110, :294-307)
        # ---------------- target points of the reaching task (observations.py:56-77; ctlp.py:184-235)
        sc.use_target_points = int(bool(cfg.use_target_points))
        self.obs_target_size = 0
        if cfg.use_target_points:
            sc.obs_add_tp_pos = int(bool(cfg.obs_add_target_point_pos))
            sc.obs_add_tp_rel = int(bool(cfg.obs_add_target_point_relative_pos))
            self.obs_target_size = 3 * sc.obs_add_tp_pos + 3 * sc.obs_add_tp_rel
            obs_size += self.obs_target_size
            sc.tp_normalize = int(bool(cfg.normalize_reward_to_initial_target_point_distance))
            sc.tp_radius = float(cfg.target_point_radius)
            sc.tp_bonus = float(cfg.target_point_reached_reward_bonus)
            sc.tp_reward_factor = float(cfg.target_point_reward_factor)
            tbox = {0: [[-0.6, 0.6], [-0.8, 0.8], [0.1, 1]], 1: [[-0.6, 0.6], [-0.3, 0.3], [0.1, 1]],
                    2: [[-0.4, 0.4], [-0.4, 0.4], [0.1, 1]]}.get(cfg.target_point_cartesian_range_scene)
            if tbox is None or cfg.target_point_relative_pos_scene != 0:
                raise NotImplementedError("target_point_cartesian_range_scene / target_point_relative_pos_scene")
            rel = [[-1.6, -2, -1.5], [1.6, 2, 1.5]]                         # ctlp.py:184
            for i in range(3):
                sc.tp_box_min[i], sc.tp_box_max[i] = tbox[i][0], tbox[i][1]
                sc.tp_rel_min[i], sc.tp_rel_max[i] = rel[0][i], rel[1][i]
            sc.tp_min_static = cfg.closest_point_safety_distance + 0.09     # ctlp.py:1661-1662
            sc.tp_min_self = cfg.closest_point_safety_distance              # ctlp.py:2220
        sc.start_at_rest = int(not cfg.collision_avoidance_mode)            # ctlp.py:1461-1500
        sc.obs_size = obs_size
        self.obs_size = obs_size

        # ---------------- start-state sampling (ctlp.py:171-183, :1461-1656)
        box = {0: [[-0.6, 0.6], [-0.8, 0.8], [0.1, 1]], 1: [[-1.1, 1.1], [-1.1, 1.1], [-0.1, 1.5]],
               2: [[-0.1, 0.6], [-0.8, 0.8], [-0.1, 1.2]]}.get(cfg.starting_point_cartesian_range_scene,
                                                               [[-0.6, 0.6], [-0.8, 0.8], [0.1, 1]])
        for i in range(3):
            sc.start_box_min[i], sc.start_box_max[i] = box[i][0], box[i][1]
        offset = cfg.target_link_offset
        if offset is None:  # robot_scene_base.py:196-208
            offset = ([0, 0, 0] if not cfg.use_target_points else [0, 0, 0.10]) if cfg.ball_machine_mode \
                else [0, 0, 0.126]
        sc.target_offset[:] = list(np.asarray(offset, dtype=np.float64))
        tl = robot.link_index(target_link)
        assert link_frame[tl] == nj, "the target link is expected to hang off the last joint"
        sc.target_R[:] = list(link_X[tl][0].reshape(-1))
        sc.target_t[:] = list(link_X[tl][1])
        sc.kinematic_sampling_probability = cfg.collision_avoidance_kinematic_state_sampling_probability \
            if cfg.collision_avoidance_kinematic_state_sampling_mode else 0.0
        sc.stay_in_state_probability = cfg.collision_avoidance_stay_in_state_probability
        sc.min_start_distance = 0.001 if cfg.collision_avoidance_mode else cfg.closest_point_safety_distance + 0.09
        sc.min_start_self = 0.001 if cfg.collision_avoidance_mode else cfg.closest_point_safety_distance + 0.04
        sc.ball_target_min_static = cfg.closest_point_safety_distance + 0.09   # ctlp.py:1725-1726
        sc.ball_target_min_self = cfg.closest_point_safety_distance            # ctlp.py:2357-2358
        sc.has_table = int(cfg.obstacle_scene != 0)
        sc.plane_z = -0.94   # robot_scene_base.py:172

    # ------------------------------------------------------------------------------------------------------
    def _add_human(self, sc, cfg, body, add_shape, shapes, robot, link_disc, mov_contact, table_sid, assets):
        """The human obstacle: kinematic tree, limits, shapes, contact thresholds, the pair list of the nested env's
        braking-trajectory check, target points, sampling constants (ctlp.py:4647-4959, human_network/params.json)."""
        h = sc.human
        he = HUMAN_ENV
        joints, link_frame, link_X = merge_chain(body)
        assert len(joints) == abi.SM_HUMAN_JOINTS
        h.enabled, h.n_joints = 1, len(joints)
        self.human_joints, self.human_link_frame, self.human_link_X, self.human_body = joints, link_frame, link_X, body
        base_r = rpy_to_matrix(HUMAN_BASE_EULER)
        h.base_R[:] = list(base_r.reshape(-1))
        h.base_t[:] = list(np.asarray(HUMAN_BASE_POSITION, dtype=np.float64))
        ts = float(cfg.trajectory_time_step)
        if abs(ts - he["trajectory_time_step"]) > 1e-12:
            raise NotImplementedError("the human's nested env runs at trajectory_time_step 0.1 (human_network/params.json)")
        for j, jt in enumerate(joints):
            h.joint_parent[j] = jt["parent_frame"]
            h.joint_R[j][:] = list(jt["R"].reshape(-1))
            h.joint_t[j][:] = list(jt["t"])
            h.joint_axis[j][:] = list(jt["axis"])
            h.pos_lo[j] = jt["limit"][0] * he["pos_limit_factor"]      # no safety buffer (robot_scene_base.py:351-352)
            h.pos_hi[j] = jt["limit"][1] * he["pos_limit_factor"]
            h.vel_max[j] = jt["limit"][3] * he["vel_limit_factor"]
            acc = MAX_ACCELERATION_HUMAN_ARM[j % 4] * he["acc_limit_factor"]
            h.acc_max[j] = acc
            h.jerk_max[j] = min(2 * acc / ts, MAX_JERK_HUMAN_ARM[j % 4]) * he["jerk_limit_factor"]
        self.human_pos_lo = np.array([h.pos_lo[j] for j in range(8)])
        self.human_pos_hi = np.array([h.pos_hi[j] for j in range(8)])
        self.human_vel_max = np.array([h.vel_max[j] for j in range(8)])
        self.human_acc_max = np.array([h.acc_max[j] for j in range(8)])
        # ---- shapes: the arms (obstacle links, ctlp.py:4772-4774) first, then the rest of the body
        arm_links = ["upper_arm_r0", "forearm_r0", "hand_r0", "upper_arm_r1", "forearm_r1", "hand_r1"]
        body_links = ["shoes", "lower_legs", "upper_legs", "body", "head"]
        hshapes = {}   # link name -> shape ids
        link_ids = {}
        h.shape_off = len(shapes)
        for name in arm_links + body_links:
            li = body.link_index(name)
            link_ids[name] = len(link_ids)
            r_x, t_x = link_X[li]
            for pi, v in enumerate(body.parts):
                if int(body.part_link[pi]) != li or body.part_kind[pi] == "point":
                    continue   # the 1e-5 m spheres of the dummy links are left out
                sid = add_shape(v @ r_x.T + t_x, 100 + int(link_frame[li]), -1)
                h.shape_link[sid - h.shape_off] = link_ids[name]
                hshapes.setdefault(name, []).append(sid)
            if name == arm_links[-1]:
                h.n_arm_shapes = len(shapes) - h.shape_off
        h.n_shapes = len(shapes) - h.shape_off
        assert h.n_shapes <= 64 and len(link_ids) <= abi.SM_MAX_HLINKS
        self.human_shapes = hshapes
        sc.obst_kind[0] = abi.SM_OBST_HUMAN
        sc.obst_shape_off[0], sc.obst_shape_cnt[0] = h.shape_off, h.n_arm_shapes
        # bounding sphere of the whole human about its base for the coarse tests: arms stretched out reach ~0.9 m
        # from the shoulders; computed from the link frames at every joint-limit corner would be tighter, a sphere
        # about the trunk centre with the arm length is enough here
        allv = np.concatenate([shapes[s]["verts"] for n in body_links for s in hshapes[n]])
        centre = 0.5 * (allv.min(0) + allv.max(0))
        sc.obst_center[0][:] = list(centre)
        sc.obst_radius[0] = float(np.linalg.norm(allv - centre, axis=1).max()) + 1.0
        # ---- contact thresholds per (human link, robot contact slot)
        for name, lid in link_ids.items():
            li = body.link_index(name)
            disc_h = self._angular_motion_disc([body.parts[pi] for pi in range(len(body.parts))
                                                if int(body.part_link[pi]) == li and body.part_kind[pi] != "point"],
                                               body.inertial_xyz[li])
            for slot, sid in enumerate(mov_contact):
                rname = robot.link_names[shapes[sid]["link"]]
                h.contact_thresh[lid][slot] = CONTACT_BREAKING_FACTOR * min(disc_h, link_disc[rname])
        # ---- braking-trajectory check of the nested env: table x {forearm, hand} (ctlp.py:2503-2518), self-collision
        # link pairs in which forearm or hand takes part (ctlp.py:1409-1437, :3347-3374)
        if table_sid is None:   # the nested env has its own table (obstacle_scene = 1 > 0) even if the main scene has none
            table = _Body(assets, "obstacle_table")
            corners = table.parts[0]
            centre_t = corners.mean(0)
            half = np.abs(corners - centre_t).max(0)
            table_sid = add_shape(centre_t + np.sign(corners - centre_t) * (half - URDF_MARGIN), 0, -1)
        active = {"forearm_r0", "hand_r0", "forearm_r1", "hand_r1"}
        pairs = []
        for name in ("forearm_r0", "hand_r0", "forearm_r1", "hand_r1"):
            pairs += [(a, table_sid) for a in hshapes[name]]
        self_links = {n: [] for n in arm_links}
        for n in arm_links[:3]:
            self_links[n] += arm_links[3:]                      # arm r0 against arm r1
        for n in arm_links:
            self_links[n] += ["body", "head"]
        for a_name in arm_links:
            for b_name in self_links[a_name]:
                if a_name in active or b_name in active:
                    pairs += [(a, b) for a in hshapes[a_name] for b in hshapes[b_name]]
        assert len(pairs) <= abi.SM_MAX_HPAIRS, len(pairs)
        h.n_brake_pairs = len(pairs)
        for i, (a, b) in enumerate(pairs):
            h.brake_pairs[i][0], h.brake_pairs[i][1] = a, b
        self.human_brake_pairs = pairs
        h.check_braking = int(bool(he["check_braking_trajectory_collisions"]))
        h.brake_checks = int(round(max(1, ts / he["collision_check_time"])))   # ctlp.py:99-103
        h.brake_safety = he["closest_point_safety_distance"]
        h.brake_timeout = BRAKING_TIMEOUT
        # ---- target points (target_point_cartesian_range_scene 9, ctlp.py:186-189, :221-223)
        for r, name in enumerate(("hand_r0", "hand_r1")):
            li = body.link_index(name)
            assert link_frame[li] == 4 * (r + 1), "the hand is expected to hang off the arm's last joint"
            r_x, t_x = link_X[li]
            h.tp_local[r][:] = list(r_x @ np.asarray(he["target_link_offset"]) + t_x)
        box = [[0.0, 0.6], [-0.8, 0.8], [0.075, 0.75]]
        rel = [[-1.4, -1.6, -1.5], [1.4, 1.6, 1.5]]
        for i in range(3):
            h.tp_box_min[i], h.tp_box_max[i] = box[i][0], box[i][1]
            h.tp_rel_min[i], h.tp_rel_max[i] = rel[0][i], rel[1][i]
            h.start_box_min[i], h.start_box_max[i] = box[i][0], box[i][1]     # ctlp.py:186-188
        h.tp_radius = he["target_point_radius"]
        h.tp_min_static = he["closest_point_safety_distance"] + 0.09         # ctlp.py:1661-1662
        h.tp_min_self = he["closest_point_safety_distance"]                  # ctlp.py:2220
        h.log_std_lo, h.log_std_hi = he["log_std_range"]
        sampling = bool(cfg.human_network_use_collision_avoidance_starting_point_sampling)
        h.kinematic_sampling_probability = \
            cfg.human_network_collision_avoidance_kinematic_state_sampling_probability if sampling else 0.0
        h.stay_in_state_probability = cfg.human_network_collision_avoidance_stay_in_state_probability if sampling else 1.0
        h.min_start_static = he["closest_point_safety_distance"] + 0.09      # ctlp.py:1475-1477
        h.min_start_self = he["closest_point_safety_distance"] + 0.04
        h.obs_size = 3 * 8 + 2 * 3 + 2 * 3 + 2                               # observations.py:54-77
        h.initial_braking_trajectory = int(sampling)                         # safe_motions_base.py:961-967

    @staticmethod
    def _angular_motion_disc(parts, inertial_xyz):
        """btCollisionShape::getAngularMotionDisc of a link's collision object: the shape is
        expressed in the link's inertial frame; disc = |aabb centre| + half the aabb diagonal (margins included)."""
        allv = np.concatenate(parts) - np.asarray(inertial_xyz)
        lo, hi = allv.min(0) - URDF_MARGIN, allv.max(0) + URDF_MARGIN
        return float(np.linalg.norm(0.5 * (lo + hi)) + 0.5 * np.linalg.norm(hi - lo))

    @staticmethod
    def _fill_ball(sc, cfg, radius):
        centre = np.asarray(cfg.moving_object_sphere_center, dtype=np.float64)
        sph_r = cfg.moving_object_sphere_radius if cfg.moving_object_sphere_radius is not None else 5
        h = cfg.moving_object_sphere_height_min_max
        if h is None:
            h = [-0.1 * sph_r, 0.1 * sph_r]
        ang = cfg.moving_object_sphere_angle_min_max if cfg.moving_object_sphere_angle_min_max is not None \
            else [0, 2 * np.pi]
        speed = float(cfg.moving_object_speed_meter_per_second)
        final = np.array([[-1.0, 1.0], [-1.3, 1.3], [-0.3, 1.5]]).T            # ctlp.py:304-308
        pos_mm = np.array([[-(centre[0] + sph_r), centre[0] + sph_r], [-(centre[1] + sph_r), centre[1] + sph_r],
                           [final[0][2], final[1][2]]]).T                        # ctlp.py:309-317
        max_initial_height = centre[2] + h[1]                                    # ctlp.py:327-328
        t_h = speed / 9.81
        pos_mm[1][2] = max_initial_height + speed * t_h - 0.5 * 9.81 * t_h ** 2  # ctlp.py:330-333
        zero_t = ball_target_height_time(max_initial_height, -speed, radius)     # ctlp.py:335-339
        min_speed = -speed - 9.81 * zero_t
        vel_mm = np.array([[-speed, speed], [-speed, speed], [min_speed, -min_speed]]).T
        for i in range(3):
            sc.ball_obs_pos_min[i], sc.ball_obs_pos_max[i] = pos_mm[0][i], pos_mm[1][i]
            sc.ball_obs_vel_min[i], sc.ball_obs_vel_max[i] = vel_mm[0][i], vel_mm[1][i]
            sc.ball_sphere_center[i] = centre[i]
            sc.ball_final_min[i], sc.ball_final_max[i] = final[0][i], final[1][i]
        sc.ball_active_xy = 1.25                                                 # ctlp.py:2830
        sc.ball_sphere_radius = sph_r
        sc.ball_height_min, sc.ball_height_max = h[0], h[1]
        sc.ball_angle_min, sc.ball_angle_max = ang[0], ang[1]
        sc.ball_speed, sc.ball_radius = speed, radius
        sc.ball_high_angle_probability = cfg.moving_object_high_launch_angle_probability
        tbox = {0: [[-0.6, 0.6], [-0.8, 0.8], [0.1, 1]], 1: [[-0.6, 0.6], [-0.3, 0.3], [0.1, 1]],
                2: [[-0.4, 0.4], [-0.4, 0.4], [0.1, 1]]}.get(cfg.target_point_cartesian_range_scene)
        if tbox is None:
            raise NotImplementedError("target_point_cartesian_range_scene {}".format(
                cfg.target_point_cartesian_range_scene))
        inv = [[-0.4, 0.0], [-0.2, 0.2], [0.0, 0.5]]                             # ctlp.py:706-710
        for i in range(3):
            sc.ball_target_box_min[i], sc.ball_target_box_max[i] = tbox[i][0], tbox[i][1]
            sc.ball_invalid_min[i], sc.ball_invalid_max[i] = inv[i][0], inv[i][1]
        sc.ball_check_invalid = int(bool(cfg.moving_object_check_invalid_target_link_point_positions))
        sc.ball_random_initial = int(bool(cfg.moving_object_random_initial_position))

    # ------------------------------------------------------------------------------------------------------
    def pointer(self):
        return C.byref(self.struct)

    def describe(self):
        sc = self.struct
        return dict(n_joints=sc.n_joints, n_shapes=sc.n_shapes, n_verts=sc.n_verts,
                    n_static_pairs=sc.n_static_pairs, n_self_pairs=sc.n_self_pairs, n_mov_reward=sc.n_mov_reward,
                    n_mov_contact=sc.n_mov_contact, n_obstacles=sc.n_obstacles, obs_size=sc.obs_size,
                    episode_steps=sc.episode_steps, substeps=sc.substeps)
