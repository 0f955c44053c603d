"""Seeded synthetic RGB-D sequences shaped like the inputs StaticFusion consumes.

The reference ships no data and the TUM sequences are unreachable offline, so every
test / bench input is rendered analytically here:

* pinhole camera with the intrinsics the solver itself assumes
  (``f = cols / (2 tan(fovh/2))`` on BOTH axes, principal point ``((cols-1)/2, (rows-1)/2)``,
  reference ``FrontEnd.cpp:378-380``);
* a 6 x 3 x 5 m textured box room, a few static boxes, optionally two person-sized
  cuboids moving laterally (the "dynamic" scenes);
* depth = z-depth + Kinect-like noise, quantised to 1 mm, > 4.5 m dropped, random holes
  and holes along depth discontinuities (inputs arrive as u16 millimetres,
  ``FrontEnd.cpp:243``, ``Utils/Datasets.cpp:178-179``);
* intensity = 8-bit RGB texture mapped to ``0.299 r + 0.587 g + 0.114 b`` in float32
  (``FrontEnd.cpp:232-236``).

Everything is a pure function of (scene name, frame index, resolution) through
``numpy.random.default_rng(7919 * config_id + frame_idx)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

FOVH = np.float32(math.pi * 62.5 / 180.0)  # FrontEnd.cpp:57


@dataclass(frozen=True)
class Scene:
    name: str
    config_id: int
    dynamic: bool
    # per-frame camera motion model
    trans_amp: tuple  # metres, sinusoid amplitude per axis
    trans_per_frame: tuple  # metres/frame constant drift per axis
    yaw_per_frame: float  # rad/frame about camera y (vertical) axis
    rot_amp: tuple  # rad, sinusoid amplitude (pitch, yaw, roll)


SCENES = {
    # config 1: static box room, 7 mm + 0.0242 rad yaw per frame
    "static_yaw": Scene("static_yaw", 1, False, (0, 0, 0), (0.007, 0.0, 0.0), 0.0242, (0, 0, 0)),
    # "fr1/360-shaped": static, rotation dominant
    "fr1_360": Scene("fr1_360", 6, False, (0.02, 0.01, 0.02), (0.004, 0.0, 0.004), 0.0242, (0.004, 0, 0.003)),
    # configs 2, 3, 5: dynamic scene, moderate motion
    "dynamic": Scene("dynamic", 2, True, (0.05, 0.03, 0.05), (0.003, 0.0, 0.002), 0.0032, (0.004, 0.006, 0.003)),
    # config 4: fr3/walking_xyz-shaped: translation dominant sinusoids, 7 mm/frame, 0.0032 rad/frame
    "walking_xyz": Scene("walking_xyz", 4, True, (0.25, 0.12, 0.18), (0, 0, 0), 0.0, (0.01, 0.012, 0.006)),
    # small static motion for unit tests
    "static_small": Scene("static_small", 7, False, (0, 0, 0), (0.004, -0.002, 0.003), 0.004, (0, 0, 0)),
}


def focal(cols: int) -> float:
    return float(cols) / (2.0 * math.tan(0.5 * float(FOVH)))


def _rot(rx: float, ry: float, rz: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Ry @ Rx @ Rz


def camera_pose(scene: Scene, t: int) -> np.ndarray:
    """World-from-camera 4x4 (float64) at frame t.  Camera axes: x right, y down, z forward."""
    if isinstance(scene, str):
        scene = SCENES[scene]
    w = 2.0 * math.pi / 90.0  # 3 s period at 30 Hz
    ax, ay, az = scene.trans_amp
    dx, dy, dz = scene.trans_per_frame
    pos = np.array([
        ax * math.sin(w * t) + dx * t,
        ay * math.sin(1.3 * w * t + 0.7) + dy * t,
        az * math.sin(0.8 * w * t + 1.9) + dz * t,
    ])
    rp, ryw, rr = scene.rot_amp
    R = _rot(rp * math.sin(0.9 * w * t + 0.3), scene.yaw_per_frame * t + ryw * math.sin(1.1 * w * t), rr * math.sin(0.7 * w * t + 1.1))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = pos
    return T


def relative_pose(scene: Scene, t_prev: int, t_cur: int) -> np.ndarray:
    """Ground-truth T_odometry for the pair: pose of the current camera in the previous camera's frame."""
    return np.linalg.inv(camera_pose(scene, t_prev)) @ camera_pose(scene, t_cur)


# ---- scene geometry: axis-aligned boxes in world coordinates (x right, y down, z forward) ----
_ROOM = (np.array([-3.0, -1.6, -2.0]), np.array([3.0, 1.4, 3.0]))  # 6 x 3 x 5 m, camera near the back
_STATIC_BOXES = [
    (np.array([-1.6, 0.5, 1.4]), np.array([-0.8, 1.4, 2.0])),
    (np.array([0.9, 0.2, 1.8]), np.array([1.7, 1.4, 2.6])),
    (np.array([-0.4, 0.9, 2.2]), np.array([0.5, 1.4, 2.9])),
    (np.array([-2.9, -0.2, 0.6]), np.array([-2.3, 1.4, 1.5])),
    (np.array([2.1, -0.6, 0.9]), np.array([2.9, 1.4, 1.6])),
]


def _moving_boxes(t: int):
    # two person-sized cuboids (0.5 x 1.7 x 0.35 m) moving laterally 0.5-1 m/s (17-33 mm/frame)
    x1 = -1.2 + 1.5 * abs(((t * 0.022) % 2.0) - 1.0) * 1.6 - 0.2
    x2 = 1.4 - 1.5 * abs(((t * 0.017 + 0.6) % 2.0) - 1.0) * 1.5
    return [
        (np.array([x1 - 0.25, -0.3, 1.55]), np.array([x1 + 0.25, 1.4, 1.9])),
        (np.array([x2 - 0.25, -0.3, 2.25]), np.array([x2 + 0.25, 1.4, 2.6])),
    ]


def _hash01(ix, iy, seed):
    h = (ix.astype(np.int64) * 374761393 + iy.astype(np.int64) * 668265263 + seed * 362437) & 0x7FFFFFFF
    h = (h ^ (h >> 13)) * 1274126177 & 0x7FFFFFFF
    return ((h ^ (h >> 16)) & 0xFFFF).astype(np.float64) / 65535.0


def _value_noise(a, b, seed):
    ia, ib = np.floor(a), np.floor(b)
    fa, fb = a - ia, b - ib
    fa = fa * fa * (3 - 2 * fa)
    fb = fb * fb * (3 - 2 * fb)
    ia = ia.astype(np.int64)
    ib = ib.astype(np.int64)
    v00 = _hash01(ia, ib, seed)
    v10 = _hash01(ia + 1, ib, seed)
    v01 = _hash01(ia, ib + 1, seed)
    v11 = _hash01(ia + 1, ib + 1, seed)
    return (v00 * (1 - fa) + v10 * fa) * (1 - fb) + (v01 * (1 - fa) + v11 * fa) * fb


def _texture(p, sid, chan):
    """Procedural texture at world points p (N,3) on surface ids sid: sinusoids + value noise, in [0,1]."""
    a = p[:, 0] + 0.37 * p[:, 1] + 0.11 * sid
    b = p[:, 2] + 0.61 * p[:, 1] - 0.23 * sid
    ph = 0.9 * chan
    s = (0.5 + 0.16 * np.sin(5.1 * a + ph + 0.3 * sid) + 0.13 * np.sin(7.3 * b - ph) + 0.09 * np.sin(13.7 * (a + b) + 2 * ph)
         + 0.07 * np.sin(23.0 * a - 17.0 * b + sid))
    s = s + 0.22 * (_value_noise(9.0 * a, 9.0 * b, 17 + chan) - 0.5) + 0.1 * (_value_noise(31.0 * a, 31.0 * b, 91 + chan) - 0.5)
    return np.clip(s, 0.02, 0.98)


def render_frame_raw(scene: Scene | str, t: int, rows: int, cols: int):
    """Frame t as the reference's loader would read it from disk (FrontEnd.cpp:216-254): (colour uint8 (rows, cols, 3) with
    channel 0 the one the loader calls r, depth uint16 millimetres (rows, cols)), both in FILE orientation, i.e. vertically
    flipped with respect to render_frame (the loader reads row H - v - 1).  Converting them with res_factor = 1 gives
    render_frame's intensity bit for bit and its depth to the last bit of (float)mm * 0.001f against (float)(mm * 0.001)."""
    rgb, mm = _render_core(scene, t, rows, cols, 0.0)
    col = np.stack([c.reshape(rows, cols) for c in rgb], axis=-1)
    return np.ascontiguousarray(col[::-1]), np.ascontiguousarray(mm[::-1])


def render_frame(scene: Scene | str, t: int, rows: int, cols: int, moving_offset_frames: float = 0.0):
    """Render frame t -> (depth float32 [m], intensity float32 [0,1]), both (rows, cols) row-major."""
    rgb, mm = _render_core(scene, t, rows, cols, moving_offset_frames)
    nf = np.float32(1.0 / 255.0)
    r, g, b = (nf * c.astype(np.float32) for c in rgb)
    inten = (np.float32(0.299) * r + np.float32(0.587) * g) + np.float32(0.114) * b
    depth = (mm.astype(np.float64) * (1.0 / 1000.0)).astype(np.float32)
    return np.ascontiguousarray(depth), np.ascontiguousarray(inten.reshape(rows, cols).astype(np.float32))


def _render_core(scene: Scene | str, t: int, rows: int, cols: int, moving_offset_frames: float = 0.0):
    """(three uint8 colour channels of rows*cols values each, uint16 depth in millimetres (rows, cols))."""
    if isinstance(scene, str):
        scene = SCENES[scene]
    f = focal(cols)
    cu, cv = 0.5 * (cols - 1), 0.5 * (rows - 1)
    uu, vv = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    dirs_c = np.stack([(uu - cu) / f, (vv - cv) / f, np.ones_like(uu)], axis=-1).reshape(-1, 3)
    Twc = camera_pose(scene, t)
    R, o = Twc[:3, :3], Twc[:3, 3]
    d = dirs_c @ R.T  # world ray directions, z_cam = ray parameter because dirs_c[:,2] == 1
    n = d.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
    # room: exit distance
    lo, hi = _ROOM
    t1, t2 = (lo - o) * inv, (hi - o) * inv
    tfar_axis = np.maximum(t1, t2)
    ax_room = np.argmin(tfar_axis, axis=1)
    best_t = tfar_axis[np.arange(n), ax_room]
    best_sid = ax_room * 2 + (d[np.arange(n), ax_room] > 0)
    boxes = list(_STATIC_BOXES)
    if scene.dynamic:
        boxes += _moving_boxes(t + moving_offset_frames)
    for bi, (blo, bhi) in enumerate(boxes):
        t1, t2 = (blo - o) * inv, (bhi - o) * inv
        tn_axis = np.minimum(t1, t2)
        tnear = tn_axis.max(axis=1)
        tfar = np.maximum(t1, t2).min(axis=1)
        hit = (tnear < tfar) & (tnear > 0.05) & (tnear < best_t)
        axn = np.argmax(tn_axis, axis=1)
        best_t = np.where(hit, tnear, best_t)
        best_sid = np.where(hit, 10 + bi * 6 + axn * 2 + (d[np.arange(n), axn] > 0), best_sid)
    z = best_t  # z-depth in the camera frame (ray parameter)
    pw = o + d * best_t[:, None]
    if scene.dynamic:  # texture moves with the moving cuboids
        nb = len(_STATIC_BOXES)
        for bi, (blo, _) in enumerate(_moving_boxes(t + moving_offset_frames)):
            m = (best_sid >= 10 + (nb + bi) * 6) & (best_sid < 10 + (nb + bi + 1) * 6)
            pw = np.where(m[:, None], pw - np.array([blo[0], 0, 0]), pw)
    rgb = [np.round(_texture(pw, best_sid.astype(np.float64), c) * 255.0).astype(np.uint8) for c in range(3)]

    rng = np.random.default_rng(7919 * scene.config_id + t)
    sigma = 0.0012 + 0.0019 * (z - 0.4) ** 2
    zn = z + rng.standard_normal(n) * sigma
    mm = np.round(zn * 1000.0)
    mm[(zn > 4.5) | (zn < 0.3)] = 0
    mm[rng.random(n) < 0.05] = 0
    mm = mm.reshape(rows, cols)
    zc = z.reshape(rows, cols)
    disc = np.zeros((rows, cols), dtype=bool)
    jump_u = np.abs(np.diff(zc, axis=1)) > 0.1
    jump_v = np.abs(np.diff(zc, axis=0)) > 0.1
    disc[:, :-1] |= jump_u
    disc[:, 1:] |= jump_u
    disc[:-1, :] |= jump_v
    disc[1:, :] |= jump_v
    mm[disc] = 0
    return rgb, mm.astype(np.uint16)


def render_sequence(scene: Scene | str, n_frames: int, rows: int, cols: int, start: int = 0):
    """(n_frames, rows, cols) depth and intensity stacks."""
    ds = np.empty((n_frames, rows, cols), np.float32)
    cs = np.empty((n_frames, rows, cols), np.float32)
    for i in range(n_frames):
        ds[i], cs[i] = render_frame(scene, start + i, rows, cols)
    return ds, cs
