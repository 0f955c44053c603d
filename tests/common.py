"""Shared helpers for the test-suite."""
import numpy as np

from staticfusion_b200 import synth

_cache = {}


def frames(scene, n, rows, cols, start=10):
    key = (scene, n, rows, cols, start)
    if key not in _cache:
        _cache[key] = synth.render_sequence(scene, n, rows, cols, start=start)
    return _cache[key]


def so3_log_angle(R):
    """Rotation angle of a (near-)rotation matrix, robust for tiny angles."""
    R = np.asarray(R, np.float64)
    s = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = 0.5 * (np.trace(R) - 1.0)
    return float(np.arctan2(np.linalg.norm(s), c))


def pose_error(Ta, Tb):
    """(translation error [m], rotation error [rad]) between two 4x4 increments."""
    Ta, Tb = np.asarray(Ta, np.float64), np.asarray(Tb, np.float64)
    dt = float(np.abs(Ta[:3, 3] - Tb[:3, 3]).max())
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    return dt, so3_log_angle(dR)


def oracle_params_from(O, p):
    """oracle Params with the same field values as a product SfParams."""
    return O.Params(**{name: getattr(p, name) for name, _ in O.Params._fields_})
