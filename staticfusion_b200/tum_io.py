"""Wire formats either side of the solver: the recorded-sequence inputs and the trajectory outputs.

Host-side mirror of the reference's file handling for the hot path's callers:

* ``load_assoc``            — StaticFusion::loadAssoc (FrontEnd.cpp:183-214): ``rgbd_assoc.txt`` lines
  ``ts_rgb rgb_path ts_depth depth_path``; comments (``#``) and empty lines skipped, parsing stops at the first
  malformed line, paths are prefixed with ``dir`` verbatim, the DEPTH timestamp is kept.
* ``read_images``           — the ``cv::imread`` calls of loadImageFromSequenceAssoc (FrontEnd.cpp:220,240): 8-bit BGR
  colour and the 16-bit depth PNG as stored (millimetres).  Decoding stays on the host; the flip / decimation /
  intensity conversion runs on the device (``StaticFusionSolver.convertFrames`` / ``upload_sequence_raw``).
* ``Trajectory``            — the pose bookkeeping of the callers: ``currPose = currPose * T_odometry`` in float
  (Reconstruction.cpp:256,265), ``Datasets::writeTrajectoryFile`` (Datasets.cpp:252-266) and the ``.freiburg`` pose
  graph of ``Reconstruction::savePly`` (Reconstruction.cpp:460-484); both are TUM-format text
  ``timestamp tx ty tz qx qy qz qw``.

The C++ twin is ``staticfusion_b200/host/TumIO.hpp``; tests check both against the oracle restatement and, for the
loader, against the reference's own functions compiled from /root/reference (tests/test_tum_io.py).
"""
from __future__ import annotations

import os

import numpy as np


def load_assoc(directory: str, assoc_file: str = "/rgbd_assoc.txt"):
    """StaticFusion::loadAssoc: returns (timestamps, filesDepth, filesColor) or None when the file cannot be opened."""
    path = directory + assoc_file
    if not path or not os.path.isfile(path):
        return None
    ts, fd, fc = [], [], []
    with open(path, "r") as f:
        for line in f.read().split("\n"):
            if line == "" or line.startswith("#"):
                continue
            tok = line.split()
            try:  # iss >> timestampColor >> fileColor >> timestampDepth >> fileDepth, break on failure
                float(tok[0]); file_color = tok[1]; t_depth = float(tok[2]); file_depth = tok[3]
            except (IndexError, ValueError):
                break
            ts.append(t_depth)
            fd.append(directory + file_depth)
            fc.append(directory + file_color)
    return ts, fd, fc


def read_images(depth_file: str, rgb_file: str):
    """cv::imread(rgb, CV_LOAD_IMAGE_COLOR) and cv::imread(depth, -1): (bgr uint8 (H, W, 3), depth uint16 (H, W)) or None
    when the colour image is missing ("End of sequence", FrontEnd.cpp:222-226)."""
    import cv2

    color = cv2.imread(rgb_file, cv2.IMREAD_COLOR)
    if color is None:
        return None
    depth = cv2.imread(depth_file, cv2.IMREAD_UNCHANGED)
    if depth is None or depth.dtype != np.uint16 or depth.ndim != 2:
        raise ValueError(f"{depth_file}: expected a 16-bit single-channel depth image")
    return color, depth


def _g(x) -> str:
    """std::ostream << float with the default precision (6 significant digits, %g)."""
    return "%g" % float(np.float32(x))


def quat_from_rotation(T):
    """Eigen::Quaternionf(Matrix3f) (Eigen/src/Geometry/Quaternion.h), float arithmetic; returns (x, y, z, w)."""
    m = np.asarray(T, np.float32)
    f = np.float32
    q = np.zeros(4, np.float32)
    t = f(f(m[0, 0] + m[1, 1]) + m[2, 2])
    if t > f(0):
        t = np.sqrt(f(t + f(1)))
        q[3] = f(0.5) * t
        t = f(0.5) / t
        q[0] = f(m[2, 1] - m[1, 2]) * t
        q[1] = f(m[0, 2] - m[2, 0]) * t
        q[2] = f(m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(f(f(f(m[i, i] - m[j, j]) - m[k, k]) + f(1)))
        q[i] = f(0.5) * t
        t = f(0.5) / t
        q[3] = f(m[k, j] - m[j, k]) * t
        q[j] = f(m[j, i] + m[i, j]) * t
        q[k] = f(m[k, i] + m[i, k]) * t
    return q


def pose_compose(A, B):
    """Eigen Matrix4f product in float with the sums taken in k order (Reconstruction.cpp:256,265)."""
    a, b = np.asarray(A, np.float32), np.asarray(B, np.float32)
    out = np.zeros((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            s = np.float32(a[i, 0] * b[0, j])
            for k in range(1, 4):
                s = np.float32(s + np.float32(a[i, k] * b[k, j]))
            out[i, j] = s
    return out


class Trajectory:
    """currPose / poseGraph / poseLogTimes of the map back-end, for callers that only want the odometry."""

    def __init__(self):
        self.currPose = np.eye(4, dtype=np.float32)  # Reconstruction.cpp:36
        self.poseGraph = []
        self.poseLogTimes = []
        # Datasets.cpp:58-60: AngleAxisf(M_PI, UnitZ).toRotationMatrix() in float: cos(pi) = -1, sin(float(pi)) = -8.742278e-08
        s, c = np.sin(np.float32(np.pi), dtype=np.float32), np.cos(np.float32(np.pi), dtype=np.float32)
        self.rotateByZ = np.eye(4, dtype=np.float32)
        self.rotateByZ[0, 0] = c; self.rotateByZ[0, 1] = -s; self.rotateByZ[1, 0] = s; self.rotateByZ[1, 1] = c

    def fuse(self, T_odometry, timestamp):
        """The pose part of Reconstruction::fuseFrame (Reconstruction.cpp:255-265, 315-321)."""
        self.currPose = pose_compose(self.currPose, T_odometry)
        self.poseGraph.append(self.currPose.copy())
        self.poseLogTimes.append(int(timestamp))
        return self.currPose

    def dataset_line(self, timestamp_obs: float, ddt_sum: float = 1.0):
        """Datasets::writeTrajectoryFile (Datasets.cpp:252-266): None when consecutive depth images were equal."""
        if not abs(ddt_sum) > 0:
            return None
        P = pose_compose(self.currPose, self.rotateByZ)
        q = quat_from_rotation(P)
        return "%.04f %s %s %s %s %s %s %s\n" % (timestamp_obs, _g(P[0, 3]), _g(P[1, 3]), _g(P[2, 3]), _g(q[0]), _g(q[1]), _g(q[2]), _g(q[3]))

    def freiburg_text(self) -> str:
        """The `.freiburg` pose graph of Reconstruction::savePly (Reconstruction.cpp:460-484)."""
        out = []
        for P, t in zip(self.poseGraph, self.poseLogTimes):
            q = quat_from_rotation(P)
            # the stream keeps std::fixed / precision 6 only for the timestamp's own stringstream
            out.append("%.6f %s %s %s %s %s %s %s\n" % (float(t) / 1000000.0, _g(P[0, 3]), _g(P[1, 3]), _g(P[2, 3]), _g(q[0]), _g(q[1]), _g(q[2]), _g(q[3])))
        return "".join(out)

    def save_freiburg(self, save_filename: str):
        with open(save_filename + ".freiburg", "w") as f:
            f.write(self.freiburg_text())
