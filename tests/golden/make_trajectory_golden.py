"""Generate tests/golden/reference_trajectory.npz with THE REFERENCE'S OWN trajectory writers: Datasets::writeTrajectoryFile
(Utils/Datasets.cpp:252-266) and the pose-graph block of Reconstruction::savePly (Reconstruction.cpp:460-485), compiled from
where they lie against oracle/ref_shim (oracle/Makefile target ref_traj; the number formatting is the real C++ library).

    python tests/golden/make_trajectory_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def trajectory_inputs(n=48, seed=9):
    """Composed camera poses (float32, as Reconstruction.cpp:256,265 forms them), timestamps and the ddt.sumAll() values."""
    from scipy.spatial.transform import Rotation

    from staticfusion_b200 import tum_io
    rng = np.random.default_rng(seed)
    P = np.eye(4, dtype=np.float32)
    poses, ts_obs, t_us, ddt = [], [], [], []
    for k in range(n):
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = Rotation.from_rotvec(rng.normal(size=3) * (0.02 if k % 7 else 2.5)).as_matrix().astype(np.float32)
        T[:3, 3] = rng.normal(size=3).astype(np.float32) * 0.01
        P = tum_io.pose_compose(P, T)
        poses.append(P.copy())
        ts_obs.append(1311868164.3631 + 0.0333 * k)
        t_us.append(1311868164363181 + 33333 * k)
        ddt.append(0.0 if k == 3 else 1.25)  # frame 3 repeats the depth image: no line (Datasets.cpp:255)
    return np.stack(poses), np.array(ts_obs, np.float64), np.array(t_us, np.uint64), np.array(ddt, np.float32)


def run_reference(poses, ts_obs, t_us, ddt):
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_traj")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref_traj"])
    lines = [str(len(poses))]
    for P, a, b, c in zip(poses, ts_obs, t_us, ddt):
        lines.append("%.17g %d %.9g " % (a, int(b), c) + " ".join("%.9g" % x for x in P.reshape(16)))
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "t")
        subprocess.run([tool, prefix], input="\n".join(lines) + "\n", text=True, check=True)
        return open(prefix + ".txt").read(), open(prefix + ".freiburg").read()


def main():
    poses, ts_obs, t_us, ddt = trajectory_inputs()
    dataset_txt, freiburg_txt = run_reference(poses, ts_obs, t_us, ddt)
    np.savez_compressed(os.path.join(HERE, "reference_trajectory.npz"), poses=poses, ts_obs=ts_obs, t_us=t_us, ddt=ddt,
                        dataset_txt=np.array(dataset_txt), freiburg_txt=np.array(freiburg_txt))
    print(dataset_txt.splitlines()[0]); print(freiburg_txt.splitlines()[0]); print(len(dataset_txt.splitlines()), len(freiburg_txt.splitlines()))


if __name__ == "__main__":
    main()
