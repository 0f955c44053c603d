"""Recorded-sequence inputs and trajectory outputs: StaticFusion::loadAssoc and
loadImageFromSequenceAssoc (FrontEnd.cpp:183-254), the pose bookkeeping of Reconstruction::fuseFrame
(Reconstruction.cpp:255-265, 315-321) and the two TUM-format writers (Datasets.cpp:252-266, Reconstruction.cpp:460-484).

Pinning: the loader restatement (oracle.convert_frame) is compared bit for bit with the reference's own function, live
where oracle/_ref exists and through fixtures generated from it (tests/golden/reference_loader_*.npz); loadAssoc likewise.
The writers need Eigen::Quaternionf and the GL back-end's pose graph (not compilable here): their restatement is checked
against scipy and between the C++ and Python twins ("parity unpinned" for the text formatting).  GPU tests: the CUDA
conversion and the raw-upload path are bit-identical to the oracle; the C++ sequence driver reproduces the Python mirror."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "staticfusion_b200", "host")
sys.path.insert(0, GOLDEN)
from make_loader_golden import ASSOC_TEXT, loader_inputs  # noqa: E402


def host_tool(name):
    subprocess.check_call(["make", "-C", HOST, "-s", name])
    return os.path.join(HOST, name)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same(a, b):
    return all(np.array_equal(bits(x), bits(y)) for x, y in zip(a, b))


# ---------------------------------------------------------------------------------------------- loader: oracle pinned
@pytest.mark.parametrize("rf", [4, 8])
def test_oracle_loader_matches_reference_fixture(oracle_mod, rf):
    g = np.load(os.path.join(GOLDEN, f"reference_loader_rf{rf}.npz"))
    bgr, depth = loader_inputs(int(g["seed"]))
    got = oracle_mod.convert_frame(bgr, depth, rf)
    assert same(got, (g["intensity"], g["depth"], g["depth_mm"], g["color_full"]))


def test_oracle_loader_matches_reference_live(oracle_mod):
    from oracle import reference as R
    if not R.available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    rng = np.random.default_rng(11)
    for rf in (1, 2, 4):
        bgr = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        depth = rng.integers(0, 65536, (480, 640), dtype=np.uint16)
        assert same(oracle_mod.convert_frame(bgr, depth, rf), R.Reference(rf).load_image_from_sequence_assoc(bgr, depth, rf))


def test_loader_semantics(oracle_mod):
    """vertical flip + decimation (:231), BGR channel naming (:232-236), colour truncation (:237), depth scale (:243)."""
    bgr = np.zeros((480, 640, 3), np.uint8)
    depth = np.zeros((480, 640), np.uint16)
    bgr[479, 0] = (255, 0, 0)      # source row H*rf - 1 -> output row 0; channel 0 carries the 0.299 weight
    bgr[477, 2] = (0, 0, 200)
    depth[479, 0] = 1234
    depth[1, 638] = 65535          # -> output (239, 319)
    inten, dep, mm, col = oracle_mod.convert_frame(bgr, depth, 2)
    assert inten.shape == (240, 320) and col.shape == (240, 320, 3)
    assert inten[0, 0] == np.float32(0.299) * (np.float32(1 / 255) * np.float32(255))
    assert inten[1, 1] == np.float32(0.114) * (np.float32(1 / 255) * np.float32(200))
    assert dep[0, 0] == np.float32(1234) * np.float32(0.001) and mm[0, 0] == 1234
    assert mm[239, 319] == 65535 and dep[239, 319] == np.float32(65535) * np.float32(0.001)
    # Vec3b(r*255, ...) truncates (c * (1/255)) * 255: in float that product returns to c for every byte value, so
    # color_full is the flipped / decimated file image with its channel order kept
    ramp = np.zeros((480, 640, 3), np.uint8)
    ramp[479, ::2, 0] = np.arange(320) % 256
    c = oracle_mod.convert_frame(ramp, depth, 2)[3][0, :256, 0].astype(int)
    assert np.array_equal(c, np.arange(256))


# ---------------------------------------------------------------------------------------------- loadAssoc
def write_assoc(tmp_path):
    d = str(tmp_path) + "/"
    with open(d + "rgbd_assoc.txt", "w") as f:
        f.write(ASSOC_TEXT)
    return d


def test_load_assoc_python_cpp_and_reference(tmp_path):
    from staticfusion_b200 import tum_io
    d = write_assoc(tmp_path)
    g = np.load(os.path.join(GOLDEN, "reference_load_assoc.npz"))
    want = (list(g["timestamps"]), [d + str(x) for x in g["files_depth"]], [d + str(x) for x in g["files_color"]])
    got = tum_io.load_assoc(d, "rgbd_assoc.txt")
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
    assert len(got[0]) == 4 and got[1][0].endswith("depth/1311868164.338541.png") and got[0][0] == 1311868164.338541
    out = subprocess.run([host_tool("tum_io_tool"), "assoc", d, "rgbd_assoc.txt"], capture_output=True, text=True, check=True).stdout
    rows = [ln.split(" ") for ln in out.strip().split("\n")]
    assert [float(r[0]) for r in rows] == want[0] and [r[1] for r in rows] == want[1] and [r[2] for r in rows] == want[2]
    assert tum_io.load_assoc(d, "missing.txt") is None
    assert subprocess.run([host_tool("tum_io_tool"), "assoc", d, "missing.txt"], capture_output=True, text=True).stdout.strip() == "FAILED"
    from oracle import reference as R
    if R.available() and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_assoc")):
        live = R.Reference.load_assoc(d, "rgbd_assoc.txt")
        assert live[0] == want[0] and live[1] == want[1] and live[2] == want[2]
        assert R.Reference.load_assoc(d, "missing.txt") is None


# ---------------------------------------------------------------------------------------------- PNG decode (C++ host)
def test_cpp_png_reader_matches_opencv(tmp_path):
    cv2 = pytest.importorskip("cv2")
    tool = host_tool("tum_io_tool")
    bgr, depth = loader_inputs(7)
    rng = np.random.default_rng(3)
    cases = {
        "rgb.png": (bgr, "color"), "noise.png": (rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), "color"),
        "grey.png": (bgr[:, :, 0].copy(), "color"), "rgba.png": (np.dstack([bgr[:100, :90], bgr[:100, :90, 0]]), "color"),
        "depth.png": (depth, "depth"), "depth_noise.png": (rng.integers(0, 65536, (41, 29), dtype=np.uint16), "depth"),
    }
    for name, (img, kind) in cases.items():
        p = str(tmp_path / name)
        level = 1 if "noise" in name else 6
        assert cv2.imwrite(p, img, [cv2.IMWRITE_PNG_COMPRESSION, level])
        want = cv2.imread(p, cv2.IMREAD_COLOR if kind == "color" else cv2.IMREAD_UNCHANGED)
        out = str(tmp_path / (name + ".bin"))
        r = subprocess.run([tool, kind, p, out], capture_output=True, text=True, check=True).stdout.split()
        assert (int(r[0]), int(r[1])) == want.shape[:2], name
        got = np.fromfile(out, np.uint8 if kind == "color" else np.uint16).reshape(want.shape)
        assert np.array_equal(got, want), name
    assert subprocess.run([tool, "color", str(tmp_path / "nope.png"), "/dev/null"], capture_output=True, text=True).stdout.strip() == "MISSING"
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    assert subprocess.run([tool, "color", str(bad), "/dev/null"], capture_output=True, text=True).returncode == 1


# ---------------------------------------------------------------------------------------------- trajectory
def random_increments(n, seed=5):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    Ts = []
    for k in range(n):
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = Rotation.from_rotvec(rng.normal(size=3) * (0.02 if k % 7 else 2.5)).as_matrix().astype(np.float32)
        T[:3, 3] = rng.normal(size=3).astype(np.float32) * 0.01
        Ts.append(T)
    return Ts


def test_quaternion_and_pose_product(oracle_mod):
    from scipy.spatial.transform import Rotation
    from staticfusion_b200 import tum_io
    Ts = random_increments(120)
    P = np.eye(4, dtype=np.float32)
    for T in Ts:
        assert np.array_equal(bits(tum_io.pose_compose(P, T)), bits(oracle_mod.pose_compose(P, T)))
        P = tum_io.pose_compose(P, T)
        q, qo = tum_io.quat_from_rotation(P), oracle_mod.quat_from_rotation(P)
        assert np.array_equal(bits(q), bits(qo))
        qs = Rotation.from_matrix(P[:3, :3].astype(np.float64)).as_quat()
        assert min(np.abs(qs - q).max(), np.abs(qs + q).max()) < 2e-5
    # both branches of Eigen's conversion are exercised (trace > 0 and each largest-diagonal case)
    for axis in range(3):
        v = np.zeros(3); v[axis] = np.pi * 0.98
        T = np.eye(4, dtype=np.float32); T[:3, :3] = Rotation.from_rotvec(v).as_matrix().astype(np.float32)
        q = tum_io.quat_from_rotation(T)
        assert np.argmax(np.abs(q[:3])) == axis and np.array_equal(bits(q), bits(oracle_mod.quat_from_rotation(T)))


def test_trajectory_writers_python_equals_cpp(tmp_path):
    from staticfusion_b200 import tum_io
    Ts = random_increments(40, seed=9)
    tr = tum_io.Trajectory()
    lines, inp = [], []
    for k, T in enumerate(Ts):
        t_us = 1311868164363181 + 33333 * k
        ts_obs = 1311868164.3631 + 0.0333 * k
        ddt = 0.0 if k == 3 else 1.25                      # frame 3 repeats the depth image: no line (Datasets.cpp:255)
        tr.fuse(T, t_us)
        ln = tr.dataset_line(ts_obs, ddt)
        if ln:
            lines.append(ln)
        inp.append("%d %.17g %.9g " % (t_us, ts_obs, ddt) + " ".join("%.9g" % x for x in T.T.reshape(16)))  # column-major, like sf_get_outputs
    f = tmp_path / "traj_in.txt"
    f.write_text("\n".join(inp) + "\n")
    out = subprocess.run([host_tool("tum_io_tool"), "traj", str(f)], capture_output=True, text=True, check=True).stdout
    cpp_dataset, cpp_freiburg = out.split("--\n")
    assert cpp_dataset == "".join(lines) and len(lines) == 39
    assert cpp_freiburg == tr.freiburg_text()
    first = tr.freiburg_text().split("\n")[0].split(" ")
    assert first[0] == "1311868164.363181" and len(first) == 8
    # the dataset writer reports poses rotated by pi about z (Datasets.cpp:58-60,257)
    P = tum_io.pose_compose(tr.currPose, tr.rotateByZ)
    assert np.allclose(P[:3, 0], -tr.currPose[:3, 0], atol=1e-6) and np.allclose(P[:3, 3], tr.currPose[:3, 3])
    tr.save_freiburg(str(tmp_path / "run"))
    assert (tmp_path / "run.freiburg").read_text() == tr.freiburg_text()


def _writers_on(poses, ts_obs, t_us, ddt, tmp_path):
    """(Datasets text, .freiburg text) of the Python and of the C++ writer for a list of composed poses."""
    from staticfusion_b200 import tum_io
    tr = tum_io.Trajectory()
    lines, inp = [], []
    prev = np.eye(4, dtype=np.float32)
    for P, a, b, c in zip(poses, ts_obs, t_us, ddt):
        tr.currPose = P.astype(np.float32).copy()  # the writers see the composed pose (Reconstruction.cpp:256,265 forms it)
        tr.poseGraph.append(tr.currPose.copy()); tr.poseLogTimes.append(int(b))
        ln = tr.dataset_line(float(a), float(c))
        if ln:
            lines.append(ln)
    return "".join(lines), tr.freiburg_text()


def test_trajectory_writers_match_the_reference_fixture(tmp_path):
    """Text of Datasets::writeTrajectoryFile (Datasets.cpp:252-266) and of savePly's pose graph (Reconstruction.cpp:460-485)
    as THE REFERENCE'S OWN code writes it (tests/golden/reference_trajectory.npz, made by make_trajectory_golden.py from the
    reference sources compiled against the shim): separators, precision, the skipped line, both quaternion branches."""
    g = np.load(os.path.join(GOLDEN, "reference_trajectory.npz"))
    dataset_txt, freiburg_txt = _writers_on(g["poses"], g["ts_obs"], g["t_us"], g["ddt"], tmp_path)
    assert dataset_txt == str(g["dataset_txt"]) and freiburg_txt == str(g["freiburg_txt"])
    assert len(dataset_txt.splitlines()) == len(g["poses"]) - 1  # one frame repeats its depth image (Datasets.cpp:255)


def test_trajectory_cpp_writer_matches_the_reference_fixture(tmp_path):
    """The C++ host twin (TumIO.hpp, via tum_io_tool) on the increments that compose the fixture's poses."""
    from make_trajectory_golden import trajectory_inputs
    from staticfusion_b200 import tum_io
    g = np.load(os.path.join(GOLDEN, "reference_trajectory.npz"))
    # regenerate the increments (same seed) and check that they compose to the stored poses
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(9)
    inp, P = [], np.eye(4, dtype=np.float32)
    for k in range(len(g["poses"])):
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = Rotation.from_rotvec(rng.normal(size=3) * (0.02 if k % 7 else 2.5)).as_matrix().astype(np.float32)
        T[:3, 3] = rng.normal(size=3).astype(np.float32) * 0.01
        P = tum_io.pose_compose(P, T)
        assert np.array_equal(P, g["poses"][k])
        inp.append("%d %.17g %.9g " % (int(g["t_us"][k]), float(g["ts_obs"][k]), float(g["ddt"][k])) + " ".join("%.9g" % x for x in T.T.reshape(16)))
    f = tmp_path / "traj_in.txt"
    f.write_text("\n".join(inp) + "\n")
    out = subprocess.run([host_tool("tum_io_tool"), "traj", str(f)], capture_output=True, text=True, check=True).stdout
    cpp_dataset, cpp_freiburg = out.split("--\n")
    assert cpp_dataset == str(g["dataset_txt"]) and cpp_freiburg == str(g["freiburg_txt"])
    assert trajectory_inputs is not None


def test_trajectory_writers_match_the_live_reference(tmp_path):
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "ref_traj")
    if not os.path.exists(ref_tool) and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/ref_traj is not present and /root/reference is not mounted")
    from make_trajectory_golden import run_reference, trajectory_inputs
    poses, ts_obs, t_us, ddt = trajectory_inputs(n=30, seed=21)
    ref_dataset, ref_freiburg = run_reference(poses, ts_obs, t_us, ddt) if os.path.isdir("/root/reference") else (None, None)
    if ref_dataset is None:
        pytest.skip("needs the reference tree to rebuild the tool")
    dataset_txt, freiburg_txt = _writers_on(poses, ts_obs, t_us, ddt, tmp_path)
    assert dataset_txt == ref_dataset and freiburg_txt == ref_freiburg


# ---------------------------------------------------------------------------------------------- GPU
def sequence_as_files(depth, inten, rf, seed=0):
    """Full-resolution decoded images whose loader output is (close to) the given float frames: grey BGR, millimetre
    depth, nearest-neighbour upsampled by rf and flipped vertically so that the loader's flip undoes it."""
    n, rows, cols = depth.shape
    g8 = np.clip(np.round(inten * 255.0), 0, 255).astype(np.uint8)
    mm = np.clip(np.round(depth * 1000.0), 0, 65535).astype(np.uint16)
    rng = np.random.default_rng(seed)
    bgr = np.repeat(np.repeat(g8, rf, axis=1), rf, axis=2)[:, ::-1]
    bgr = np.stack([bgr, bgr, bgr], axis=-1)
    if rf > 1:  # the pixels the decimation skips must not matter
        junk = rng.integers(0, 256, bgr.shape, dtype=np.uint8)
        keep = np.zeros(bgr.shape[1:3], bool)
        keep[(rows * rf - 1 - rf * np.arange(rows))[:, None], (rf * np.arange(cols))[None, :]] = True
        bgr = np.where(keep[None, :, :, None], bgr, junk)
    d16 = np.repeat(np.repeat(mm, rf, axis=1), rf, axis=2)[:, ::-1]
    return np.ascontiguousarray(bgr), np.ascontiguousarray(d16)


@pytest.mark.gpu
@pytest.mark.parametrize("rf,rows,cols", [(2, 240, 320), (1, 480, 640), (4, 120, 160)])
def test_cuda_convert_frames_bit_exact(sf_mod, oracle_mod, rf, rows, cols):
    rng = np.random.default_rng(rf)
    n = 3
    bgr = rng.integers(0, 256, (n, rows * rf, cols * rf, 3), dtype=np.uint8)
    depth = rng.integers(0, 65536, (n, rows * rf, cols * rf), dtype=np.uint16)
    s = sf_mod.StaticFusionSolver(sf_mod.default_params(rows, cols), max_batch=2)
    got = s.convertFrames(bgr, depth, rf)
    for k in range(n):
        assert same([g[k] for g in got], oracle_mod.convert_frame(bgr[k], depth[k], rf))
    assert same(s.convertFrames(bgr[1], depth[1], rf), oracle_mod.convert_frame(bgr[1], depth[1], rf))  # single image
    with pytest.raises(ValueError):
        s.convertFrames(bgr[:, :-2], depth[:, :-2], rf)
    s.close()


@pytest.mark.gpu
def test_raw_sequence_upload_equals_float_upload(sf_mod, oracle_mod):
    import torch
    rows, cols, rf = 240, 320, 2
    d, c = frames("dynamic", 5, rows, cols)
    bgr, d16 = sequence_as_files(d, c, rf)
    conv = [oracle_mod.convert_frame(bgr[k], d16[k], rf) for k in range(5)]
    fi, fd = np.stack([x[0] for x in conv]), np.stack([x[1] for x in conv])
    s = sf_mod.StaticFusionSolver(sf_mod.default_params(rows, cols), max_batch=4)
    want = s.solve_sequence(fd, fi)
    s.upload_sequence_raw(bgr, d16, rf); s.launch()
    got = s.download(want_images=True)
    assert np.array_equal(got.T, want.T) and np.array_equal(got.labels, want.labels) and np.array_equal(got.b_perpixel, want.b_perpixel)
    gb = torch.from_numpy(bgr).cuda()
    gd = torch.from_numpy(d16.view(np.int16)).cuda().view(torch.uint16)
    s.upload_sequence_raw(gb, gd, rf); s.launch()
    assert np.array_equal(s.download().T, want.T)
    # and the pair results are the oracle's on the converted frames
    o = oracle_mod.Oracle(oracle_mod.driver_params(rows, cols), oracle_mod.ACCUM_EXACT)
    assert np.array_equal(want.T_matrices()[0], o.solve_pair(fd[1], fi[1], fd[0], fi[0]))
    s.close()


@pytest.mark.gpu
def test_cpp_sequence_driver_on_a_tum_folder(sf_mod, oracle_mod, tmp_path):
    """StaticFusion-imagesequenceassoc.cpp's loop through the C++ mirror on a synthetic TUM-format folder; the Python mirror
    replays the same calls and must produce the same `.freiburg` text."""
    cv2 = pytest.importorskip("cv2")
    from staticfusion_b200 import tum_io
    rows, cols, rf, n = 240, 320, 2, 9
    d, c = frames("dynamic", n, rows, cols)
    bgr, d16 = sequence_as_files(d, c, rf)
    seq = str(tmp_path) + "/"
    os.makedirs(seq + "rgb"); os.makedirs(seq + "depth")
    with open(seq + "rgbd_assoc.txt", "w") as f:
        f.write("# synthetic sequence\n")
        for k in range(n):
            t = 100.0 + k / 30.0
            assert cv2.imwrite(seq + f"rgb/{t:.6f}.png", bgr[k]) and cv2.imwrite(seq + f"depth/{t:.6f}.png", d16[k])
            f.write(f"{t:.6f} rgb/{t:.6f}.png {t + 0.01:.6f} depth/{t:.6f}.png\n")
    driver = host_tool("sequence_driver")
    prefix = str(tmp_path / "cpp_run")
    r = subprocess.run([driver, seq, "1000", prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("frame ") == n - 2  # the reference driver starts at index 1 (StaticFusion-imagesequenceassoc.cpp:83)
    # Python mirror, same call order
    ts, fdep, fcol = tum_io.load_assoc(seq, "/rgbd_assoc.txt")
    s = sf_mod.StaticFusionSolver(sf_mod.default_params(rows, cols), max_batch=1)
    tr = tum_io.Trajectory()

    def load(k):
        b, z = tum_io.read_images(fdep[k], fcol[k])
        return s.convertFrames(b, z, rf)

    im = 1
    i0, d0, _, _ = load(im)
    s.bufferSet(im, d0, i0)
    s.depthPrediction, s.intensityPrediction = d0, i0
    im += 1
    s.intensityCurrent, s.depthCurrent, mm, _ = load(im)
    p = sf_mod.default_params(rows, cols); p.kb = 1.05; s.set_params(p)
    s.createImagePyramid(True); s.runSolver(True); s.buildSegmImage(); s.bufferPush(im)
    tr.fuse(s.T_odometry, im)
    p.kb = 1.5; s.set_params(p)
    while im + 1 < n:
        im += 1
        s.depthPrediction, s.intensityPrediction = s.depthCurrent, s.intensityCurrent
        s.intensityCurrent, _, mm, _ = load(im)
        s.depthCurrent = s.getFilteredDepth(mm)
        s.createImagePyramid(True); s.runSolver(True)
        if im - 5 >= 0:
            s.computeResidualsAgainstPreviousImage(im)
        s.buildSegmImage(); s.bufferPush(im)
        tr.fuse(s.T_odometry, im)
    s.close()
    assert open(prefix + ".freiburg").read() == tr.freiburg_text()
    assert len(tr.poseGraph) == n - 2 and np.abs(tr.currPose[:3, 3]).max() > 1e-3
