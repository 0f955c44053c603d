"""Generates tests/golden/reference_loader_*.npz from the REFERENCE's own StaticFusion::loadImageFromSequenceAssoc and
StaticFusion::loadAssoc (FrontEnd.cpp:183-254, compiled unmodified by `make -C oracle ref`; needs /root/reference).
Inputs are regenerated from the seed by the test, only the reference's outputs are stored.

    python tests/golden/make_loader_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference as R  # noqa: E402

ASSOC_TEXT = """# colour and depth, associated
# ts_rgb rgb ts_depth depth
1311868164.363181 rgb/1311868164.363181.png 1311868164.338541 depth/1311868164.338541.png

1311868164.399026 rgb/1311868164.399026.png 1311868164.373557 depth/1311868164.373557.png
1311868164.430940\trgb/1311868164.430940.png\t1311868164.407784\tdepth/1311868164.407784.png
#1311868164.463055 rgb/skipped.png 1311868164.437021 depth/skipped.png
1311868164.463055 rgb/1311868164.463055.png 1311868164.437021 depth/1311868164.437021.png   trailing tokens are ignored
this line is malformed and ends the parse
1311868164.531025 rgb/never.png 1311868164.508285 depth/never.png
"""


def loader_inputs(seed):
    """640x480 decoded images: smooth colour field + noise, millimetre depth with holes and saturated values."""
    rng = np.random.default_rng(seed)
    v, u = np.mgrid[0:480, 0:640]
    base = np.stack([127 + 120 * np.sin(u / 37.0 + c) * np.cos(v / 23.0 - c) for c in range(3)], axis=-1)
    bgr = np.clip(base + rng.integers(-6, 7, (480, 640, 3)), 0, 255).astype(np.uint8)
    depth = (1500 + 900 * np.sin(u / 91.0) + 400 * np.cos(v / 57.0) + rng.integers(-3, 4, (480, 640))).astype(np.uint16)
    depth[rng.random((480, 640)) < 0.05] = 0
    depth[0:3, 0:5] = 65535
    bgr[0, 0] = (255, 255, 255); bgr[0, 1] = (0, 0, 0); bgr[479, 639] = (1, 2, 3)
    return bgr, depth


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for rf in (4, 8):
        bgr, depth = loader_inputs(1000 + rf)
        r = R.Reference(rf)
        inten, dep, mm, col = r.load_image_from_sequence_assoc(bgr, depth, rf)
        np.savez_compressed(os.path.join(out, f"reference_loader_rf{rf}.npz"), seed=1000 + rf, res_factor=rf, intensity=inten, depth=dep,
                            depth_mm=mm, color_full=col)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "rgbd_assoc.txt"), "w").write(ASSOC_TEXT)
        ts, fd, fc = R.Reference.load_assoc(d, "/rgbd_assoc.txt")
    np.savez_compressed(os.path.join(out, "reference_load_assoc.npz"), timestamps=np.array(ts, np.float64),
                        files_depth=np.array([x[len(d):] for x in fd]), files_color=np.array([x[len(d):] for x in fc]))
    print("ok", len(ts))


if __name__ == "__main__":
    main()
