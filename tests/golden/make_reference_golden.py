"""Generate tests/golden/reference_golden_*.npz by RUNNING THE REFERENCE ITSELF: its own KMeans.cpp,
SegmentationBackground.cpp, StaticFusion.h and the solver part of FrontEnd.cpp, compiled unmodified from
/root/reference against the header shim (oracle/Makefile target `ref`, oracle/ref_shim/sf_ref_shim.h).

The reference tree exists only in the build container, so its outputs are committed here as small fixtures;
tests/test_oracle_golden.py checks the oracle's reference-literal policy against them bit for bit wherever
the suite runs.

    python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference as R  # noqa: E402
from staticfusion_b200 import synth  # noqa: E402

CASES = {
    # name: (scene, first frame, res_factor, twist_old, param overrides)
    "dynamic_160x120": ("dynamic", 20, 4, None, {}),
    "fr1_360_160x120": ("fr1_360", 5, 4, None, {}),
    "walking_xyz_160x120_twist": ("walking_xyz", 12, 4, [0.004, -0.002, 0.001, 0.001, 0.003, -0.002], {}),
    "dynamic_320x240": ("dynamic", 33, 2, None, {}),
    "static_small_80x60_ctor_defaults": ("static_small", 3, 8, None, dict(max_iter_per_level=2, max_iter_irls=10, irls_delta_threshold=1e-6,
                                                                          use_motion_filter=0, kb=1.25, previous_speed_const_weight=0.05,
                                                                          previous_speed_eig_weight=0.5)),
}


def main():
    for name, (scene, t0, rf, twist_old, kw) in CASES.items():
        r = R.Reference(rf, **kw)
        d, c = synth.render_sequence(scene, 2, r.rows, r.cols, start=t0)
        T = r.solve_pair(d[1], c[1], d[0], c[0], twist_old=twist_old)
        ps, sg = r.pose(), r.seg()
        cen, conn = r.kmeans_state()
        levels = r.params.ctf_levels
        out = dict(
            depth_mm=np.round(d * 1000.0).astype(np.uint16), intensity=c, res_factor=np.int32(rf),
            twist_old_in=np.zeros(6, np.float32) if twist_old is None else np.asarray(twist_old, np.float32),
            params=np.array([getattr(r.params, f) for f, _ in R.RefParams._fields_], np.float64),
            T=T, twist_odometry=ps["twist_odometry"], twist_old=ps["twist_old"], twist_level=ps["twist_level"], est_cov=ps["est_cov"],
            b_segm=sg["b_segm"], b_prior=sg["b_prior"], lambda_t_w=sg["lambda_t_w"], b_perpixel=r.b_perpixel(),
            labels0=r.labels(0).astype(np.uint8), labels_coarse=r.labels(levels - 1).astype(np.uint8), kmeans=cen, connectivity=conn,
            depth_pyr_coarse=r.image("depth", levels - 1), intensity_pyr_1=r.image("intensity", 1),
            depth_warped0=r.image("depth_warped", 0), dcu0=r.image("dcu", 0), ddt0=r.image("ddt", 0), weights_d0=r.image("weights_d", 0),
        )
        if rf <= 2:  # keep the committed fixture small: full-resolution float planes only for the 160x120 / 80x60 cases
            for k in ("depth_warped0", "dcu0", "ddt0", "weights_d0", "intensity_pyr_1"):
                del out[k]
        np.savez_compressed(os.path.join(HERE, f"reference_golden_{name}.npz"), **out)
        print(name, r.rows, r.cols, "T[:3,3]", T[:3, 3])


def history():
    """The drivers' steady-state loop over 8 frames with the 5-frame ring buffers (StaticFusion-datasets.cpp:109-184):
    frames 5..7 run computeResidualsAgainstPreviousImage (FrontEnd.cpp:896) before buildSegmImage."""
    for name, scene, t0, rf in (("history_dynamic_160x120", "dynamic", 40, 4), ("history_walking_xyz_80x60", "walking_xyz", 8, 8)):
        r = R.Reference(rf)
        n = 9
        d, c = synth.render_sequence(scene, n, r.rows, r.cols, start=t0)
        r.buffer_set(0, d[0], c[0])
        Ts, pcs, bps, wref = [], [], [], []
        for t in range(1, n):
            Ts.append(r.track_frame(t, d[t], c[t], d[t - 1], c[t - 1]))
            pcs.append(r.per_cluster_average_residual())
            bps.append(r.b_perpixel())
        np.savez_compressed(
            os.path.join(HERE, f"reference_{name}.npz"), depth_mm=np.round(d * 1000.0).astype(np.uint16), intensity=c,
            res_factor=np.int32(rf), T=np.stack(Ts), per_cluster=np.stack(pcs), b_perpixel=np.stack(bps),
            depth_warped_ref=r.residual_image("depth_warped_ref"), cumulative=r.residual_image("cumulative"))
        print(name, "flipped clusters in last frame:", int((pcs[-1] < 0.017).sum()))


if __name__ == "__main__":
    main()
    history()
