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


def oracle_sequence(O, params, depth, inten, history=False, accum=None, want_images=True):
    """The batched sequence semantics restated with the oracle: pair k = frames k, k+1 solved on its own
    (twist_odometry_old = 0); with `history`, pair k >= 4 also runs computeResidualsAgainstPreviousImage(k+1) against
    frame k-4 through the increments of pairs k-4..k (ring buffers filled as the drivers would have,
    StaticFusion-datasets.cpp:182-184) before buildSegmImage.  Returns a dict of stacked outputs."""
    accum = O.ACCUM_EXACT if accum is None else accum
    n = depth.shape[0] - 1
    out = dict(T=[], b_segm=[], per_cluster=[], b_perpixel=[], labels=[], irls=[], status=[])
    Ts = []
    for k in range(n):
        o = O.Oracle(params, accum)
        o.set_current(depth[k + 1], inten[k + 1])
        o.set_prediction(depth[k], inten[k])
        o.set_twist_old(np.zeros(6, np.float32))
        o.create_image_pyramid(True)
        o.run_solver(True)
        Ts.append(o.T())
        if history and k >= 4:
            index = k + 1
            o.buffer_set(index % 5, depth[k - 4], inten[k - 4])
            for i in range(index - 4, index):
                o.buffer_set(i % 5, depth[i], inten[i], Ts[i - 1])
            o.compute_residuals_against_previous_image(index)
        o.build_segm_image()
        out["T"].append(o.T()); out["b_segm"].append(o.b_segm()); out["per_cluster"].append(o.per_cluster_average_residual())
        out["irls"].append(o.total_irls()); out["status"].append(o.status())
        if want_images:
            out["b_perpixel"].append(o.b_perpixel()); out["labels"].append(o.labels(0))
    return {k: np.stack(v) for k, v in out.items() if len(v)}


def _oracle_job(args):
    from oracle import oracle as O
    fields, accum, dc, ic, dp, ip = args
    o = O.Oracle(O.Params(**fields), accum)
    T = o.solve_pair(dc, ic, dp, ip)
    return dict(T=T.copy(), labels=o.labels(0), b_perpixel=o.b_perpixel(), b_segm=o.b_segm(), twist_old=o.twists()[1],
                irls=o.total_irls(), status=o.status())


def oracle_pairs(O, params, jobs, accum=None, workers=None):
    """Solve independent pairs [(depth_cur, inten_cur, depth_pred, inten_pred), ...] with the oracle on the host cores
    (twist_odometry_old = 0, as the batched sequence path does); returns one dict of outputs per pair."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    accum = O.ACCUM_EXACT if accum is None else accum
    fields = {name: getattr(params, name) for name, _ in O.Params._fields_}
    args = [(fields, accum) + tuple(j) for j in jobs]
    workers = workers or min(len(args), os.cpu_count() or 1, 32)
    if workers <= 1:
        return [_oracle_job(a) for a in args]
    with ProcessPoolExecutor(workers) as ex:
        return list(ex.map(_oracle_job, args, chunksize=1))
