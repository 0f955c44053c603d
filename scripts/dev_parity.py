"""Developer smoke: GPU solver vs CPU oracle on a few synthetic pairs (prints diffs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import staticfusion_b200 as sf
from staticfusion_b200 import synth
from oracle import oracle as O

def rot_angle(R):
    return float(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))

def main():
    rows, cols = 240, 320
    scene = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
    nf = 4
    ds, cs = synth.render_sequence(scene, nf + 1, rows, cols, start=10)
    p = sf.default_params(rows, cols)
    s = sf.StaticFusionSolver(p, max_batch=nf, trace=True)
    t0 = time.time()
    r = s.solve_sequence(ds, cs)
    print("gpu solve_sequence", time.time() - t0, "launches", s.last_launch_count, "status", r.status, "iters", r.irls_iters)
    op = O.driver_params(rows, cols)
    for k in range(nf):
        o = O.Oracle(op, O.ACCUM_EXACT)
        To = o.solve_pair(ds[k + 1], cs[k + 1], ds[k], cs[k])
        Tg = r.T_matrices()[k]
        dT = np.linalg.inv(To.astype(np.float64)) @ Tg.astype(np.float64)
        print(f"pair {k}: dtrans {np.abs(To[:3,3]-Tg[:3,3]).max():.3e} drot {rot_angle(dT[:3,:3]):.3e} iters gpu {r.irls_iters[k]} orc {o.total_irls()}")
        for L in range(p.ctf_levels):
            for nm in ("depth", "intensity", "depth_pred", "intensity_pred"):
                g = s.debug_plane(nm, k, L); c = o.image(nm, L)
                if not np.array_equal(g, c): print("   MISMATCH", nm, L, np.abs(g - c).max(), (g != c).sum())
            lg = s.debug_labels(k, L); lo = o.labels(L)
            if not np.array_equal(lg, lo): print("   LABEL MISMATCH level", L, (lg != lo).sum())
        cen, conn = s.debug_kmeans(k)
        print("   kmeans equal", np.array_equal(cen, o.kmeans_centres()), "conn equal", np.array_equal(conn, o.connectivity()))
        print("   b_segm maxdiff", np.abs(r.b_segm[k] - o.b_segm()).max(), "mask equal", np.array_equal(r.b_perpixel[k] > 0.5, o.b_perpixel() > 0.5),
              "bpp maxdiff", np.abs(r.b_perpixel[k] - o.b_perpixel()).max())
        tg = s.debug_trace(k); to = o.trace()
        for st in range(tg.shape[0]):
            if to[st, 0] != tg[st, 0]: print("   step executed mismatch", st, to[st, 0], tg[st, 0]); continue
            if not to[st, 0]: continue
            hd = np.abs(tg[st, :85] - to[st, :85])
            print(f"   step {st}: N {int(tg[st,3])}/{int(to[st,3])} it {int(tg[st,4])}/{int(to[st,4])} hdr maxdiff {hd.max():.2e} @ {hd.argmax()}", end="")
            nit = int(to[st, 4])
            ig = tg[st, 96:96 + 34 * nit].reshape(nit, 34); io = to[st, 96:96 + 34 * nit].reshape(nit, 34)
            print(f" var maxdiff {np.abs(ig[:, :6]-io[:, :6]).max():.2e} b {np.abs(ig[:, 6:30]-io[:, 6:30]).max():.2e} aver {np.abs(ig[:,30]-io[:,30]).max():.2e} ressq rel {np.abs(ig[:,32]/io[:,32]-1).max():.2e}")

if __name__ == "__main__":
    main()
