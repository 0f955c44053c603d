"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU, NCCL): one frame sequence sharded
over the ranks with the 4-pair history halo (StaticFusion-datasets.cpp:109-184 semantics, frame-to-frame form), result rows
gathered on the device; rank 0 compares the gathered table with the whole sequence solved on its own GPU, bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import staticfusion_b200 as sf
    from staticfusion_b200 import sharding, synth

    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 42
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rows, cols = 240, 320
    d, c = synth.render_sequence("walking_xyz", n_frames, rows, cols, start=0)
    p = sf.default_params(rows, cols)  # reference default 5 levels (BASELINE config 4)
    p0, p1 = sharding.shard_pairs(n_frames - 1, rank, world)
    s = sf.StaticFusionSolver(p, device=local, max_batch=p1 - p0 + sharding.HISTORY_HALO)
    table, localres = sharding.solve_sequence_sharded(s, d, c, device=dev, want_images=True, history=True)
    # every rank holds the same table
    t = torch.from_numpy(np.concatenate([table["T"], table["twist_old"], table["b_segm"]], axis=1)).to(dev)
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), "ranks disagree on the gathered table"
    s.close()
    if rank == 0:
        whole = sf.StaticFusionSolver(p, device=local, max_batch=n_frames - 1)
        r = whole.solve_sequence(d, c, history=True)
        assert np.array_equal(table["T"], r.T), "poses differ between the sharded and the single-GPU run"
        assert np.array_equal(table["twist_old"], r.twist_old) and np.array_equal(table["b_segm"], r.b_segm)
        assert np.array_equal(table["irls_iters"], r.irls_iters) and np.array_equal(table["status"], r.status)
        # this rank's own pairs: per-pixel outputs and the 5-frame residuals (the halo reproduces the history across the cut)
        assert np.array_equal(localres.b_perpixel, r.b_perpixel[p0:p1]) and np.array_equal(localres.labels, r.labels[p0:p1])
        assert np.array_equal(localres.per_cluster_residual, r.per_cluster_residual[p0:p1], equal_nan=True)
        whole.close()
        print(f"MGPU_OK world={world} pairs={n_frames - 1}")
    # the last rank checks its per-pixel outputs too (its first pairs sit right behind a cut)
    if rank == world - 1 and world > 1:
        whole = sf.StaticFusionSolver(p, device=local, max_batch=n_frames - 1)
        r = whole.solve_sequence(d, c, history=True)
        assert np.array_equal(localres.b_perpixel, r.b_perpixel[p0:p1])
        assert np.array_equal(localres.per_cluster_residual, r.per_cluster_residual[p0:p1], equal_nan=True)
        assert np.isfinite(localres.per_cluster_residual).any()
        whole.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
