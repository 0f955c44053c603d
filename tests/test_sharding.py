"""Multi-GPU host logic on CPU: world_size-2 gloo run of the frame sharding + result gather.

The solve itself needs a B200; here the per-rank solve is stood in by the CPU oracle on tiny frames so
that the partitioning, halo frames, all-gather and trajectory composition are exercised end to end.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition_properties():
    from staticfusion_b200.sharding import shard_frames, shard_pairs
    for n in (1, 2, 7, 8, 125, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_pairs(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
            for r in range(world):
                f0, f1 = shard_frames(n + 1, r, world)
                assert (f0, f1) == spans[r]  # pairs s..e-1 need frames s..e (one halo frame)


def test_compose_trajectory_is_prefix_product():
    from staticfusion_b200.sharding import compose_trajectory
    rng = np.random.default_rng(0)
    Ts = []
    for _ in range(5):
        A = np.eye(4)
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        A[:3, :3] = q
        A[:3, 3] = rng.standard_normal(3)
        Ts.append(A)
    colmajor = np.stack([T.T.reshape(16) for T in Ts]).astype(np.float32)
    poses = compose_trajectory(colmajor)
    ref = np.eye(4)
    for k, T in enumerate(Ts):
        ref = ref @ T.astype(np.float32).astype(np.float64)
        assert np.allclose(poses[k + 1], ref, atol=1e-6)


def _worker(rank, world, port, n_frames, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from staticfusion_b200 import sharding, synth
    from staticfusion_b200.solver import BatchResult

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, cols = 60, 80
    d, c = synth.render_sequence("dynamic", n_frames, rows, cols, start=30)

    import common

    class OracleSolver:  # same solve_sequence contract as StaticFusionSolver
        def solve_sequence(self, depth, inten, twist_old=None, want_images=False, history=False, halo=0):
            n = depth.shape[0] - 1 - halo
            r = BatchResult(n, rows, cols, False)
            o = common.oracle_sequence(O, O.driver_params(rows, cols, ctf_levels=3), depth, inten, history=history, want_images=False)
            r.T[:] = o["T"][halo:].transpose(0, 2, 1).reshape(-1, 16)
            r.twist_old[:] = 0
            r.b_segm[:] = o["b_segm"][halo:]
            r.irls_iters[:] = o["irls"][halo:]
            r.status[:] = o["status"][halo:]
            r.per_cluster_residual[:] = o["per_cluster"][halo:]
            return r

    table, local = sharding.solve_sequence_sharded(OracleSolver(), d, c, history=True)
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), **table, n_local=local.T.shape[0], per_cluster=local.per_cluster_residual)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from staticfusion_b200 import synth

    n_frames, world = 12, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert int(r0["n_local"]) + int(r1["n_local"]) == n_frames - 1
    for k in ("T", "twist_old", "b_segm", "irls_iters", "status"):
        assert np.array_equal(r0[k], r1[k])  # every rank ends with the same global table
    # single-process reference
    rows, cols = 60, 80
    d, c = synth.render_sequence("dynamic", n_frames, rows, cols, start=30)
    for k in range(n_frames - 1):
        o = O.Oracle(O.driver_params(rows, cols, ctf_levels=3), O.ACCUM_EXACT)
        T = o.solve_pair(d[k + 1], c[k + 1], d[k], c[k])
        assert np.array_equal(r0["T"][k], T.T.reshape(16))
        assert r0["irls_iters"][k] == o.total_irls()
    # 5-frame history across the shard boundary: rank 1 re-solves four halo pairs, so its first pairs see the same
    # history as in a single-process run
    import common
    full = common.oracle_sequence(O, O.driver_params(rows, cols, ctf_levels=3), d, c, history=True, want_images=False)
    n0 = int(r0["n_local"])
    assert np.array_equal(r0["per_cluster"], full["per_cluster"][:n0], equal_nan=True)
    assert np.array_equal(r1["per_cluster"], full["per_cluster"][n0:], equal_nan=True)
    assert np.isnan(full["per_cluster"][:4]).all() and not np.isnan(full["per_cluster"][4:]).all()
