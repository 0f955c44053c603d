"""Frame sharding over GPUs: one process per GPU, contiguous blocks of frame pairs per rank.

The reference is strictly sequential frame-to-model SLAM; what shards is its frame-to-frame form
(prediction := previous raw frame, ``StaticFusion-datasets.cpp:109-144``) where pair (t-1, t) is an
independent solve.  There is no data-path collective: each rank solves its own pairs and ONE
all-gather of the small per-pair result rows (pose, twist, b_segm, counters; 48 floats) follows the
whole batch.  The global trajectory is the prefix product of the gathered increments
(``cam_pose = cam_pose + pose_aux``, ``FrontEnd.cpp:1134-1137``), done on the host in float64.
"""
from __future__ import annotations

import numpy as np

ROW = 16 + 6 + 24 + 2  # T (column-major 4x4), twist_old, b_segm, irls_iters, status


def shard_pairs(n_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the pairs rank `rank` solves: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_frames(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last] inclusive frames rank `rank` needs for its pairs of a sequence (one halo frame)."""
    s, e = shard_pairs(n_frames - 1, rank, world)
    return s, e  # pairs s..e-1 use frames s..e


def pack_rows(result) -> np.ndarray:
    n = result.T.shape[0]
    rows = np.zeros((n, ROW), np.float32)
    rows[:, 0:16] = result.T
    rows[:, 16:22] = result.twist_old
    rows[:, 22:46] = result.b_segm
    rows[:, 46] = result.irls_iters
    rows[:, 47] = result.status
    return rows


def unpack_rows(rows: np.ndarray) -> dict:
    return dict(T=rows[:, 0:16].copy(), twist_old=rows[:, 16:22].copy(), b_segm=rows[:, 22:46].copy(),
                irls_iters=rows[:, 46].astype(np.int32), status=rows[:, 47].astype(np.int32))


def gather_rows(local_rows: np.ndarray, n_pairs: int, device=None) -> np.ndarray:
    """All-gather the per-pair rows of every rank into the global (n_pairs, ROW) table (same on all ranks).

    Uses the initialised ``torch.distributed`` group: NCCL when `device` is a CUDA device (the rows are
    staged through a device tensor), gloo on CPU.  Single process / no group: returns the input.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local_rows.shape[0] == n_pairs
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_pairs(n_pairs, r, world) for r in range(world)]
    cap = max(e - s for s, e in counts)
    buf = torch.zeros((cap, ROW), dtype=torch.float32, device=device if device is not None else "cpu")
    s, e = counts[rank]
    assert local_rows.shape[0] == e - s
    if e > s:
        buf[: e - s] = torch.from_numpy(np.ascontiguousarray(local_rows)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    table = np.zeros((n_pairs, ROW), np.float32)
    for r, (rs, re) in enumerate(counts):
        if re > rs:
            table[rs:re] = out[r][: re - rs].cpu().numpy()
    return table


def rows_from_device_layout(rows: np.ndarray) -> np.ndarray:
    """Device result rows (sf_result_rows_device: T row-major, counters as int32 bit patterns) -> the host row layout of
    pack_rows (T column-major like Eigen::Matrix4f, counters as floats)."""
    out = rows.astype(np.float32, copy=True)
    out[:, 0:16] = rows[:, 0:16].reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1, 16)
    out[:, 46:48] = rows[:, 46:48].copy().view(np.int32).astype(np.float32)
    return out


class DeviceRowGather:
    """All-gather of the per-pair result rows that never leaves the device (NCCL over NVLink).

    ``enqueue(solver)`` is called right after ``solver.launch()``: a 48-float-per-pair copy behind the solve on the solver's
    own stream puts the rows into a staging slot, an event hands them to a side stream and ``all_gather_into_tensor`` fills
    the slot's pre-allocated table there.  Nothing blocks the host, so the next batch is enqueued at once and the collective
    overlaps it; ``table(slot)`` is the only synchronising call (one download of the gathered table).  Ranks hold the same
    number of rows per slot (`rows_per_rank`, shorter shards are zero-padded)."""

    def __init__(self, rows_per_rank: int, device, slots: int = 3):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.cap, self.device, self.slots = rows_per_rank, device, slots
        self.stage = [torch.zeros((rows_per_rank, ROW), dtype=torch.float32, device=device) for _ in range(slots)]
        self.tables = [torch.zeros((self.world * rows_per_rank, ROW), dtype=torch.float32, device=device) for _ in range(slots)]
        self.side = torch.cuda.Stream(device=device)
        self.done = [None] * slots
        self.count = 0

    def enqueue(self, solver, first: int = 0) -> int:
        """`first`: leading rows of the solver's batch to leave out (the history halo pairs of a sharded sequence)."""
        self.stage_rows(solver, first, None, 0)
        return self.launch_gather()

    def stage_rows(self, solver, first: int, n, at: int):
        """Copy rows [first, first + n) of the solver's last batch into rows [at, at + n) of the current slot's staging buffer,
        behind the solve on the solver's stream (several solver contexts may each contribute a part of one table)."""
        torch = self.torch
        slot = self.count % self.slots
        st = torch.cuda.ExternalStream(solver.stream, device=self.device)
        rows = solver.result_rows_device()
        rows = rows[first:] if n is None else rows[first:first + n]
        if self.done[slot] is not None:
            st.wait_event(self.done[slot])  # the slot's previous table has been produced (and, by contract, read)
        with torch.cuda.stream(st):
            self.stage[slot][at: at + rows.shape[0]].copy_(rows, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(st)
        self.side.wait_event(ready)

    def launch_gather(self) -> int:
        """One all-gather of the current slot (after every stage_rows of it); the same number of calls on every rank."""
        torch = self.torch
        slot = self.count % self.slots
        self.count += 1
        with torch.cuda.stream(self.side):
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.tables[slot], self.stage[slot])
            else:
                self.tables[slot].copy_(self.stage[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[slot] = ev
        return slot

    def table(self, slot: int, counts=None) -> np.ndarray:
        """The gathered (world * rows_per_rank, 48) table of `slot` in the host row layout; `counts` = rows each rank really
        holds (padding removed)."""
        self.done[slot].synchronize()
        t = self.tables[slot].cpu().numpy()
        if counts is not None:
            t = np.concatenate([t[r * self.cap: r * self.cap + n] for r, n in enumerate(counts)], axis=0)
        return rows_from_device_layout(t)


def compose_trajectory(T_colmajor: np.ndarray, start: np.ndarray | None = None) -> np.ndarray:
    """Prefix product of the increments: pose_k = pose_{k-1} @ T_k (float64), shape (n+1, 4, 4)."""
    n = T_colmajor.shape[0]
    poses = np.zeros((n + 1, 4, 4), np.float64)
    poses[0] = np.eye(4) if start is None else start
    for k in range(n):
        poses[k + 1] = poses[k] @ T_colmajor[k].reshape(4, 4).T.astype(np.float64)
    return poses


HISTORY_HALO = 4  # computeResidualsAgainstPreviousImage looks five frames back (FrontEnd.cpp:898-909)


def solve_sequence_sharded(solver, depth, inten, device=None, want_images=False, history=False):
    """Solve the pairs of a frame sequence owned by this rank and gather every rank's result rows.

    `depth` / `inten` hold the WHOLE sequence (n_frames, rows, cols) on every rank (synthetic data is
    generated locally); only the rank's frame block (plus a 4-pair halo when `history` is on) is uploaded.
    Returns (global table dict, local BatchResult).
    """
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n_frames = int(depth.shape[0])
    f0, f1 = shard_frames(n_frames, rank, world)
    # with the 5-frame history every rank re-solves the last four pairs of its predecessor instead of exchanging them
    halo = min(HISTORY_HALO, f0) if history else 0
    local = solver.solve_sequence(depth[f0 - halo:f1 + 1], inten[f0 - halo:f1 + 1], want_images=want_images, history=history, halo=halo)
    if device is not None and str(device).startswith("cuda"):
        # rows stay on the device: staging copy behind the solve, NCCL all-gather on a side stream, one download
        counts = [shard_pairs(n_frames - 1, r, world)[1] - shard_pairs(n_frames - 1, r, world)[0] for r in range(world)]
        g = DeviceRowGather(max(counts), device, slots=1)
        table = g.table(g.enqueue(solver, first=halo), counts)
    else:
        table = gather_rows(pack_rows(local), n_frames - 1, device=device)
    return unpack_rows(table), local
