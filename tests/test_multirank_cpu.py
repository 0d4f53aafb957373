"""world_size-2 (gloo, CPU) test of the multi-rank plan: contiguous index shards (cc_shard_range), count exchange ->
exclusive-scan file offsets, and the (theta_u, index) merge of use case 2 (cc_merge_order) reproduce the
single-rank output exactly.  No GPU here, so each rank's shard cutout is computed by the oracle standing in for
the kernel; everything around it is the product's host logic."""
import os
import socket
import sys

import numpy as np
import pytest

import harness as H
from __graft_entry__ import REPO, load_package

cc = load_package()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch
    import torch.distributed as dist
    import harness as H  # noqa
    from __graft_entry__ import load_package
    cc = load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cols = H.concat_cols(H.shell_fixture(300_001, seed=41), H.shell_fixture(100_000, seed=42, octant=False))
        cols["id"] = np.arange(cols["x"].shape[0], dtype=np.int64)  # id == global index
        n = cols["x"].shape[0]
        lo, hi = cc.shard_range(n, world, rank)
        mine = H.slice_cols(cols, lo, hi)

        # ---- use case 1: blocks at exclusive-scan offsets == single-rank file
        tc, pc = cc.cli_bounds(70, 15), cc.cli_bounds(40, 15)
        part = H.oracle_cutout_const(mine, tc, pc)
        k = torch.tensor([part["id"].shape[0]], dtype=torch.int64)
        allk = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allk, k)
        counts = np.array([int(t.item()) for t in allk])
        offset = int(counts[:rank].sum())
        for nme in ("id", "theta", "x"):
            path = os.path.join(tmpdir, nme + ".bin")
            fd = os.open(path, os.O_CREAT | os.O_WRONLY, 0o644)
            os.pwrite(fd, part[nme].tobytes(), offset * part[nme].dtype.itemsize)
            os.close(fd)
        dist.barrier()
        if rank == 0:
            want = H.oracle_cutout_const(cols, tc, pc)
            assert counts.sum() == want["id"].shape[0] > 1000
            for nme in ("id", "theta", "x"):
                got = np.fromfile(os.path.join(tmpdir, nme + ".bin"), dtype=want[nme].dtype)
                assert got.tobytes() == want[nme].tobytes(), nme

        # ---- use case 2: merge of the per-rank (theta_u, index)-ordered lists == single-rank order
        halo = H.oracle_halo_setup((150, 20, 12), 240.0)
        part, rough = H.oracle_cutout_halo(mine, halo)
        th_u, _ = H.oracle_theta_phi(part["x"], part["y"], part["z"], guards=True)
        payload = dict(id=part["id"], theta=part["theta"], theta_u=th_u, index=part["id"].copy(), rough=rough)
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
        if rank == 0:
            want, want_rough = H.oracle_cutout_halo(cols, halo)
            assert sum(g["rough"] for g in gathered) == want_rough
            src_rank, src_row = cc.merge_order([g["theta_u"] for g in gathered], [g["index"] for g in gathered])
            ids = np.array([gathered[r]["id"][j] for r, j in zip(src_rank, src_row)], dtype=np.int64)
            th = np.array([gathered[r]["theta"][j] for r, j in zip(src_rank, src_row)], dtype=np.float32)
            assert want["id"].shape[0] > 500 and len(set(src_rank.tolist())) == world
            assert ids.tobytes() == want["id"].tobytes() and th.tobytes() == want["theta"].tobytes()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_plan_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


def test_shard_ranges_tile_the_step():
    for n in (0, 1, 7, 1000, 2 ** 24 + 5, 1_000_000_007, 4_000_000_000):
        for world in (1, 2, 3, 4, 8):
            edges = [cc.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_merge_order_ties_and_empty_ranks():
    t = [np.array([1.0, 1.0, 2.0], np.float32), np.array([], np.float32), np.array([1.0, 2.0, 2.0], np.float32)]
    i = [np.array([5, 9, 1], np.int64), np.array([], np.int64), np.array([7, 0, 3], np.int64)]
    r, j = cc.merge_order(t, i)
    merged = [(float(t[a][b]), int(i[a][b])) for a, b in zip(r, j)]
    assert merged == sorted(merged) == [(1.0, 5), (1.0, 7), (1.0, 9), (2.0, 0), (2.0, 1), (2.0, 3)]


def test_merge_order_of_a_halo_without_survivors():
    """a halo whose cutout is empty on every rank (a small box in a sparse step): nothing to merge, and the empty
    output arrays a caller passes may have NULL data pointers (std::vector<>::data() of an empty vector)"""
    import ctypes as C
    r, j = cc.merge_order([np.zeros(0, np.float32)] * 3, [np.zeros(0, np.int64)] * 3)
    assert r.shape == (0,) and j.shape == (0,)
    counts = np.zeros(2, dtype=np.int64)
    null2 = (C.c_void_p * 2)(None, None)
    assert cc.lib().cc_merge_order(2, counts.ctypes.data_as(C.c_void_p), null2, null2, None, None) == 0
    counts[0] = 1  # survivors but no key columns: refused with a message
    assert cc.lib().cc_merge_order(2, counts.ctypes.data_as(C.c_void_p), null2, null2, None, None) != 0
    assert b"cc_merge_order" in cc.lib().cc_last_error()
