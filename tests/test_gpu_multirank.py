"""Two ranks on two GPUs (skipped on a single-GPU box): index shards + the count exchange, through NVLink peer
stores (k4_exchange_peer) and through NCCL, against the single-rank oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from __graft_entry__ import REPO, load_package

cc = load_package()
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode):
    use_nccl = mode == "nccl"
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch.distributed as dist
    import harness as H
    from __graft_entry__ import load_package
    cc = load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cols = H.concat_cols(H.shell_fixture(600_001, seed=41), H.shell_fixture(200_000, seed=42, octant=False))
        cols["id"] = np.arange(cols["x"].shape[0], dtype=np.int64)
        n = cols["x"].shape[0]
        lo, hi = cc.shard_range(n, world, rank)
        ctx = cc.Context(rank)
        if use_nccl:
            ident = [cc.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ident, src=0)
            ctx.comm_init(world, rank, ident[0])
        else:
            handles = [None] * world
            dist.all_gather_object(handles, ctx.comm_peer_handle(world, rank))
            ctx.comm_peer_attach(handles)
            ctx.set_option("exchange", mode)
        dist.barrier()
        ctx.upload(H.slice_cols(cols, lo, hi), index_base=lo)
        tc, pc = cc.cli_bounds(70, 15), cc.cli_bounds(40, 15)
        shard_counts = [H.oracle_cutout_const(H.slice_cols(cols, *cc.shard_range(n, world, r)), tc, pc)["id"].shape[0]
                        for r in range(world)]
        for rep in range(3):  # several epochs through the same mailboxes
            ctx.cutout_const(tc, pc)
            allc, allr, off = ctx.exchange_counts()
            assert allc[:, 0].tolist() == shard_counts, (allc, shard_counts)
            assert int(off[0]) == sum(shard_counts[:rank])
        halos = [((150, 20, 12), 120.0), ((30, 170, 100), 60.0), ((100, 100, 100), 240.0)]
        ctx.cutout_halo([cc.halo_setup(p, b).halo for p, b in halos], n_limit_global=n)
        allc, allr, off = ctx.exchange_counts()
        tot = 0
        for h, (p, b) in enumerate(halos):
            want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(p, b))
            assert int(allc[:, h].sum()) == want["id"].shape[0] and int(allr[:, h].sum()) == want_rough
            assert int(off[h]) == int(allc[:rank, h].sum())
            # this rank's survivors are exactly the oracle's survivors that live in its shard (same relative order)
            mine = ctx.fetch(h, names=("id", "theta"))
            sel = (want["id"] >= lo) & (want["id"] < hi)
            assert mine["id"].tobytes() == want["id"][sel].tobytes() and mine["theta"].tobytes() == want["theta"][sel].tobytes()
            tot += want["id"].shape[0]
        assert tot > 1000
        # cluster-sized cutouts, repeatedly: from the second one on the single-CTA kernel orders + gathers, and with
        # exchange=eager it also carries the count exchange (posted first, collected after the rows are written)
        small = [((150, 20, 12), 25.0), ((100, 100, 100), 30.0)]
        wants = [H.oracle_cutout_halo(cols, H.oracle_halo_setup(p, b)) for p, b in small]
        for rep in range(4):
            ctx.cutout_halo([cc.halo_setup(p, b).halo for p, b in small], n_limit_global=n)
            allc, allr, off = ctx.exchange_counts()
            for h, (want, want_rough) in enumerate(wants):
                assert int(allc[:, h].sum()) == want["id"].shape[0] and int(allr[:, h].sum()) == want_rough, (rep, h)
                assert int(off[h]) == int(allc[:rank, h].sum())
                mine = ctx.fetch(h, names=("id", "phi"))
                sel = (want["id"] >= lo) & (want["id"] < hi)
                assert mine["id"].tobytes() == want["id"][sel].tobytes() and mine["phi"].tobytes() == want["phi"][sel].tobytes()
        dist.barrier()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "eager", "nccl"])
def test_two_gpus_count_exchange(mode):
    if cc.lib().cc_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), mode), nprocs=2, join=True)
