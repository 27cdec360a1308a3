"""N > 1 path on CPU: world_size-2 gloo processes shard a batch with shard_range, each computes
its slice (the oracle stands in for the kernel: this test is about the partition / timing /
gather plumbing), and the gathered result equals the unsharded one bit for bit."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casclik_b200.sharding import shard_range, max_over_ranks, gather_columns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 23) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bridge import oracle_pinv
    from casclik_b200 import scenarios
    sc = scenarios.get("ur5_track")
    inp = sc.sample(n_total, seed=0)
    lo, hi = shard_range(n_total, rank, world)
    v, _ = oracle_pinv(sc.spec, {"t": inp["t"][lo:hi], "q": inp["q"][:, lo:hi], "y": inp["y"][:, lo:hi]})
    full = gather_columns(torch.from_numpy(np.ascontiguousarray(v)), n_total)
    slow = max_over_ranks(1.0 + rank)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), np.array([slow]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_total = 1001
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bridge import oracle_pinv
    from casclik_b200 import scenarios
    sc = scenarios.get("ur5_track")
    ref, _ = oracle_pinv(sc.spec, sc.sample(n_total, seed=0))
    got = np.load(tmp_path / "gathered.npy")
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert float(np.load(tmp_path / "tmax.npy")[0]) == 2.0
