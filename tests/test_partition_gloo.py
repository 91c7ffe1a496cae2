"""The N > 1 host-side logic, run as two real processes over torch.distributed (gloo, CPU):
sample partition, ray sharding, the NCCL-id style broadcast, the max-over-ranks timing rule and the
film reduction (sum of RGBW accumulators == the single-process film)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spica_b200 import partition, scenes


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _fake_film(sample_ids, w=8, h=6):
    """Deterministic stand-in for a rendered film: every sample adds a pattern and weight 1."""
    film = np.zeros((h, w, 4), dtype=np.float64)
    for s in sample_ids:
        rng = np.random.default_rng(1000 + s)
        film[..., :3] += rng.random((h, w, 3))
        film[..., 3] += 1.0
    return film


def _worker(rank, world, port, spp, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the unique id is created on rank 0 and shipped to the others (bench.py does this for NCCL)
    ids = [b"id-from-rank-0" if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == b"id-from-rank-0"
    mine = partition.sample_indices(spp, rank, world)
    film = torch.from_numpy(_fake_film(mine))
    dist.all_reduce(film, op=dist.ReduceOp.SUM)                 # K7: one sum over the film
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)  # rank-local device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                    # bench.py: max over ranks
    start, count = partition.ray_shard(1000, rank)
    rays = scenes.incoherent_rays(count, [-1, -1, -1], [1, 1, 1], seed=2, start=start)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, float(rays[:, :3].sum())))
    if rank == 0:
        np.save(out, film.numpy())
        with open(out + ".txt", "w") as f:
            f.write(repr((gathered, float(t.item()))))
    dist.destroy_process_group()


@pytest.mark.parametrize("spp", [7, 8])
def test_two_rank_partition_and_film_reduce(tmp_path, spp):
    world = 2
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    film = np.load(out)
    assert np.allclose(film, _fake_film(range(spp)))            # union of the ranks == single process
    assert np.array_equal(film[..., 3], np.full(film.shape[:2], float(spp)))
    gathered, tmax = eval(open(out + ".txt").read())
    ids = sorted(gathered[0][0] + gathered[1][0])
    assert ids == list(range(spp))                              # every sample exactly once
    assert tmax == pytest.approx(0.020)
    # ray shards are disjoint ranges of one generator: concatenation == one big call
    full = scenes.incoherent_rays(2000, [-1, -1, -1], [1, 1, 1], seed=2)
    assert gathered[0][1] == pytest.approx(float(full[:1000, :3].sum()))
    assert gathered[1][1] == pytest.approx(float(full[1000:, :3].sum()))


def test_partition_edges():
    assert partition.sample_partition(5, 3, 8) == (3, 1, 8)
    assert partition.sample_partition(3, 5, 8) == (5, 0, 8)
    assert sum(partition.sample_partition(1024, r, 8)[1] for r in range(8)) == 1024
    with pytest.raises(ValueError):
        partition.sample_partition(4, 2, 2)
