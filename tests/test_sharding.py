"""Exposure-wise sharding: host logic under a 2-rank gloo group on CPU."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from wayne_b200 import sharding


def test_partition_covers_every_exposure_once():
    for n, g in ((121, 8), (128, 4), (5, 8), (0, 2), (1024, 1)):
        seen = sorted(i for r in range(g) for i in sharding.shard_indices(n, g, r))
        assert seen == list(range(n))
        sizes = [len(sharding.shard_indices(n, g, r)) for r in range(g)]
        assert max(sizes) - min(sizes) <= 1


def test_keys_do_not_depend_on_world_size():
    keys1 = {i: sharding.exposure_key(1963, i) for i in sharding.shard_indices(16, 1, 0)}
    keys4 = {}
    for r in range(4):
        keys4.update({i: sharding.exposure_key(1963, i) for i in sharding.shard_indices(16, 4, r)})
    assert keys1 == keys4 and len(set(keys1.values())) == 16
    assert sharding.filename_key(1963, '0001_raw.fits') != sharding.filename_key(1963, '0002_raw.fits')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)

    def make(i, key):       # stands in for one exposure: a deterministic function of the key only
        rng = np.random.Generator(np.random.Philox(key=(key[0] << 32) | key[1]))
        return {'rank': rank, 'checksum': int(rng.integers(0, 2 ** 31))}

    merged = sharding.run_sharded(13, make, world, rank, visit_seed=1963)
    np.save(os.path.join(out_dir, 'r%d.npy' % rank),
            np.array([[i, merged[i]['rank'], merged[i]['checksum']] for i in sorted(merged)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_matches_single_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / 'r0.npy')
    b = np.load(tmp_path / 'r1.npy')
    assert np.array_equal(a, b)                       # every rank holds the merged result
    assert list(a[:, 0]) == list(range(13))
    assert list(a[:, 1]) == [i % 2 for i in range(13)]   # strided ownership
    single = sharding.run_sharded(
        13, lambda i, key: int(np.random.Generator(np.random.Philox(key=(key[0] << 32) | key[1]))
                               .integers(0, 2 ** 31)), 1, 0, visit_seed=1963)
    assert [single[i] for i in range(13)] == list(a[:, 2])   # same frames whatever the GPU count
