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


def _seed_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from wayne_b200 import run_visit
    os.environ.pop('WAYNE_B200_VISIT_SEED', None)
    s = run_visit.shared_visit_seed(None)           # no seed in the parameter file
    with open(os.path.join(out_dir, 'seed%d.txt' % rank), 'w') as f:
        f.write(str(s))
    assert run_visit.shared_visit_seed(1963) == 1963


def test_unseeded_visit_gets_one_seed_for_all_ranks(tmp_path, monkeypatch):
    """Without general.seed every rank used to draw its own (RANDSEED headers and Philox keys
    differed per rank): rank 0 draws, the others receive it."""
    port = _free_port()
    mp.spawn(_seed_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    seeds = [int(open(tmp_path / ('seed%d.txt' % r)).read()) for r in range(2)]
    assert seeds[0] == seeds[1] and 0 < seeds[0] < 2 ** 31
    from wayne_b200 import run_visit
    monkeypatch.setenv('WAYNE_B200_VISIT_SEED', '4242')     # what `wayne --gpus n` hands its children
    assert run_visit.shared_visit_seed(None) == 4242


def test_compat_streams_cannot_be_sharded():
    """rng='numpy' consumes one sequential numpy stream across a visit (run_visit.py:73-77):
    ranks seeded alike would repeat each other's noise, so the driver refuses."""
    import pytest
    from wayne_b200 import observation
    obs = observation.Observation()
    obs.rng = 'numpy'
    obs.exp_start_times = [0.0, 1.0]
    with pytest.raises(ValueError, match="cannot be sharded"):
        obs.run_observation(shard=(0, 2))


def test_native_ssv_draws_depend_on_the_exposure_key_only():
    """SSV generators draw from numpy's global stream; in native mode the exposure path keys it
    per exposure, so an exposure's durations do not depend on what the process drew before."""
    from wayne_b200 import exposure_generator as eg_mod
    src = open(eg_mod.__file__).read()
    assert "np.random.seed([key[0], key[1], 0x55F])" in src and "np.random.set_state(saved)" in src
    from wayne.trend_generators.scan_speed_varations import SSVModulatedSine
    from wayne import units as u
    rt = np.array([0.278, 7.624, 14.971, 22.317]) * u.s

    def draw(pre):
        np.random.seed(5)
        np.random.random(pre)                        # a different history of the global stream
        saved = np.random.get_state()
        np.random.seed([1963, 17, 0x55F])
        d, ri = SSVModulatedSine(10, 1.1, 50).get_subsample_exposure_times(None, None, rt, 0.05 * u.s)
        np.random.set_state(saved)
        return np.asarray(u.value_in(d, u.ms)), ri, np.random.random()

    a, b = draw(0), draw(1000)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] != b[2]
