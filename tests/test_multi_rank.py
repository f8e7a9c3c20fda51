"""N > 1 host logic on CPU: world_size-2 gloo processes shard clips, run a stand-in predictor and gather the
fixed-size records.  The real path uses the same code with NCCL on GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from upliftingtabletennis_b200 import sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 50000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_record_roundtrip():
    spin = torch.tensor([1.0, -2.0, 3.0])
    pos = np.arange(21, dtype=np.float32).reshape(7, 3)
    s, p = sharding.unpack_record(sharding.pack_record(spin, pos))
    assert torch.equal(s, spin) and np.array_equal(p.numpy(), pos)
    assert sharding.RECORD_FLOATS * 4 == 616        # bytes per clip, SURVEY.md section 8e


def _fake_predict(clip):
    rng = np.random.default_rng(clip)
    n = 10 + clip % 40
    return torch.tensor(rng.standard_normal(3), dtype=torch.float32), rng.standard_normal((n, 3)).astype(np.float32)


def _worker(rank, world, port, n_clips, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    table = sharding.run_clips(_fake_predict, list(range(n_clips)), torch.device('cpu'))
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), table.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('n_clips', [7, 8, 1])
def test_two_ranks_gloo(tmp_path, n_clips):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_clips, str(tmp_path)), nprocs=2, join=True)
    expect = torch.stack([sharding.pack_record(*_fake_predict(c)) for c in range(n_clips)]).numpy()
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), 'rank%d.npy' % r))
        assert np.array_equal(got, expect)
