"""N > 1 on real GPUs: two NCCL ranks shard trajectories / clips with sharding.shard_range, run the library on their shard and
gather the fixed-size records; every rank must hold exactly what one rank computes alone.  Skipped on a one-GPU box
(`gpurun --gpus 2` runs it); tests/test_multi_rank.py covers the same host logic with gloo on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _records_single(dev, n):
    from oracle import uplift as oup
    from upliftingtabletennis_b200 import ops, sharding, synthetic
    from upliftingtabletennis_b200.uplift import get_model
    m = get_model('connectstage', 'large', 'dynamic', 'new').to(dev).eval()
    m.load_state_dict(oup.random_state_dict(5))
    args = [torch.from_numpy(a).to(dev) for a in synthetic.trajectories(n, seed=21)]
    return m, args, ops, sharding


def _worker(rank, world, port, n, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    m, args, ops, sharding = _records_single(dev, n)
    a, b = sharding.shard_range(n, rank, world)
    rot, pos = m(*(x[a:b] for x in args))
    spin = ops.rotation_local(rot, pos)
    lens = args[2][a:b].sum(dim=1).long().tolist()
    recs = torch.stack([sharding.pack_record(spin[i], pos[i, :lens[i]]) for i in range(b - a)])
    table = sharding.gather_records(recs, n, world)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), table.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('n', [37, 64])
def test_two_ranks_nccl_gather_equals_single_rank(tmp_path, n):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    dev = torch.device('cuda:0')
    m, args, ops, sharding = _records_single(dev, n)
    rot, pos = m(*args)
    spin = ops.rotation_local(rot, pos)
    lens = args[2].sum(dim=1).long().tolist()
    expect = torch.stack([sharding.pack_record(spin[i], pos[i, :lens[i]]) for i in range(n)]).cpu().numpy()
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), 'rank%d.npy' % r))
        assert got.shape == (n, sharding.RECORD_FLOATS)
        # fp32 path: a trajectory's result does not depend on which other trajectories share its launch
        np.testing.assert_array_equal(got, expect)
