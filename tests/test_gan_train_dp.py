"""Data-parallel SRGAN training (the one path with a collective): the gradient exchange of
dsr_b200.gan_train -- broadcast of rank 0's state, mean all-reduce of the flat gradient buffers -- with world_size 2 on
the gloo backend (CPU).  The GPU box runs the same class over NCCL (bench.py --workload gan_train --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'deep-super-resolution_b200'))
    from dsr_b200.gan_train import GradExchange
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    x = GradExchange(device='cpu')
    g = torch.Generator().manual_seed(100 + rank)
    params = torch.rand(1000, generator=g)                 # every rank starts from different parameters ...
    x.broadcast(params)                                    # ... and takes rank 0's
    gD = torch.full((4096,), float(rank + 1))              # "discriminator gradient" of this replica
    gG = torch.arange(16, dtype=torch.float32) * (rank + 1)
    w = x.start(gD)                                        # asynchronous, as the fused step does around the generator phase
    other = gG.clone()                                     # (work that overlaps)
    x.finish(w, gD)
    x.allreduce_mean(gG)
    torch.save({'params': params, 'gD': gD, 'gG': gG, 'other': other, 'bytes': x.bytes, 'world': x.world},
               f'{out}.{rank}')
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_exchange_world2(tmp_path):
    out = str(tmp_path / 'r')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + '.0'), torch.load(out + '.1')
    want = torch.rand(1000, generator=torch.Generator().manual_seed(100))
    assert torch.equal(r0['params'], want) and torch.equal(r1['params'], want)
    for r in (r0, r1):
        assert r['world'] == 2
        assert torch.equal(r['gD'], torch.full((4096,), 1.5))                          # mean of 1 and 2
        assert torch.equal(r['gG'], torch.arange(16, dtype=torch.float32) * 1.5)
        assert r['bytes'] == 4096 * 4 + 16 * 4
