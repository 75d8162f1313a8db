"""Multi-GPU path of the DIP workload: independent images are sharded over ranks with no
data-path collective; only host-side scalar metrics are gathered.  Exercised here with world_size 2 on the
gloo backend (CPU), which is what the host logic needs."""
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


def _worker(rank, world, port, n_images, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'deep-super-resolution_b200'))
    from dsr_b200 import sharder
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = sharder.images_for_rank(n_images, rank, world)
    results = {i: {'psnr': 20.0 + i, 'iters': 10 * (i + 1)} for i in mine}
    merged = sharder.gather_results(results)
    t = sharder.max_over_ranks(float(rank + 1))
    if rank == 0:
        torch.save({'merged': merged, 'mine': mine, 't': t}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_image_sharding_world2(tmp_path):
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), 7, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r['mine'] == [0, 2, 4, 6]
    assert sorted(r['merged']) == list(range(7))
    assert r['merged'][5] == {'psnr': 25.0, 'iters': 60}
    assert r['t'] == 2.0


def test_partition_is_exact():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'deep-super-resolution_b200'))
    from dsr_b200 import sharder
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            parts = [sharder.images_for_rank(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_run_sharded_in_flight_matches_sequential():
    """in_flight > 1 runs the rank's images on worker threads (each with its own current stream on a GPU box); the
    gathered results are those of the sequential loop."""
    import sys
    import threading
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'deep-super-resolution_b200'))
    from dsr_b200 import sharder
    seen = set()

    def run_image(i):
        seen.add(threading.get_ident())
        return {'index': i, 'value': float(torch.full((4,), float(i)).sum())}

    seq = sharder.run_sharded(9, run_image, rank=1, world_size=2)
    par = sharder.run_sharded(9, run_image, rank=1, world_size=2, in_flight=3)
    assert seq == par and sorted(par) == [1, 3, 5, 7]
    assert len(seen) >= 2
