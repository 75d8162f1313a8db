"""BASELINE configs[2]: DIP 4x SR of a batch of 64 independent synthetic 512x512 images sharded over the GPUs of one
node (image i -> rank i mod N, no collective, DIP.py:164-190), `--in-flight` images of a rank concurrently on its GPU.

    python tools/config3_run.py [--images 64] [--iters 3000] [--size 512] [--in-flight 3] [--out file.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config3_run.py ...

Every image gets a fresh same-seed network (torch.manual_seed(i) before get_net / get_noise), runs
dsr_b200.dip_sr_fused (one C call per iteration, device-side noise) and reports the PSNR of the resolved image against
its HR original; rank 0 prints the aggregate iterations / s (wall clock around the whole job, max over ranks, device
synchronised on both sides) and the mean / spread of the final PSNR over the batch (DIP.py:183-190 averages them)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
import torch.distributed as dist                # noqa: E402
import dsr_b200                                 # noqa: E402
from dsr_b200 import sharder                    # noqa: E402
from oracle import dip_oracle as O              # noqa: E402  (synthetic images only: the workload generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=64)
    ap.add_argument('--iters', type=int, default=3000)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--in-flight', type=int, default=3)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    cfg = {'learning_rate': 0.01, 'num_iter': args.iters, 'reg_noise_std': 0.05}
    size = args.size
    # inputs prepared up front (host work outside the timed region, as the reference's dataset loader is)
    mine = sharder.images_for_rank(args.images, rank, world)
    pairs = {i: O.synthetic_pair(i, size) for i in mine}
    nets = {}
    for i in mine:
        torch.manual_seed(i)
        nets[i] = (dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                                    upsample_mode='bilinear'), dsr_b200.get_noise(32, 'noise', (size, size)))

    def run_image(i):
        lr_img, hr = pairs[i]
        net, z = nets[i]
        out, losses = dsr_b200.dip_sr_fused(net, lr_img, (size, size), 4, cfg, dev, seed=1000 + i, net_input=z)
        torch.cuda.current_stream().synchronize()
        mse = float(((out - hr.unsqueeze(0).to(dev)) ** 2).mean())
        res = {'psnr': 10.0 * torch.log10(torch.tensor(1.0 / mse)).item(), 'loss_last': float(losses[-1])}
        net.cpu()                                   # DIP.py:109
        del nets[i]
        return res

    # warm-up image (plan construction, graph capture, allocator) outside the timed region
    torch.manual_seed(12345)
    wnet = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                            upsample_mode='bilinear')
    wl, _ = O.synthetic_pair(99, size)
    dsr_b200.dip_sr_fused(wnet, wl, (size, size), 4, dict(cfg, num_iter=5), dev, seed=1)
    del wnet
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    results = sharder.run_sharded(args.images, run_image, rank, world, in_flight=args.in_flight)
    torch.cuda.synchronize()
    dt = sharder.max_over_ranks(time.time() - t0, device=dev if world > 1 else None)
    if rank == 0:
        ps = torch.tensor([results[i]['psnr'] for i in sorted(results)])
        line = {'workload': f'BASELINE configs[2]: {args.images} independent {size}x{size} images, {args.iters} iterations each, '
                            f'sharded over {world} GPU(s), {args.in_flight} image(s) in flight per GPU, no collective',
                'n_gpus': world, 'images': args.images, 'iters_per_image': args.iters,
                'seconds': dt, 'value': args.images * args.iters / dt, 'unit': 'it/s (aggregate)',
                'psnr_mean_db': float(ps.mean()), 'psnr_std_db': float(ps.std()), 'psnr_min_db': float(ps.min()),
                'psnr_max_db': float(ps.max()), 'psnr_per_image': [round(float(v), 3) for v in ps]}
        s = json.dumps(line)
        print(s, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, 'w') as f:
                f.write(s + '\n')
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
