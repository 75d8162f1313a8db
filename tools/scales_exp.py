"""Where does the iteration time go by LEVEL?  Times dsr_dip_run (graph replay) at one size for num_scales = 1..5:
the difference between consecutive lines is the in-graph cost of one more (4x smaller) level.
    python tools/scales_exp.py [size] [iters]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
import torch, dsr_b200
from dsr_b200._lib import lib, check, StepBuffers
from dsr_b200 import _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
scales_list = [int(s) for s in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 2, 3, 4, 5]
dev = torch.device('cuda', 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
for scales in scales_list:
    torch.manual_seed(0)
    ds = dsr_b200.Downsampler(3, 4, 'lanczos2', phase=0.5, preserve_size=True)
    hr = torch.rand(1, 3, size, size).to(dev)
    lr_img = ds(hr)[0].contiguous()
    net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=scales,
                           upsample_mode='bilinear').to(dev)
    z_saved = dsr_b200.get_noise(32, 'noise', (size, size)).to(dev).contiguous()
    z = z_saved.clone()
    net(z)
    net.zero_grad()
    plan = net._plans[(size, size)]
    tables = ds._tables_for(size, size, dev)
    oh, ow = ds.out_size(size, size)
    flat, gflat = net.flat_buffers()
    f32 = dict(dtype=torch.float32, device=dev)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    out_hr = torch.empty((1, 3, size, size), **f32)
    out_lr, g_lr, g_hr = torch.empty((3, oh, ow), **f32), torch.empty((3, oh, ow), **f32), torch.empty((1, 3, size, size), **f32)
    losses = torch.zeros(iters + 64, **f32)
    b = StepBuffers(flat.data_ptr(), gflat.data_ptr(), m.data_ptr(), v.data_ptr(), net._bnflat.data_ptr(),
                    z_saved.data_ptr(), z.data_ptr(), lr_img.data_ptr(), out_hr.data_ptr(), out_lr.data_ptr(),
                    g_lr.data_ptr(), g_hr.data_ptr(), losses.data_ptr())
    stream = _lib.stream_ptr()
    check(lib.dsr_dip_run(plan.handle, tables.handle, C.byref(b), 0.01, 0.05, 7, 1, 5, stream), 'run')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.dsr_dip_run(plan.handle, tables.handle, C.byref(b), 0.01, 0.05, 7, 6, iters, stream), 'run')
    e1.record()
    torch.cuda.synchronize()
    print(f'size {size} num_scales {scales}: {e0.elapsed_time(e1) / iters * 1e3:8.1f} us / iteration, '
          f'{lib.dsr_plan_last_launches(plan.handle)} launches, loss {float(losses[iters + 4]):.5f}', flush=True)
    if os.environ.get('DSR_TIMELINE') == '2':          # in-kernel stamps (make kstamp; DSR_B200_LIB=...kstamp.so)
        buf = C.create_string_buffer(1 << 20)
        lib.dsr_timeline_dump(buf, 1 << 20)
        import re
        print('  #  lead  total   body  grid  name     (lead: entry -> wait done; total: wait done -> next wait done; body: block 0)')
        for l in buf.value.decode().splitlines():
            if l.startswith('#'):
                print(l); continue
            i, lead, tot, body, gb, g, name = l.split('\t')[:7]
            marks = name[name.index(' m'):] if ' m' in name else ''
            name = name.split(' ')[0]
            print(f'{int(i):4d} {float(lead):6.2f} {float(tot):6.2f} {float(body):6.2f} {gb:>9s}  ' +
                  re.sub(r'\(.*', '', name).replace('void ', '').replace('dsr::', '')[:60] + marks)
    if os.environ.get('DSR_TIMELINE') == '1':
        buf = C.create_string_buffer(1 << 20)
        nbytes = lib.dsr_timeline_dump(buf, 1 << 20)
        import collections, re
        rows = [l.split('\t') for l in buf.value.decode().splitlines()]
        agg, cnt = collections.OrderedDict(), collections.Counter()
        for _i, us, _g, name in rows:
            k = re.sub(r'\(.*', '', name).replace('void ', '').replace('dsr::', '')
            if float(us) < 0:
                k = '[side stream] ' + k
            agg[k] = agg.get(k, 0.0) + abs(float(us)); cnt[k] += 1
        print(f'--- in-graph timeline, {len(rows)} stamped launches, sum {sum(agg.values()):.1f} us')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            print(f'{v:9.1f} us  n={cnt[k]:3d}  mean {v / cnt[k]:6.2f}  {k}')
        if os.environ.get('DSR_TIMELINE_ORDER'):
            for i, us, g, name in rows:
                print(f'{int(i):4d} {float(us):7.2f} {g:>6s}  ' + re.sub(r'\(.*', '', name).replace('void ', '').replace('dsr::', ''))
    del net, plan, tables
