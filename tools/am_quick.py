"""approxmatch time at several sizes and point densities (unit-variance clouds and clouds ten times denser, where the
zero-term skipping of the fine levels rarely triggers); prints row-sum error as a sanity check"""
import sys, torch
sys.path.insert(0, '.')
from monopsr_b200 import lib as mlib
L = mlib.load(); dev = torch.device('cuda:0')
for scale in (1.0, 0.1):
    for n in (1024, 2304):
        g = torch.Generator(device='cpu').manual_seed(200)
        x = (torch.randn(32, n, 3, generator=g) * scale).to(dev); y = (torch.randn(32, n, 3, generator=g) * scale).to(dev)
        mt = torch.empty(32, n, n, device=dev)
        f = lambda: L.mpb_approxmatch(32, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), None, mlib.stream_ptr())
        f(); f(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): f()
        b.record(); torch.cuda.synchronize()
        print('scale', scale, 'n', n, round(a.elapsed_time(b) / 5 * 1e3, 1), 'us', 'rowsum err', float((mt.sum(2) - 1).abs().max()),
              'checksum %.9e' % float(mt.double().sum()))
