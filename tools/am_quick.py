import sys, torch, ctypes
sys.path.insert(0, '.')
from monopsr_b200 import lib as mlib
L = mlib.load(); dev = torch.device('cuda:0')
for n in (1024, 1536, 2048, 2304):
    g = torch.Generator(device='cpu').manual_seed(200)
    x = torch.randn(32, n, 3, generator=g).to(dev); y = torch.randn(32, n, 3, generator=g).to(dev)
    mt = torch.empty(32, n, n, device=dev)
    f = lambda: L.mpb_approxmatch(32, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), None, mlib.stream_ptr())
    f(); f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    print(n, round(a.elapsed_time(b) / 5 * 1e3, 1), 'us', 'rowsum err', float((mt.sum(2) - 1).abs().max()))
