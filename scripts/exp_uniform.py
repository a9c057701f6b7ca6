import sys, torch, statistics
sys.path.insert(0, ".")
from torchquad_b200 import _lib, ops
dev = torch.device("cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e-3)
    return statistics.mean(ts)
for dt, dims in [(torch.float32, [1,3,4,8,10,16]), (torch.float64, [1,2,3,4,8,16])]:
    for dim in dims:
        rows = int(8e9 // (dim * (4 if dt==torch.float32 else 8)))
        dom = torch.tensor([[0.0,1.0]]*dim, dtype=dt, device=dev)
        buf = torch.empty((rows, dim), dtype=dt, device=dev)
        for aff in (True, False):
            f = (lambda: _lib.call("tq_mc_sample", buf.data_ptr(), dom.data_ptr(), 0, rows, dim, _lib.dtype_code(dt), 1, 0, _lib.stream_ptr(dev))) if aff else \
                (lambda: _lib.call("tq_philox_uniform", buf.data_ptr(), 0, rows, dim, _lib.dtype_code(dt), 1, 0, _lib.stream_ptr(dev)))
            s = t(f)
            print(f"{str(dt):14s} dim={dim:2d} affine={aff!s:5s} {buf.numel()*buf.element_size()/s/1e9:8.1f} GB/s  ({s*1e3:.2f} ms)")
        del buf
