"""Kernel-alone rates (CUDA events) of one unfused VEGAS pass without the integrand: the materialised pipeline
(strat_sample -> map_forward_packed -> accumulate_fused) against the two-kernel one (sample_map, accumulate_regen)."""
import statistics
import sys

sys.path.insert(0, ".")
import torch

from torchquad_b200 import ops
from torchquad_b200.integration.vegas_map import VEGASMap

dev = torch.device("cuda")


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return statistics.median(ts)


def case(label, dim, dt, ns, ni, nh_mean):
    C = ns**dim
    dh = torch.full((C,), 1.0 / C, dtype=dt, device=dev)
    _nh, offsets = ops.strat_nh(dh, nh_mean * C)
    M = int(offsets[-1])
    vm = VEGASMap(ni, dim, "torch", dt, device=dev)
    vm.weights.copy_(torch.rand_like(vm.weights) + 0.1)
    vm.counts.fill_(1)
    vm.update_map()
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    f = torch.rand(M, dtype=dt, device=dev)
    elt = f.element_size()
    edges = vm.packed_edges()
    h = vm.hist_pairs()
    y = ops.strat_sample(offsets, ns, dim, dt, 0, M, seed=1, call_idx=7)
    x, jac, _ = ops.map_forward_packed(y, edges, dom)
    res = {}
    res["strat_sample"] = (timeit(lambda: ops.strat_sample(offsets, ns, dim, dt, 0, M, seed=1, call_idx=7)), dim * elt)
    res["map_forward_packed"] = (timeit(lambda: ops.map_forward_packed(y, edges, dom)), (2 * dim + 1) * elt)
    res["accumulate_fused"] = (timeit(lambda: ops.accumulate_fused(y, f, jac, 1.0, vm.weights, vm.counts)), (dim + 3) * elt)
    del x
    res["sample_map"] = (timeit(lambda: ops.sample_map(offsets, ns, dim, dt, 0, M, 1, 7, dom, edges_packed=edges)), (dim + 1) * elt)
    res["accumulate_regen"] = (timeit(lambda: ops.accumulate_regen(offsets, ns, dim, 0, M, ni, f, jac, 1.0, 1, 7, hist_pairs=h)), 3 * elt)
    old = sum(res[k][0] for k in ("strat_sample", "map_forward_packed", "accumulate_fused"))
    new = res["sample_map"][0] + res["accumulate_regen"][0]
    for k, (t, b) in res.items():
        print(f"{label:30s} {k:20s} M={M:.3e} {t*1e3:8.2f} ms  {M/t:.3e} rows/s  {M*b/t/1e9:7.0f} GB/s algorithmic", flush=True)
    print(f"{label:30s} materialised {old*1e3:.2f} ms -> two kernels {new*1e3:.2f} ms ({old/new:.2f}x)", flush=True)


case("8D f64 Ns=8 Ni=4096 nh=5", 8, torch.float64, 8, 4096, 5)
case("6D f32 Ns=13 Ni=4096 nh=9", 6, torch.float32, 13, 4096, 9)
case("16D f32 Ns=3 Ni=4096 nh=3", 16, torch.float32, 3, 4096, 3)
