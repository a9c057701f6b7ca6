"""Kernel-alone rate of one fused stratified VEGAS pass (CUDA events) for the histogram modes of tq_fused_vegas,
with a cross-check that every mode produces the same counts (exactly) and weights (to rounding)."""
import os
import statistics
import sys

sys.path.insert(0, ".")
import torch

from torchquad_b200 import integrands as F
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


def case(label, fn, dim, dt, ns, ni, nh_mean, modes=("arrays", "pairs")):
    if os.environ.get("TQ_EXP_MODES"):
        modes = tuple(m for m in modes if m in os.environ["TQ_EXP_MODES"].split(","))
    C = ns**dim
    dh = torch.full((C,), 1.0 / C, dtype=dt, device=dev)
    nh, offsets = ops.strat_nh(dh, nh_mean * C)
    M = int(offsets[-1])
    vm = VEGASMap(ni, dim, "torch", dt, device=dev)
    # adapt the map a little so that bins are not uniform
    vm.weights.copy_(torch.rand_like(vm.weights) + 0.1)
    vm.counts.fill_(1)
    vm.update_map()
    s = fn.to_struct([0.0] * dim, [1.0] * dim, 1.0)
    JF = torch.zeros((2, C), dtype=dt, device=dev)
    res = {}
    for mode in modes:
        vm._reset_weight()
        JF.zero_()
        if mode == "arrays":
            run = lambda: ops.fused_vegas(s, vm.packed_edges(), vm.weights, vm.counts, 0, M, 1, 7, offsets=offsets, n_strat=ns,
                                          JF=JF[0], JF2=JF[1])
        elif mode == "pairs":
            h = vm.hist_pairs()
            h.zero_()
            run = lambda: ops.fused_vegas(s, vm.packed_edges(), None, None, 0, M, 1, 7, offsets=offsets, n_strat=ns, JF=JF[0],
                                          JF2=JF[1], hist_pairs=h)
        elif mode == "tile":  # whole pass (row_end < 0: count read on the device) -> band-privatised tile kernel when it applies
            h = vm.hist_pairs()
            h.zero_()
            run = lambda: ops.fused_vegas(s, vm.packed_edges(), None, None, 0, -M, 1, 7, offsets=offsets, n_strat=ns, JF=JF[0],
                                          JF2=JF[1], hist_pairs=h)
        elif mode == "records":
            rec = vm.records()
            run = lambda: ops.fused_vegas(s, None, None, None, 0, M, 1, 7, offsets=offsets, n_strat=ns, JF=JF[0], JF2=JF[1],
                                          records=rec, dtype=dt, n_intervals=ni)
        elif mode == "sweep":
            h = vm.hist_pairs()
            h.zero_()
            jf2 = torch.empty(M, dtype=dt, device=dev)
            g = int(os.environ.get('TQ_SWEEP_G', 0)) or vm.sweep_group(ns) or 1

            def run():
                ops.fused_vegas_deferred(s, vm.packed_edges(), 0, M, 1, 7, offsets, ns, JF[0], JF[1], jf2)
                ops.hist_sweep(offsets, ns, dim, jf2, ni, h, g, 1, 7)
        elif mode == "deferred":  # the pass alone (no histogram work at all)
            jf2 = torch.empty(M, dtype=dt, device=dev)
            run = lambda: ops.fused_vegas_deferred(s, vm.packed_edges(), 0, M, 1, 7, offsets, ns, JF[0], JF[1], jf2)
        elif mode == "nohist":
            run = lambda: ops.fused_vegas(s, vm.packed_edges(), None, None, 0, M, 1, 7, offsets=offsets, n_strat=ns, JF=JF[0],
                                          JF2=JF[1])
        run()
        torch.cuda.synchronize()
        if mode in ("pairs", "sweep", "tile"):
            vm.unpack_hist()
        if mode == "records":
            vm.unpack_records()
        res[mode] = (vm.weights.clone(), vm.counts.clone(), JF.clone())
        t = timeit(run)
        print(f"{label:34s} {mode:8s} M={M:.3e} {t*1e3:8.2f} ms  {M/t:.3e} samples/s", flush=True)
    base = res.get("arrays")
    for mode, (w, c, jf) in res.items():
        if base is None or mode in ("arrays", "nohist", "deferred"):
            continue
        same_counts = torch.equal(c, base[1])
        werr = float(((w - base[0]).abs() / base[0].abs().clamp_min(1e-300)).max())
        jerr = float((jf - base[2]).abs().max() / base[2].abs().max())
        print(f"    {mode} vs arrays: counts equal={same_counts} (sum {int(c.sum())} = {dim}*M {dim*M}), max rel weight diff {werr:.2e}, JF diff {jerr:.2e}", flush=True)
        assert same_counts


which = sys.argv[1:] or ["v8", "v16", "v8ref", "v4"]
if "v8" in which:
    case("8D f64 osc Ns=8 Ni=4096 nh=5", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float64, 8, 4096, 5, ("arrays", "pairs", "tile", "nohist"))
    case("8D f32 osc Ns=8 Ni=4096 nh=5", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float32, 8, 4096, 5, ("arrays", "pairs", "tile", "nohist"))
if "v16" in which:
    case("16D f32 prodpeak Ns=3 Ni=4096 nh=9", F.GenzProductPeak(16, a=2.0, u=0.5), 16, torch.float32, 3, 4096, 9, ("arrays", "pairs", "tile", "nohist"))
if "v8ref" in which:
    case("8D f64 osc Ns=8 Ni=1e7 nh=5", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float64, 8, 10_000_000, 5, ("arrays", "pairs", "records", "sweep", "deferred"))
    case("7D f64 osc Ns=5 Ni=3e6 nh=7", F.GenzOscillatory(7, a=0.5, u=0.3), 7, torch.float64, 5, 3_000_000, 7, ("arrays", "records", "sweep"))
    case("6D f32 gauss Ns=7 Ni=2e6 nh=4", F.GenzGaussian(6, a=3.0, u=0.5), 6, torch.float32, 7, 2_000_000, 4, ("arrays", "records", "sweep"))
if "v4" in which:
    case("4D f64 gauss Ns=10 Ni=4000 nh=5", F.GenzGaussian(4, a=5.0, u=0.5), 4, torch.float64, 10, 4000, 5, ("arrays", "pairs"))
    case("4D f64 gauss Ns=56 Ni=65536 nh=10", F.GenzGaussian(4, a=5.0, u=0.5), 4, torch.float64, 56, 65536, 10, ("arrays", "pairs"))
if "small_sweep" in which:
    case("4D f64 gauss Ns=10 Ni=4000 nh=5 (sweep)", F.GenzGaussian(4, a=5.0, u=0.5), 4, torch.float64, 10, 4000, 5, ("arrays", "sweep"))
    case("5D f32 gauss Ns=6 Ni=3000 nh=5 (sweep)", F.GenzGaussian(5, a=5.0, u=0.5), 5, torch.float32, 6, 3000, 5, ("arrays", "sweep"))
if "v8ref_big" in which:
    case("8D f64 osc Ns=8 Ni=1e7 nh=11.5", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float64, 8, 10_000_000, 11.5, ("sweep", "deferred"))
if "l2gran" in which:
    from torchquad_b200 import _lib
    for gran in (32, 64, 128):
        prev = _lib.l2_fetch_granularity(dev, gran)
        print(f"-- cudaLimitMaxL2FetchGranularity {gran} (was {prev})", flush=True)
        case("8D f64 osc Ns=8 Ni=1e7 nh=5", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float64, 8, 10_000_000, 5, ("sweep", "deferred"))
if "tile_small" in which:
    case("6D f64 gauss Ns=6 Ni=1000 nh=6 (tile)", F.GenzGaussian(6, a=3.0, u=0.5), 6, torch.float64, 6, 1000, 6, ("arrays", "pairs", "tile"))
    case("7D f32 osc Ns=5 Ni=777 nh=4 (tile)", F.GenzOscillatory(7, a=0.5, u=0.3), 7, torch.float32, 5, 777, 4, ("arrays", "pairs", "tile"))
