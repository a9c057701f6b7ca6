"""Multi-rank plumbing on CPU: world_size 2 over gloo.  The kernels cannot run here, so the per-rank local
statistics are produced by the oracle on each rank's shard; what is under test is the sharding arithmetic
and collective layer of torchquad_b200.distributed (row/cube partitions cover every sample exactly once,
packed all-reduce reproduces the single-process statistics, autograd-aware all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchquad_b200 as tq
        from torchquad_b200 import distributed as tqdist
        from torchquad_b200 import ops

        tqdist.enable()
        assert tqdist.rank_and_world() == (rank, world)
        dt = torch.float64
        # ---- Monte Carlo: disjoint row ranges of one Philox call == the single-process sample set
        N, dim = 10_001, 3
        b, e = tqdist.shard_range(N)
        dom = torch.tensor([[0.0, 2.0], [-1.0, 1.0], [0.5, 1.5]], dtype=dt)
        u = O.philox_uniform(5, 0, b, e - b, dim, dt)
        f = torch.sum(torch.sin(O.mc_sample_points(u, dom)), dim=1)
        total = f.sum().reshape(1)
        tqdist.all_reduce_sum_(total)
        full = torch.sum(torch.sin(O.mc_sample_points(O.philox_uniform(5, 0, 0, N, dim, dt), dom)), dim=1).sum()
        assert abs(float(total) - float(full)) < 1e-9
        # ---- VEGAS: cube-aligned shards from the device-side formula, stats all-reduced in one packed call
        ns, nc, vc = O.strat_config(2000, 3)
        g = torch.Generator().manual_seed(1)
        dh = torch.rand(nc, generator=g, dtype=dt) ** 4
        dh = dh / dh.sum()
        nh = O.strat_get_nh(dh, 6000)
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(nh, 0)])
        v = tq.VEGAS()
        M, c0, c1, r0, r1 = v._cube_aligned_shard(offsets)
        assert M == int(nh.sum()) and r0 == int(offsets[c0]) and r1 == int(offsets[c1])
        spans = [None] * world
        dist.all_gather_object(spans, (c0, c1, r0, r1))
        assert spans[0][0] == 0 and spans[-1][1] == nc and all(a[1] == b2[0] for a, b2 in zip(spans, spans[1:]))
        assert abs((r1 - r0) - M / world) <= int(nh.max())
        y_full = O.strat_get_y(nh, ns, 3, O.philox_uniform(9, 0, 0, M, 3, dt))
        xe, dxe, w_full, c_full = O.map_init(64, 3, dt)
        jf_full = torch.prod(torch.exp(2.0 * y_full), dim=1)
        O.map_accumulate(w_full, c_full, y_full, jf_full**2)
        JF_full, JF2_full = O.strat_accumulate(nh, jf_full)
        w, c = O.map_reset(64, 3, dt)
        O.map_accumulate(w, c, y_full[r0:r1], jf_full[r0:r1] ** 2)
        nh_local = torch.zeros_like(nh)
        nh_local[c0:c1] = nh[c0:c1]
        JF, JF2 = O.strat_accumulate(nh_local, jf_full[r0:r1])
        both = torch.stack([JF, JF2])
        tqdist.pack_all_reduce_sum_([w, both])
        tqdist.all_reduce_sum_(c)
        assert torch.equal(c, c_full)
        assert torch.allclose(w, w_full, rtol=1e-13) and torch.equal(both[0], JF_full) and torch.equal(both[1], JF2_full)
        # ---- fused VEGAS: block-cyclic cube ownership (tq_vegas_run_fused_sharded).  Every cube has exactly one owner, so
        # the per-cube sums need no collective; one packed fp64 all-reduce [hist pairs | I, sigma2, sum d^beta, sum nh]
        # reproduces the single-process statistics.
        lb, n_local = tqdist.cube_shard(nc)
        ids = tqdist.global_cube_ids(n_local, lb, rank, world)
        owned = [None] * world
        dist.all_gather_object(owned, ids.tolist())
        assert sorted(i for part in owned for i in part) == list(range(nc))
        cube_of_row = torch.repeat_interleave(torch.arange(nc), nh)
        mine = torch.isin(cube_of_row, ids)
        pairs = torch.zeros(3 * 64 * 2 + 8, dtype=dt)
        w_l, c_l = O.map_reset(64, 3, dt)
        O.map_accumulate(w_l, c_l, y_full[mine], jf_full[mine] ** 2)
        pairs[: 3 * 64 * 2] = torch.stack([w_l, c_l.to(dt)], dim=-1).reshape(-1)
        JF_l, JF2_l, nh_l = JF_full[ids], JF2_full[ids], nh[ids].to(dt)  # complete per-cube sums, owned cubes only
        ih = JF_l * vc / nh_l
        sig2 = torch.abs(JF2_l * vc * vc / nh_l - ih * ih)
        pairs[-8:-4] = torch.stack([ih.sum(), (sig2 / nh_l).sum(), torch.zeros((), dtype=dt), nh_l.sum()])
        tqdist.all_reduce_sum_(pairs)
        summed = pairs[: 3 * 64 * 2].view(3, 64, 2)
        assert torch.equal(summed[..., 1].to(torch.int64), c_full) and torch.allclose(summed[..., 0], w_full, rtol=1e-13)
        ih_f = JF_full * vc / nh.to(dt)
        assert abs(float(pairs[-8]) - float(ih_f.sum())) < 1e-12 * abs(float(ih_f.sum())) and float(pairs[-5]) == float(nh.sum())
        # ---- Newton-Cotes: point ranges of the flattened grid
        pts, hs, n = O.nc_grid("simpson", 9**3, torch.tensor([[0.0, 1.0]] * 3, dtype=dt))
        fvals = torch.prod(torch.cos(pts), dim=1)
        W = torch.einsum("i,j,k->ijk", *([tq.Simpson._rule_weights_1d(n, dt, "cpu")] * 3)).reshape(-1)
        pb, pe = tqdist.shard_range(n**3)
        part = (fvals[pb:pe] * W[pb:pe]).sum().reshape(1)
        tqdist.all_reduce_sum_(part)
        whole = O.nc_result("simpson", fvals, 3, n, hs)
        assert abs(float(part[0] * torch.prod(hs / 3.0)) - float(whole)) < 1e-14
        # ---- autograd-aware all-reduce: value is the global sum, gradient stays local
        p = torch.tensor([float(rank + 1)], dtype=dt, requires_grad=True)
        tot = ops.all_reduce_sum_autograd(p * 2.0)
        assert float(tot) == 2.0 * sum(range(1, world + 1))
        tot.backward()
        assert float(p.grad) == 2.0
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_and_collectives():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
