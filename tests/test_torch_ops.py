"""torch.library registration of the C-ABI kernels (torchquad_b200/torch_ops.py): schemas and fake implementations are
checked on the CPU (FakeTensorMode needs no GPU); values, autograd and torch.compile interop on the GPU."""
import pytest
import torch


def test_operators_are_registered_with_schemas_and_fake_impls():
    from torch._subclasses.fake_tensor import FakeTensorMode

    import torchquad_b200.torch_ops as T

    for name in T.OPERATORS:
        assert hasattr(torch.ops.tqb200, name), name
    assert str(torch.ops.tqb200.vegas_map_accumulate_.default._schema).count("!") == 2  # in-place on weights and counts
    with FakeTensorMode():
        dom = torch.empty((3, 2), dtype=torch.float64, device="cuda")
        pts = torch.ops.tqb200.mc_sample(dom, 1000, 1, 0, 0)
        assert pts.shape == (1000, 3) and pts.dtype == torch.float64 and pts.device.type == "cuda"
        assert torch.ops.tqb200.sum_columns(pts).shape == (3,)
        x, jac = torch.ops.tqb200.vegas_map_forward(pts, dom.new_empty((3, 11)), dom.new_empty((3, 10)))
        assert x.shape == (1000, 3) and jac.shape == (1000,)
        off = torch.empty(28, dtype=torch.int64, device="cuda")
        assert torch.ops.tqb200.vegas_strat_sample(off, 3, 3, torch.float32, 77, 1, 0).shape == (77, 3)
        edges = dom.new_empty((3, 10, 2))
        x2, jac2 = torch.ops.tqb200.vegas_sample_map(off, 3, 3, 77, edges, dom, 1, 0)
        assert x2.shape == (77, 3) and jac2.shape == (77,) and x2.dtype == torch.float64
        hist = dom.new_empty((3, 10, 2))
        assert torch.ops.tqb200.vegas_accumulate_regen_(off, 3, 3, 77, jac2, jac2, 1.0, hist, 1, 0).shape == (77,)
        JF, JF2 = torch.ops.tqb200.vegas_strat_accumulate(dom.new_empty(77), off)
        assert JF.shape == JF2.shape == (27,)
        assert torch.ops.tqb200.nc_grid_points(dom.new_empty((3, 5))).shape == (125, 3)
        assert torch.ops.tqb200.nc_contract(dom.new_empty((125, 2)), dom.new_empty((3, 5))).shape == (2,)
        assert torch.ops.tqb200.philox_uniform(10, 4, torch.float32, torch.device("cuda"), 1, 0, 0).dtype == torch.float32


@pytest.mark.gpu
def test_custom_ops_match_the_ctypes_path_and_compose_with_torch_compile(cuda):
    import torchquad_b200.torch_ops  # noqa: F401
    from torchquad_b200 import ops

    dom = torch.tensor([[0.0, 2.0], [-1.0, 1.0], [0.5, 1.5]], dtype=torch.float64, device=cuda)
    pts = torch.ops.tqb200.mc_sample(dom, 5000, 7, 0, 0)
    assert torch.equal(pts, ops.mc_sample(dom, 5000, 7, 0, 0))
    f = torch.sin(pts).sum(dim=1)
    assert torch.equal(torch.ops.tqb200.sum_columns(f), ops.reduce_sum(f))
    # autograd through mc_sample (domain) and sum_columns, as monte_carlo.py differentiates through its steps
    d = dom.clone().requires_grad_(True)
    vol = torch.prod(d[:, 1] - d[:, 0])
    (vol * torch.ops.tqb200.sum_columns(torch.sin(torch.ops.tqb200.mc_sample(d, 5000, 7, 0, 0)).sum(dim=1)) / 5000).backward()
    d2 = dom.clone().requires_grad_(True)
    vol2 = torch.prod(d2[:, 1] - d2[:, 0])
    (vol2 * ops.reduce_sum(torch.sin(ops.mc_sample(d2, 5000, 7, 0, 0)).sum(dim=1)) / 5000).backward()
    assert torch.allclose(d.grad, d2.grad, rtol=1e-12)
    # Newton-Cotes pair
    nodes = torch.linspace(0, 1, 9, dtype=torch.float64, device=cuda).repeat(3, 1).contiguous()
    w = torch.rand(3, 9, dtype=torch.float64, device=cuda)
    g = torch.ops.tqb200.nc_grid_points(nodes)
    assert torch.equal(g, ops.nc_grid_points(nodes))
    vals = torch.cos(g).prod(dim=1).requires_grad_(True)
    r = torch.ops.tqb200.nc_contract(vals, w)
    assert torch.equal(r.detach(), ops.nc_contract(vals.detach(), w))
    r.backward()
    assert torch.allclose(vals.grad, ops.nc_point_weights(w, 0, 9**3))
    torch.library.opcheck(torch.ops.tqb200.sum_columns, (f,))
    torch.library.opcheck(torch.ops.tqb200.mc_sample, (dom, 100, 1, 0, 0))

    # one traced program: sample -> integrand -> reduce (fake tensors + functionalisation must accept the operators)
    @torch.compile(backend="aot_eager", fullgraph=True)
    def mc(domain):
        p = torch.ops.tqb200.mc_sample(domain, 20000, 3, 0, 0)
        return torch.prod(domain[:, 1] - domain[:, 0]) * torch.ops.tqb200.sum_columns(torch.sin(p).sum(dim=1)) / 20000

    want = torch.prod(dom[:, 1] - dom[:, 0]) * ops.reduce_sum(torch.sin(ops.mc_sample(dom, 20000, 3, 0, 0)).sum(dim=1)) / 20000
    assert torch.allclose(mc(dom), want, rtol=1e-13)
