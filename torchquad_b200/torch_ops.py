"""The C-ABI kernels of the unfused path as PyTorch custom operators (`torch.ops.tqb200.*`).

north_star asks for "a thin C-ABI torch custom-op extension": the integrator classes call the library through ctypes
(`ops.py`, no dispatcher overhead on the latency-bound small-problem paths); this module registers the same entry points
with `torch.library` -- schema, CUDA implementation, fake (meta) implementation and, where the reference differentiates
through the step, an autograd formula -- so that user code can put them inside `torch.compile`d / traced / fake-tensor
programs (e.g. a compiled integrand pipeline that samples, evaluates and reduces in one graph).

    import torchquad_b200.torch_ops            # registers the operators
    pts = torch.ops.tqb200.mc_sample(domain, 1_000_000, seed, 0, 0)
    total = torch.ops.tqb200.sum_columns(fn(pts))

Every operator launches on the current stream and never synchronises.  No CPU implementation is registered (there is
no CPU path in this package); the fake implementations make shape / dtype inference work on any device.
"""
import torch

from . import ops

_LIB = "tqb200"


def _define(name, mutates=()):
    return torch.library.custom_op(f"{_LIB}::{name}", mutates_args=mutates, device_types="cuda")


# ------------------------------------------------------------------------------------------- RNG / Monte Carlo
@_define("philox_uniform")
def philox_uniform(rows: int, dim: int, dtype: torch.dtype, device: torch.device, seed: int, call_idx: int, row_begin: int) -> torch.Tensor:
    """U[0,1) block [rows, dim] of Philox stream (seed, call_idx), global rows row_begin.. (rng.py:119-125)."""
    return ops.philox_uniform(rows, dim, dtype, device, seed, call_idx, row_begin)


@philox_uniform.register_fake
def _(rows, dim, dtype, device, seed, call_idx, row_begin):
    return torch.empty((rows, dim), dtype=dtype, device=device)


@_define("mc_sample")
def mc_sample(domain: torch.Tensor, rows: int, seed: int, call_idx: int, row_begin: int) -> torch.Tensor:
    """points = u * (b - a) + a for the [dim, 2] domain (monte_carlo.py:84-106)."""
    return ops._MCSample.apply(domain.detach(), rows, seed, call_idx, row_begin)


@mc_sample.register_fake
def _(domain, rows, seed, call_idx, row_begin):
    return domain.new_empty((rows, domain.shape[0]))


def _mc_sample_setup(ctx, inputs, output):
    _, ctx.rows, ctx.seed, ctx.call_idx, ctx.row_begin = inputs
    ctx.dtype = inputs[0].dtype


def _mc_sample_backward(ctx, grad):
    # d points / d domain: regenerated uniforms, reduced in fp64 (tq_mc_sample_backward)
    g = grad.contiguous()
    gd = torch.ops.tqb200.mc_sample_backward(g, ctx.rows, ctx.seed, ctx.call_idx, ctx.row_begin)
    return gd.to(ctx.dtype), None, None, None, None


@_define("mc_sample_backward")
def mc_sample_backward(grad: torch.Tensor, rows: int, seed: int, call_idx: int, row_begin: int) -> torch.Tensor:
    from ._lib import call, dtype_code, on_device, ptr, stream_ptr

    dim = grad.shape[1]
    gd = torch.zeros((dim, 2), dtype=torch.float64, device=grad.device)
    if rows > 0:
        with on_device(grad.device):
            wsp, wsn = ops._ws(grad.device)
            call("tq_mc_sample_backward", ptr(grad), row_begin, row_begin + rows, dim, dtype_code(grad.dtype),
                 seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, ptr(gd), wsp, wsn, stream_ptr(grad.device))
    return gd


@mc_sample_backward.register_fake
def _(grad, rows, seed, call_idx, row_begin):
    return grad.new_empty((grad.shape[1], 2), dtype=torch.float64)


mc_sample.register_autograd(_mc_sample_backward, setup_context=_mc_sample_setup)


@_define("sum_columns")
def sum_columns(f: torch.Tensor) -> torch.Tensor:
    """sum(f, axis 0) accumulated in fp64 and rounded once to f.dtype (monte_carlo.py:77)."""
    s, _ = ops.sum_columns(f)
    return s.to(f.dtype).reshape(f.shape[1:])


@sum_columns.register_fake
def _(f):
    return f.new_empty(f.shape[1:])


sum_columns.register_autograd(lambda ctx, g: g.unsqueeze(0).expand(ctx.shape),
                              setup_context=lambda ctx, inputs, output: setattr(ctx, "shape", inputs[0].shape))


# ------------------------------------------------------------------------------------------- VEGAS
@_define("vegas_map_forward")
def vegas_map_forward(y: torch.Tensor, x_edges: torch.Tensor, dx_edges: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """(x, jac) of VEGASMap.get_X / get_Jac in one pass (vegas_map.py:44-74)."""
    x, jac, _ = ops.map_forward(y, x_edges, dx_edges)
    return x, jac


@vegas_map_forward.register_fake
def _(y, x_edges, dx_edges):
    return torch.empty_like(y), y.new_empty((y.shape[0],))


@_define("vegas_map_accumulate_", mutates=("weights", "counts"))
def vegas_map_accumulate_(y: torch.Tensor, jf2: torch.Tensor, weights: torch.Tensor, counts: torch.Tensor) -> None:
    """weights[d, k] += jf2, counts[d, k] += 1 in place (vegas_map.py:99-111)."""
    ops.map_accumulate(y, jf2, weights, counts)


@vegas_map_accumulate_.register_fake
def _(y, jf2, weights, counts):
    return None


@_define("vegas_strat_sample")
def vegas_strat_sample(offsets: torch.Tensor, n_strat: int, dim: int, dtype: torch.dtype, rows: int, seed: int, call_idx: int) -> torch.Tensor:
    """Stratified points y [rows, dim], rows sorted by cube, cube-keyed Philox (vegas_stratification.py:140-165)."""
    return ops.strat_sample(offsets, n_strat, dim, dtype, 0, rows, seed=seed, call_idx=call_idx)


@vegas_strat_sample.register_fake
def _(offsets, n_strat, dim, dtype, rows, seed, call_idx):
    return offsets.new_empty((rows, dim), dtype=dtype)


@_define("vegas_strat_accumulate")
def vegas_strat_accumulate(jf: torch.Tensor, offsets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-cube sums (JF, JF2) of jf and jf^2 over each cube's rows (vegas_stratification.py:46-70)."""
    JF, JF2 = ops._StratAccumulate.apply(jf.detach(), offsets, 0, 0, offsets.shape[0] - 1)
    return JF, JF2


@vegas_strat_accumulate.register_fake
def _(jf, offsets):
    n = offsets.shape[0] - 1
    return jf.new_empty((n,)), jf.new_empty((n,))


@_define("vegas_sample_map")
def vegas_sample_map(offsets: torch.Tensor, n_strat: int, dim: int, rows: int, edges_packed: torch.Tensor, domain: torch.Tensor,
                     seed: int, call_idx: int) -> tuple[torch.Tensor, torch.Tensor]:
    """(x [rows, dim] in domain coordinates, jac [rows]) of a stratified pass straight from the Philox stream: get_Y + get_X +
    get_Jac + x*size+start in one kernel, y never materialised (vegas_stratification.py:140-165, vegas_map.py:44-74)."""
    return ops.sample_map(offsets, n_strat, dim, edges_packed.dtype, 0, rows, seed, call_idx, domain, edges_packed=edges_packed)


@vegas_sample_map.register_fake
def _(offsets, n_strat, dim, rows, edges_packed, domain, seed, call_idx):
    return edges_packed.new_empty((rows, dim)), edges_packed.new_empty((rows,))


@_define("vegas_accumulate_regen_", mutates=("hist_pairs",))
def vegas_accumulate_regen_(offsets: torch.Tensor, n_strat: int, dim: int, rows: int, f: torch.Tensor, jac: torch.Tensor,
                            volume: float, hist_pairs: torch.Tensor, seed: int, call_idx: int) -> torch.Tensor:
    """jf = (f*volume)*jac for the rows of `vegas_sample_map`, and hist_pairs[d, k] += {jf^2, 1} with the bin ids regenerated
    from the same Philox blocks (vegas.py:104-112,284-287, vegas_map.py:99-111).  Returns jf."""
    jf, _ = ops.accumulate_regen(offsets, n_strat, dim, 0, rows, hist_pairs.shape[1], f, jac, volume, seed, call_idx,
                                 hist_pairs=hist_pairs)
    return jf


@vegas_accumulate_regen_.register_fake
def _(offsets, n_strat, dim, rows, f, jac, volume, hist_pairs, seed, call_idx):
    return f.new_empty((rows,))


# ------------------------------------------------------------------------------------------- Newton-Cotes
@_define("nc_grid_points")
def nc_grid_points(nodes: torch.Tensor) -> torch.Tensor:
    """points[p, d] = nodes[d, i_d(p)], dimension 0 slowest (integration_grid.py:98-99)."""
    return ops._GridPoints.apply(nodes.detach(), 0, nodes.shape[1] ** nodes.shape[0])


@nc_grid_points.register_fake
def _(nodes):
    dim, n = nodes.shape
    return nodes.new_empty((n**dim, dim))


@_define("nc_contract")
def nc_contract(f: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """sum_p f[p, ...] * prod_d w[d, i_d(p)] in fp64, rounded once (grid_integrator.py:57-91)."""
    dim, n = w.shape
    return ops._Contract.apply(f.detach(), w, 0, n**dim, False)


@nc_contract.register_fake
def _(f, w):
    return f.new_empty(f.shape[1:])


def _nc_contract_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])
    ctx.fshape = inputs[0].shape


def _nc_contract_backward(ctx, g):
    (w,) = ctx.saved_tensors
    dim, n = w.shape
    W = ops.nc_point_weights(w.contiguous(), 0, n**dim)
    gf = W.reshape([-1] + [1] * (len(ctx.fshape) - 1)) * g.unsqueeze(0)
    return gf.expand(ctx.fshape), None


nc_contract.register_autograd(_nc_contract_backward, setup_context=_nc_contract_setup)

OPERATORS = ["philox_uniform", "mc_sample", "mc_sample_backward", "sum_columns", "vegas_map_forward", "vegas_map_accumulate_",
             "vegas_strat_sample", "vegas_strat_accumulate", "vegas_sample_map", "vegas_accumulate_regen_", "nc_grid_points",
             "nc_contract"]
