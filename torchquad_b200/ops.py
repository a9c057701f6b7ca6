"""Tensor-level wrappers of the C ABI (include/tqb200.h) plus the autograd glue of the unfused path.

Every function takes CUDA tensors, launches on the current stream and returns without synchronising.
Autograd: the reference gets gradients for free because it is written in differentiable ATen ops; here
the three places where a gradient crosses one of our kernels get an explicit backward
(Monte Carlo samples wrt the domain, per-cube sums wrt jf, grid points wrt the 1-D nodes, weighted
contraction / column sums wrt the integrand values) -- pinned by tests mirroring
/root/reference/tests/gradient_test.py:162-259.
"""
import torch

from . import _lib
from ._lib import call, dtype_code, on_device, ptr, require_cuda, stream_ptr, workspace


def _ws(device):
    ws = workspace(device)
    return ws.data_ptr(), ws.numel()


# ------------------------------------------------------------------------------------------- RNG / MC
def philox_uniform(rows, dim, dtype, device, seed, call_idx, row_begin=0):
    """U[0,1) block [rows, dim] of stream (seed, call_idx), global rows row_begin.. (rng.py:119-125)."""
    out = torch.empty((rows, dim), dtype=dtype, device=device)
    require_cuda(out)
    if rows > 0:
        with on_device(out.device):
            call("tq_philox_uniform", ptr(out), row_begin, row_begin + rows, dim, dtype_code(dtype),
                 seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, stream_ptr(out.device))
    return out


class _MCSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, domain, rows, seed, call_idx, row_begin):
        require_cuda(domain)
        dom = domain.detach().contiguous()
        dim = dom.shape[0]
        out = torch.empty((rows, dim), dtype=dom.dtype, device=dom.device)
        if rows > 0:
            with on_device(dom.device):
                call("tq_mc_sample", ptr(out), ptr(dom), row_begin, row_begin + rows, dim, dtype_code(dom.dtype),
                     seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, stream_ptr(dom.device))
        ctx.meta = (rows, seed, call_idx, row_begin, dom.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        rows, seed, call_idx, row_begin, dtype = ctx.meta
        g = grad_out.contiguous()
        dim = g.shape[1]
        gd = torch.zeros((dim, 2), dtype=torch.float64, device=g.device)
        if rows > 0:
            with on_device(g.device):
                wsp, wsn = _ws(g.device)
                call("tq_mc_sample_backward", ptr(g), row_begin, row_begin + rows, dim, dtype_code(dtype),
                     seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, ptr(gd), wsp, wsn, stream_ptr(g.device))
        return gd.to(dtype), None, None, None, None


def mc_sample(domain, rows, seed, call_idx, row_begin=0, call_offset=None):
    """points = u*(b-a)+a for the [dim,2] domain (monte_carlo.py:84-106); differentiable wrt domain.

    `call_offset` (uint32/int32 device tensor, 1 element) is added to `call_idx` ON THE DEVICE: launches replayed
    from a CUDA graph draw fresh samples after the word is incremented (no autograd on that path)."""
    if call_offset is not None and torch.is_grad_enabled() and domain.requires_grad:
        # the differentiable path regenerates its uniforms in backward from a HOST call index: read the device word
        # (one sync; this is the eager, gradient-carrying path of a compiled integrate)
        call_idx, call_offset = call_idx + int(call_offset.item()), None
    if call_offset is None:
        return _MCSample.apply(domain, rows, seed, call_idx, row_begin)
    require_cuda(domain, call_offset)
    dom = domain.detach().contiguous()
    dim = dom.shape[0]
    out = torch.empty((rows, dim), dtype=dom.dtype, device=dom.device)
    if rows > 0:
        with on_device(dom.device):
            call("tq_mc_sample_replayable", ptr(out), ptr(dom), row_begin, row_begin + rows, dim, dtype_code(dom.dtype),
                 seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, ptr(call_offset), stream_ptr(dom.device))
    return out


def sum_columns(f, want_sumsq=False):
    """fp64 column sums (and sums of squares) of f[rows] or f[rows, cols]; no autograd."""
    require_cuda(f)
    f2 = f.detach().contiguous()
    rows = f2.shape[0]
    cols = 1 if f2.dim() == 1 else int(f2[0].numel()) if rows > 0 else int(torch.Size(f2.shape[1:]).numel())
    make = torch.empty if rows > 0 else torch.zeros  # the kernels write every output element
    s = make(cols, dtype=torch.float64, device=f2.device)
    q = make(cols, dtype=torch.float64, device=f2.device) if want_sumsq else None
    if rows > 0:
        with on_device(f2.device):
            wsp, wsn = _ws(f2.device)
            call("tq_sum_columns", ptr(f2), rows, cols, dtype_code(f2.dtype), ptr(s), ptr(q), wsp, wsn,
                 stream_ptr(f2.device))
    return s, q


class _ReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f):
        s, _ = sum_columns(f)
        ctx.shape = f.shape
        return s.to(f.dtype).reshape(f.shape[1:])

    @staticmethod
    def backward(ctx, g):
        return g.unsqueeze(0).expand(ctx.shape)


def _complex_through_real(op, f, *args):
    """Complex integrand values (the reference's tests integrate a few, tests/integration_test_functions.py) go through
    the real kernels as [..., 2] real views: sums and weighted contractions are linear, so re / im parts are columns."""
    out = op(torch.view_as_real(f.contiguous() if not f.is_contiguous() else f), *args)
    return torch.view_as_complex(out.contiguous())


def reduce_sum(f):
    """sum(f, axis=0) accumulated in fp64 and rounded once to f.dtype (monte_carlo.py:77); differentiable."""
    if f.is_complex():
        return _complex_through_real(_ReduceSum.apply, f)
    return _ReduceSum.apply(f)


class _ReduceSumF64(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f):
        s, _ = sum_columns(f)
        ctx.shape, ctx.dtype = f.shape, f.dtype
        return s.reshape(f.shape[1:])

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).unsqueeze(0).expand(ctx.shape)


def reduce_sum_f64(f):
    """Like reduce_sum but returns the unrounded fp64 sums (for accumulation over chunks and ranks)."""
    if f.is_complex():
        return _complex_through_real(_ReduceSumF64.apply, f)
    return _ReduceSumF64.apply(f)


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        from . import distributed as tqdist

        out = t.detach().clone()
        tqdist.all_reduce_sum_(out)
        return out

    @staticmethod
    def backward(ctx, g):
        # every rank holds the same replicated result; each rank backpropagates its own contribution
        return g


def all_reduce_sum_autograd(t):
    """Sum over ranks that keeps the local autograd graph (d(total)/d(local part) = 1)."""
    return _AllReduceSum.apply(t)


# ------------------------------------------------------------------------------------------- VEGAS map
def map_forward(y, x_edges, dx_edges, want_x=True, want_jac=True, want_ids=False, want_offset=False):
    """(x, jac, ids[, offset]) of vegas_map.py:44-97 in one pass."""
    require_cuda(y, x_edges, dx_edges)
    y = y.contiguous()
    rows, dim = y.shape
    ni = dx_edges.shape[1]
    x = torch.empty_like(y) if want_x else None
    jac = torch.empty(rows, dtype=y.dtype, device=y.device) if want_jac else None
    ids = torch.empty((rows, dim), dtype=torch.int32, device=y.device) if want_ids else None
    off = torch.empty_like(y) if want_offset else None
    if rows > 0:
        with on_device(y.device):
            call("tq_vegas_map_forward", ptr(y), ptr(x_edges), ptr(dx_edges), ptr(x), ptr(jac), ptr(ids), ptr(off), rows,
                 dim, ni, dtype_code(y.dtype), stream_ptr(y.device))
    if want_offset:
        return x, jac, ids, off
    return x, jac, ids


def map_forward_packed(y, edges_packed, domain=None, want_ids=False, records=None, n_intervals=None):
    """(x, jac, ids) from packed edges; with `domain` ([dim, 2]) x is already x*size + start (vegas.py:109-110).
    `records` (large maps, see pack_records) replaces `edges_packed`; `n_intervals` is then required."""
    table = records if records is not None else edges_packed
    require_cuda(y, table, domain)
    y = y.contiguous()
    rows, dim = y.shape
    x = torch.empty_like(y)
    jac = torch.empty(rows, dtype=y.dtype, device=y.device)
    ids = torch.empty((rows, dim), dtype=torch.int32, device=y.device) if want_ids else None
    if rows > 0:
        ni = n_intervals if records is not None else edges_packed.shape[1]
        with on_device(y.device):
            call("tq_vegas_map_forward_packed", ptr(y), ptr(table),
                 _lib.TQ_EDGES_RECORDS if records is not None else _lib.TQ_EDGES_PAIRS, ptr(domain), ptr(x), ptr(jac), ptr(ids),
                 rows, dim, ni, dtype_code(y.dtype), stream_ptr(y.device))
    return x, jac, ids


def accumulate_fused(y, f, jac, volume, weights, counts, want_jf=True, records=None, n_intervals=None):
    """jf = (f*volume)*jac, weights[d,k] += jf^2, counts[d,k] += 1 in one pass; returns jf (or None).
    With `records` the histogram goes into the record table instead (weights / counts untouched)."""
    require_cuda(y, f, jac, weights, counts, records)
    y = y.contiguous()
    f = f.detach().contiguous()
    rows, dim = y.shape
    if f.shape != (rows,) or f.dtype != y.dtype:
        raise ValueError(f"integrand values must have shape ({rows},) and dtype {y.dtype}, got {tuple(f.shape)} / {f.dtype}")
    jf = torch.empty(rows, dtype=y.dtype, device=y.device) if want_jf else None
    if rows > 0:
        ni = n_intervals if records is not None else weights.shape[1]
        with on_device(y.device):
            call("tq_vegas_accumulate_fused", ptr(y), ptr(f), ptr(jac), float(volume), ptr(jf),
                 None if records is not None else ptr(weights), None if records is not None else ptr(counts), ptr(records),
                 rows, dim, ni, dtype_code(y.dtype), stream_ptr(y.device))
    return jf


def sample_map(offsets, n_strat, dim, dtype, row_begin, row_end, seed, call_idx, domain, edges_packed=None, records=None,
               n_intervals=None):
    """Rows [row_begin, row_end) of a VEGAS pass straight to (x [rows, dim] in domain coordinates, jac [rows]): get_Y + get_X +
    get_Jac + the x*size + start transform in one kernel, y never materialised (tq_vegas_sample_map).  `offsets` None: a
    warm-up pass (y = u * 0.999999 from the row-keyed stream)."""
    table = records if records is not None else edges_packed
    require_cuda(offsets, table, domain)
    if domain.dtype != dtype or tuple(domain.shape) != (dim, 2):
        raise ValueError(f"domain must be a [{dim}, 2] tensor of dtype {dtype}, got {tuple(domain.shape)} / {domain.dtype}")
    if edges_packed is not None and (edges_packed.dtype != dtype or edges_packed.shape[0] != dim):
        raise ValueError(f"edges_packed must be [{dim}, Ni, 2] of dtype {dtype}, got {tuple(edges_packed.shape)} / {edges_packed.dtype}")
    rows = row_end - row_begin
    x = torch.empty((rows, dim), dtype=dtype, device=table.device)
    jac = torch.empty(rows, dtype=dtype, device=table.device)
    if rows > 0:
        ni = n_intervals if records is not None else edges_packed.shape[1]
        dom = domain.detach().contiguous()
        with on_device(table.device):
            call("tq_vegas_sample_map", ptr(offsets), 0 if offsets is None else offsets.shape[0] - 1, n_strat, dim, dtype_code(dtype),
                 row_begin, row_end, ptr(table), _lib.TQ_EDGES_RECORDS if records is not None else _lib.TQ_EDGES_PAIRS, ni, ptr(dom),
                 seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, ptr(x), ptr(jac), stream_ptr(table.device))
    return x, jac


def accumulate_regen(offsets, n_strat, dim, row_begin, row_end, n_intervals, f, jac, volume, seed, call_idx, hist_pairs=None,
                     records=None, weights=None, counts=None, want_jf=True, want_jf2=False):
    """jf = (f*volume)*jac for the rows of `sample_map` and, with a target, the map histogram of those rows with the bin
    ids regenerated from the Philox stream (tq_vegas_accumulate_regen).  Returns (jf or None, jf^2 rows or None)."""
    require_cuda(offsets, f, jac, hist_pairs, records, weights, counts)
    rows = row_end - row_begin
    f = f.detach().contiguous()
    if f.shape != (rows,) or f.dtype != jac.dtype:
        raise ValueError(f"integrand values must have shape ({rows},) and dtype {jac.dtype}, got {tuple(f.shape)} / {f.dtype}")
    jf = torch.empty(rows, dtype=f.dtype, device=f.device) if want_jf else None
    jf2 = torch.empty(rows, dtype=f.dtype, device=f.device) if want_jf2 else None
    if rows > 0 and (want_jf or want_jf2 or hist_pairs is not None or records is not None or weights is not None):
        with on_device(f.device):
            call("tq_vegas_accumulate_regen", ptr(offsets), 0 if offsets is None else offsets.shape[0] - 1, n_strat, dim,
                 dtype_code(f.dtype), row_begin, row_end, n_intervals, ptr(f), ptr(jac), float(volume), ptr(jf), ptr(jf2),
                 ptr(hist_pairs), ptr(records), ptr(weights), ptr(counts), seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF,
                 stream_ptr(f.device))
    return jf, jf2


def map_accumulate(y, jf2, weights, counts):
    """weights[d,k] += jf2, counts[d,k] += 1 in place (vegas_map.py:99-111)."""
    require_cuda(y, jf2, weights, counts)
    y = y.contiguous()
    jf2 = jf2.detach().contiguous()
    rows, dim = y.shape
    if rows > 0:
        with on_device(y.device):
            call("tq_vegas_map_accumulate", ptr(y), ptr(jf2), ptr(weights), ptr(counts), rows, dim, weights.shape[1],
                 dtype_code(y.dtype), stream_ptr(y.device))


def _map_scratch(dim, ni, dtype, device):
    nbytes = _lib.load().tq_vegas_map_workspace_bytes(dim, ni, dtype_code(dtype))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def map_smooth(weights, counts, alpha):
    """_smooth_map (vegas_map.py:113-172).  Returns (smoothed, status[4] int32 device tensor)."""
    require_cuda(weights, counts)
    dim, ni = weights.shape
    out = torch.empty_like(weights)
    status = torch.zeros(4, dtype=torch.int32, device=weights.device)
    scratch = _map_scratch(dim, ni, weights.dtype, weights.device)
    with on_device(weights.device):
        call("tq_vegas_map_smooth", ptr(weights.contiguous()), ptr(counts.contiguous()), ptr(out), dim, ni, float(alpha),
             dtype_code(weights.dtype), ptr(status), ptr(scratch), scratch.numel(), stream_ptr(weights.device))
    return out, status


def map_scratch(dim, ni, dtype, device):
    """Scratch buffer for map_smooth / map_update of a [dim, ni] map (reusable across calls on one stream)."""
    return _map_scratch(dim, ni, dtype, device)


def map_update(x_edges, dx_edges, weights, counts, alpha, status, edges_packed=None, scratch=None):
    """update_map (vegas_map.py:185-261) in place; `status` is an int32[4] device tensor (see header);
    `edges_packed` ([dim, Ni, 2], optional) receives the new edges in the packed gather layout."""
    require_cuda(x_edges, dx_edges, weights, counts, status, edges_packed)
    dim, ni = weights.shape
    if scratch is None:
        scratch = _map_scratch(dim, ni, weights.dtype, weights.device)
    with on_device(weights.device):
        call("tq_vegas_map_update", ptr(x_edges), ptr(dx_edges), ptr(weights), ptr(counts), ptr(edges_packed), dim, ni,
             float(alpha), dtype_code(weights.dtype), ptr(status), ptr(scratch), scratch.numel(),
             stream_ptr(weights.device))


# ------------------------------------------------------------------------------------------- stratification
def strat_nh(dh, nevals_exp):
    """nh (int64 [C]) and its exclusive scan offsets (int64 [C+1]) (vegas_stratification.py:92-103)."""
    require_cuda(dh)
    n = dh.shape[0]
    nh = torch.empty(n, dtype=torch.int64, device=dh.device)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dh.device)
    with on_device(dh.device):
        wsp, wsn = _ws(dh.device)
        call("tq_vegas_strat_nh", ptr(dh.contiguous()), n, float(nevals_exp), dtype_code(dh.dtype), ptr(nh), ptr(offsets),
             wsp, wsn, stream_ptr(dh.device))
    return nh, offsets


def strat_offsets(nh):
    """Exclusive scan of a user-provided nh (int64 [C]) -> offsets int64 [C+1]."""
    require_cuda(nh)
    nh = nh.contiguous()
    n = nh.shape[0]
    offsets = torch.empty(n + 1, dtype=torch.int64, device=nh.device)
    with on_device(nh.device):
        wsp, wsn = _ws(nh.device)
        call("tq_vegas_strat_offsets", ptr(nh), n, ptr(offsets), wsp, wsn, stream_ptr(nh.device))
    return offsets


def strat_sample(offsets, n_strat, dim, dtype, row_begin, row_end, u_in=None, seed=0, call_idx=0):
    """y rows [row_begin,row_end) of vegas_stratification.py:140-165 (injected `u_in` or cube-keyed Philox)."""
    require_cuda(offsets, u_in)
    rows = row_end - row_begin
    y = torch.empty((rows, dim), dtype=dtype, device=offsets.device)
    if u_in is not None:
        u_in = u_in.contiguous()
        if tuple(u_in.shape) != (rows, dim) or u_in.dtype != dtype:
            raise ValueError(f"rng.uniform returned shape {tuple(u_in.shape)} / {u_in.dtype}, expected {(rows, dim)} / {dtype}")
    if rows > 0:
        with on_device(offsets.device):
            call("tq_vegas_strat_sample", ptr(offsets), offsets.shape[0] - 1, n_strat, dim, dtype_code(dtype), ptr(u_in),
                 seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, row_begin, row_end, ptr(y), stream_ptr(offsets.device))
    return y


class _StratAccumulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, jf, offsets, row_base, cube_begin, cube_end):
        require_cuda(jf, offsets)
        v = jf.detach().contiguous()
        n_cubes = offsets.shape[0] - 1
        JF = torch.zeros(n_cubes, dtype=v.dtype, device=v.device)
        JF2 = torch.zeros(n_cubes, dtype=v.dtype, device=v.device)
        with on_device(v.device):
            call("tq_vegas_strat_accumulate", ptr(v), row_base, ptr(offsets), cube_begin, cube_end, ptr(JF), ptr(JF2),
                 dtype_code(v.dtype), stream_ptr(v.device))
        ctx.save_for_backward(offsets)
        ctx.meta = (row_base, v.shape[0])
        ctx.mark_non_differentiable(JF2)
        return JF, JF2

    @staticmethod
    def backward(ctx, gJF, _gJF2):
        (offsets,) = ctx.saved_tensors
        row_base, rows = ctx.meta
        g = torch.empty(rows, dtype=gJF.dtype, device=gJF.device)
        if rows > 0:
            with on_device(g.device):
                call("tq_vegas_strat_accumulate_backward", ptr(gJF.contiguous()), ptr(offsets), offsets.shape[0] - 1,
                     row_base, row_base + rows, ptr(g), dtype_code(g.dtype), stream_ptr(g.device))
        return g, None, None, None, None


def strat_accumulate(jf, offsets, row_base=0, cube_begin=0, cube_end=None):
    """JF[c], JF2[c] over each cube's rows in order (vegas_stratification.py:46-70); JF differentiable wrt jf.

    JF/JF2 are full-length [C] tensors, zero outside [cube_begin, cube_end); the sum of squares is not
    differentiated (the reference detaches everything derived from it, vegas.py:298-299)."""
    cube_end = offsets.shape[0] - 1 if cube_end is None else cube_end
    return _StratAccumulate.apply(jf, offsets, row_base, cube_begin, cube_end)


def strat_update(JF, JF2, nh, v_cubes, beta):
    """Estimator + update_DH: returns (dh [C], scalars fp64 [4] = I, sigma2, sum d^beta, sum nh)."""
    require_cuda(JF, JF2, nh)
    n = JF.shape[0]
    dh = torch.empty(n, dtype=JF.dtype, device=JF.device)
    scalars = torch.empty(4, dtype=torch.float64, device=JF.device)  # fully written by the kernel
    with on_device(JF.device):
        wsp, wsn = _ws(JF.device)
        call("tq_vegas_strat_update", ptr(JF.detach().contiguous()), ptr(JF2.detach().contiguous()), ptr(nh), n,
             float(v_cubes), float(beta), dtype_code(JF.dtype), ptr(dh), ptr(scalars), wsp, wsn, stream_ptr(JF.device))
    return dh, scalars


# ------------------------------------------------------------------------------------------- Newton-Cotes
class _GridPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, nodes, p_begin, p_end):
        require_cuda(nodes)
        nd = nodes.detach().contiguous()
        dim, n = nd.shape
        pts = torch.empty((p_end - p_begin, dim), dtype=nd.dtype, device=nd.device)
        if p_end > p_begin:
            with on_device(nd.device):
                call("tq_nc_grid_points", ptr(nd), n, dim, p_begin, p_end, ptr(pts), dtype_code(nd.dtype),
                     stream_ptr(nd.device))
        ctx.meta = (n, dim, p_begin, p_end, nd.dtype)
        return pts

    @staticmethod
    def backward(ctx, g):
        n, dim, p_begin, p_end, dtype = ctx.meta
        g = g.contiguous()
        gn = torch.zeros((dim, n), dtype=torch.float64, device=g.device)
        with on_device(g.device):
            call("tq_nc_grid_points_backward", ptr(g), n, dim, p_begin, p_end, ptr(gn), dtype_code(dtype),
                 stream_ptr(g.device))
        return gn.to(dtype), None, None


def nc_grid_points(nodes, p_begin=0, p_end=None):
    """points[p, d] = nodes[d, i_d(p)] (integration_grid.py:98-99 ordering); differentiable wrt nodes."""
    dim, n = nodes.shape
    p_end = n**dim if p_end is None else p_end
    return _GridPoints.apply(nodes, p_begin, p_end)


def nc_point_weights(w, p_begin, p_end):
    require_cuda(w)
    w = w.contiguous()
    dim, n = w.shape
    out = torch.empty(p_end - p_begin, dtype=w.dtype, device=w.device)
    if p_end > p_begin:
        with on_device(w.device):
            call("tq_nc_point_weights", ptr(w), n, dim, p_begin, p_end, ptr(out), dtype_code(w.dtype), stream_ptr(w.device))
    return out


class _Contract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, w, p_begin, p_end, keep_f64):
        require_cuda(f, w)
        fv = f.detach().contiguous()
        wv = w.detach().contiguous()
        dim, n = wv.shape
        rows = p_end - p_begin
        if fv.shape[0] != rows:
            raise ValueError(f"expected {rows} function values, got {fv.shape[0]}")
        cols = 1 if fv.dim() == 1 else int(torch.Size(fv.shape[1:]).numel())
        out = torch.zeros(cols, dtype=torch.float64, device=fv.device)
        if rows > 0:
            with on_device(fv.device):
                wsp, wsn = _ws(fv.device)
                call("tq_nc_contract", ptr(fv), ptr(wv), n, dim, p_begin, p_end, cols, dtype_code(fv.dtype), ptr(out),
                     wsp, wsn, stream_ptr(fv.device))
        ctx.save_for_backward(wv)
        ctx.meta = (p_begin, p_end, f.shape, fv.dtype)
        out = out if keep_f64 else out.to(fv.dtype)
        return out.reshape(f.shape[1:])

    @staticmethod
    def backward(ctx, g):
        (wv,) = ctx.saved_tensors
        p_begin, p_end, shape, dtype = ctx.meta
        W = nc_point_weights(wv, p_begin, p_end)
        gf = W.reshape([-1] + [1] * (len(shape) - 1)) * g.to(dtype).unsqueeze(0)
        return gf.expand(shape), None, None, None, None


def nc_contract(f, w, p_begin=0, p_end=None):
    """sum_p f[p, ...] * prod_d w[d, i_d(p)] in fp64, rounded once; differentiable wrt f."""
    dim, n = w.shape
    p_end = n**dim if p_end is None else p_end
    if f.is_complex():
        return _complex_through_real(_Contract.apply, f, w, p_begin, p_end, False)
    return _Contract.apply(f, w, p_begin, p_end, False)


def nc_contract_f64(f, w, p_begin, p_end):
    """Same contraction, unrounded fp64 partial sums (for accumulation over chunks and ranks)."""
    if f.is_complex():
        return _complex_through_real(_Contract.apply, f, w, p_begin, p_end, True)
    return _Contract.apply(f, w, p_begin, p_end, True)


# ------------------------------------------------------------------------------------------- fused
def fused_mc(fn_struct, dtype, device, row_begin, row_end, seed, call_idx):
    """{sum f, sum f^2} (fp64 [2]) of a built-in integrand over rows of the row-keyed stream."""
    out = torch.zeros(2, dtype=torch.float64, device=device)
    require_cuda(out)
    with on_device(out.device):
        wsp, wsn = _ws(out.device)
        call("tq_fused_mc", fn_struct, dtype_code(dtype), row_begin, row_end, seed & 0xFFFFFFFFFFFFFFFF,
             call_idx & 0xFFFFFFFF, ptr(out), wsp, wsn, stream_ptr(out.device))
    return out


def fused_nc(fn_struct, nodes, w, p_begin, p_end):
    require_cuda(nodes, w)
    dim, n = nodes.shape
    out = torch.zeros(1, dtype=torch.float64, device=nodes.device)
    with on_device(nodes.device):
        wsp, wsn = _ws(nodes.device)
        call("tq_fused_nc", fn_struct, ptr(nodes.contiguous()), ptr(w.contiguous()), n, dtype_code(nodes.dtype), p_begin,
             p_end, ptr(out), wsp, wsn, stream_ptr(nodes.device))
    return out


def pack_edges(x_edges, dx_edges, out=None):
    """Interleave {x_edges[d,k], dx_edges[d,k]} into [dim, Ni, 2] (the fused kernel's gather layout)."""
    require_cuda(x_edges, dx_edges)
    dim, ni = dx_edges.shape
    if out is None:
        out = torch.empty((dim, ni, 2), dtype=dx_edges.dtype, device=dx_edges.device)
    with on_device(dx_edges.device):
        call("tq_vegas_map_pack_edges", ptr(x_edges), ptr(dx_edges), ptr(out), dim, ni, dtype_code(dx_edges.dtype),
             stream_ptr(dx_edges.device))
    return out


def map_records_bytes(dim, ni, dtype):
    return _lib.load().tq_vegas_map_records_bytes(dim, ni, dtype_code(dtype))


def pack_records(x_edges, dx_edges, out=None):
    """{x_edge, dx_edge, 0, 0} records of a large map (see tq_vegas_map_pack_records): opaque uint8 [dim*Ni*rec]."""
    require_cuda(x_edges, dx_edges)
    dim, ni = dx_edges.shape
    if out is None:
        out = torch.empty(map_records_bytes(dim, ni, dx_edges.dtype), dtype=torch.uint8, device=dx_edges.device)
    with on_device(dx_edges.device):
        call("tq_vegas_map_pack_records", ptr(x_edges), ptr(dx_edges), ptr(out), dim, ni, dtype_code(dx_edges.dtype),
             stream_ptr(dx_edges.device))
    return out


def unpack_records(records, weights, counts):
    """weights += records.weight, counts += records.count, record fields back to zero."""
    require_cuda(records, weights, counts)
    dim, ni = weights.shape
    with on_device(weights.device):
        call("tq_vegas_map_unpack_records", ptr(records), ptr(weights), ptr(counts), dim, ni, dtype_code(weights.dtype),
             stream_ptr(weights.device))


def unpack_hist(hist_pairs, weights, counts):
    """weights += hist[..., 0] (rounded once), counts += hist[..., 1], hist = 0 (fp64 pair table of fused_vegas)."""
    require_cuda(hist_pairs, weights, counts)
    dim, ni = weights.shape
    with on_device(weights.device):
        call("tq_vegas_map_unpack_hist", ptr(hist_pairs), ptr(weights), ptr(counts), dim, ni, dtype_code(weights.dtype),
             stream_ptr(weights.device))


def fused_vegas(fn_struct, edges_packed, weights, counts, row_begin, row_end, seed, call_idx,
                offsets=None, n_strat=1, JF=None, JF2=None, records=None, dtype=None, n_intervals=None, hist_pairs=None):
    """One fused VEGAS pass (warm-up when offsets is None).  Returns fp64 [2] = {sum jf, sum jf^2} (warm-up only).
    With `records` (large maps) the histogram goes into the records and weights/counts are not touched; with
    `hist_pairs` (fp64 [dim, Ni, 2], see `unpack_hist`) it goes there instead of weights/counts."""
    require_cuda(edges_packed, weights, counts, offsets, JF, JF2, records, hist_pairs)
    if hist_pairs is not None:
        weights = counts = None
    table = records if records is not None else edges_packed
    dt = dtype if records is not None else edges_packed.dtype
    ni = n_intervals if records is not None else edges_packed.shape[1]
    # only the warm-up pass reduces {sum jf, sum jf^2}; the stratified pass writes JF/JF2
    out = torch.empty(2, dtype=torch.float64, device=table.device) if offsets is None else None
    n_cubes = 0 if offsets is None else offsets.shape[0] - 1
    with on_device(table.device):
        wsp, wsn = _ws(table.device)
        call("tq_fused_vegas", fn_struct, dtype_code(dt), ptr(offsets), n_cubes, n_strat, row_begin, row_end,
             ptr(table), _lib.TQ_EDGES_RECORDS if records is not None else _lib.TQ_EDGES_PAIRS, ni,
             None if records is not None or weights is None else ptr(weights),
             None if records is not None or weights is None else ptr(counts), ptr(hist_pairs), ptr(JF), ptr(JF2),
             seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, ptr(out), wsp, wsn, stream_ptr(table.device))
    return out


def fused_vegas_deferred(fn_struct, edges_packed, row_begin, row_end, seed, call_idx, offsets, n_strat, JF, JF2, jf2_rows):
    """A stratified fused pass WITHOUT the histogram: jf^2 of row r goes to jf2_rows[r - row_begin] (tq_fused_vegas_deferred)."""
    require_cuda(edges_packed, offsets, JF, JF2, jf2_rows)
    with on_device(edges_packed.device):
        wsp, wsn = _ws(edges_packed.device)
        call("tq_fused_vegas_deferred", fn_struct, dtype_code(edges_packed.dtype), ptr(offsets), offsets.shape[0] - 1, n_strat,
             row_begin, row_end, ptr(edges_packed), edges_packed.shape[1], ptr(jf2_rows), ptr(JF), ptr(JF2),
             seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, wsp, wsn, stream_ptr(edges_packed.device))


def hist_sweep(offsets, n_strat, dim, jf2_rows, n_intervals, hist_pairs, dims_per_group, seed, call_idx):
    """Bin the rows of a deferred pass band by band into the fp64 pair table (tq_vegas_hist_sweep)."""
    require_cuda(offsets, jf2_rows, hist_pairs)
    with on_device(offsets.device):
        call("tq_vegas_hist_sweep", ptr(offsets), offsets.shape[0] - 1, n_strat, dim, dtype_code(jf2_rows.dtype), ptr(jf2_rows),
             n_intervals, ptr(hist_pairs), dims_per_group, seed & 0xFFFFFFFFFFFFFFFF, call_idx & 0xFFFFFFFF, *_ws(offsets.device),
             stream_ptr(offsets.device))


def _vegas_state(vmap, strat, use_records, sweep=None):
    """(tq_vegas_state, keep-alive tensors) over the map / stratification tensors of a native-loop run.
    `sweep` = (dims_per_group, rows capacity) enables the deferred + band-sweep histogram of maps beyond L2."""
    dev, dt = vmap.device, vmap.dtype
    n_cubes = strat.N_cubes_local
    # JF and JF2 must be adjacent ([2, n_cubes]); the constructor's layout is, a user-replaced pair may not be
    if strat.JF.data_ptr() + n_cubes * strat.JF.element_size() != strat.JF2.data_ptr() or strat.JF.dtype != dt:
        pair = torch.empty((2, n_cubes), dtype=dt, device=dev)
        strat.JF, strat.JF2 = pair[0], pair[1]
    # every buffer below is fully written on the device before it is read (the driver clears what it needs)
    ints = torch.empty(2 * n_cubes + 1, dtype=torch.int64, device=dev)
    nh, offsets = ints[:n_cubes], ints[n_cubes:]
    records = torch.empty(_lib.TQ_VEGAS_MAX_PASSES * 4, dtype=torch.float64, device=dev)
    status = torch.empty(_lib.TQ_VEGAS_MAX_PASSES * 4, dtype=torch.int32, device=dev)
    scratch = _map_scratch(vmap.dim, vmap.N_intervals, dt, dev)
    ws = workspace(dev)
    packed = vmap.records() if use_records else vmap.packed_edges()
    hist = None if use_records else vmap.hist_pairs()
    jf2 = torch.empty(sweep[1], dtype=dt, device=dev) if sweep else None
    state = _lib.tq_vegas_state(
        ptr(vmap.x_edges), ptr(vmap.dx_edges), ptr(packed), ptr(vmap.weights), ptr(vmap.counts), ptr(hist), ptr(jf2),
        sweep[1] if sweep else 0, sweep[0] if sweep else 0, 0, ptr(strat.dh), ptr(nh),
        ptr(offsets), ptr(strat.JF), ptr(strat.JF2), ptr(records), ptr(status), ptr(scratch), scratch.numel(), ws.data_ptr(),
        ws.numel(), _lib.TQ_EDGES_RECORDS if use_records else _lib.TQ_EDGES_PAIRS)
    return state, (ints, records, status, scratch, packed, hist, jf2), nh, offsets


def _vegas_finish(vmap, strat, use_records, nh, offsets):
    if use_records:
        vmap._edges2_stale = True   # the pair table was not maintained; the records were
    else:
        vmap._records_stale = True
    strat._nh, strat._offsets = nh, offsets
    strat._counts_stale = True  # strat_counts = float(nh) is materialised on first access


def vegas_run_fused(fn_struct, vmap, strat, N, max_iterations, eps_rel, eps_abs, use_grid_improve, use_warmup, seed,
                    first_call, shard=None):
    """Whole fused VEGAS run through `tq_vegas_run_fused` (host loop in C++).  `vmap` / `strat` are the
    VEGASMap / VEGASStratification objects whose tensors hold the state (mutated in place).
    Returns the filled `tq_vegas_result`; synchronises (the schedule needs the per-block estimates).

    `shard` = (rank, world, log2 block, all_reduce) runs this rank's share of a multi-GPU job through
    `tq_vegas_run_fused_sharded`: `strat` then holds this rank's cubes only and `all_reduce(tensor)` must sum a fp64
    device tensor over the ranks in place on the current stream (torch.distributed.all_reduce)."""
    dev, dt = vmap.device, vmap.dtype
    use_records = bool(use_grid_improve) and vmap.wants_records()
    sweep = None
    if use_records and vmap.sweep_group(strat.N_strat) >= 1 and N // (max_iterations + 5) >= (1 << 20):
        # maps beyond L2: pair tables + deferred histogram, binned band by band (tq_vegas_hist_sweep); on several GPUs every rank
        # sweeps its own cubes.  A pass has at most budget * sum(dh) + 2 * cubes rows and the budget stays below 4 increments.
        use_records = False
        sweep = (vmap.sweep_group(strat.N_strat), 4 * (N // (max_iterations + 5)) + 2 * strat.N_cubes_local + 1024)
    state, keep, nh, offsets = _vegas_state(vmap, strat, use_records, sweep)
    result = _lib.tq_vegas_result()
    common = (fn_struct, dtype_code(dt), N, max_iterations, float(eps_rel), float(eps_abs), int(bool(use_grid_improve)),
              int(bool(use_warmup)), vmap.N_intervals, strat.N_strat, strat.N_cubes, float(strat.V_cubes), float(vmap.alpha),
              float(strat.beta), seed & 0xFFFFFFFFFFFFFFFF, first_call & 0xFFFFFFFF, state)
    if shard is None:
        with on_device(dev):
            call("tq_vegas_run_fused", *common, result, stream_ptr(dev))
    else:
        rank, world, lb, all_reduce = shard
        vmap.hist_pairs()
        comm = vmap._hist_flat
        comm.zero_()
        failure = []

        def _cb(_user, offset, count):
            try:
                all_reduce(comm[offset:offset + count])
                return 0
            except BaseException as exc:  # noqa: BLE001 - re-raised below
                failure.append(exc)
                return 1

        callback = _lib.tq_allreduce_callback(_cb)
        desc = _lib.tq_vegas_shard(rank, world, lb, 0, strat.N_cubes_local, ptr(comm), callback, None)
        try:
            with on_device(dev):
                call("tq_vegas_run_fused_sharded", *common, desc, result, stream_ptr(dev))
        except RuntimeError:
            if failure:
                raise failure[0]
            raise
    _vegas_finish(vmap, strat, use_records, nh, offsets)
    del keep
    return result


def vegas_run_unfused(evaluate, vmap, strat, domain, volume, cap_rows, N, max_iterations, eps_rel, eps_abs, use_grid_improve,
                      use_warmup, seed, first_call):
    """Whole VEGAS run with a callback integrand through `tq_vegas_run_unfused` (host loop in C++).
    `evaluate(x)` receives the [rows, dim] view of the sample buffer and returns the `rows` integrand values as a
    contiguous tensor of the working dtype on the same device (it runs on the current stream).  Exceptions raised
    by `evaluate` abort the run and propagate.  Returns the filled `tq_vegas_result`."""
    dev, dt = vmap.device, vmap.dtype
    dim = vmap.dim
    use_records = bool(use_grid_improve) and vmap.wants_records()
    state, keep, nh, offsets = _vegas_state(vmap, strat, use_records)
    x = torch.empty((cap_rows, dim), dtype=dt, device=dev)  # the only [rows, dim] buffer: y is never materialised
    jac = torch.empty(cap_rows, dtype=dt, device=dev)
    jf = torch.empty(cap_rows, dtype=dt, device=dev)
    warm = torch.tensor([[0.0, 0.999999]] * dim, dtype=dt, device=dev)
    dom = domain.detach().contiguous()
    buffers = _lib.tq_vegas_unfused_buffers(None, ptr(x), ptr(jac), ptr(jf), ptr(dom), ptr(warm), cap_rows, float(volume))
    failure = []
    last = [None]  # keeps the latest value tensor alive until the kernels that read it are queued behind it

    def _cb(_user, rows, f_out):
        try:
            f = evaluate(x[:rows])
            if f.dtype != dt or f.device != dev or f.numel() != rows:
                raise ValueError(f"integrand values must be {rows} values of dtype {dt} on {dev}, "
                                 f"got {tuple(f.shape)} / {f.dtype} / {f.device}")
            last[0] = f = f.contiguous()
            f_out[0] = f.data_ptr()
            return 0
        except BaseException as exc:  # noqa: BLE001 - re-raised by the caller below
            failure.append(exc)
            return 1

    callback = _lib.tq_eval_callback(_cb)
    result = _lib.tq_vegas_result()
    try:
        with on_device(dev):
            call("tq_vegas_run_unfused", callback, None, dim, dtype_code(dt), N, max_iterations, float(eps_rel), float(eps_abs),
                 int(bool(use_grid_improve)), int(bool(use_warmup)), vmap.N_intervals, strat.N_strat, strat.N_cubes,
                 float(strat.V_cubes), float(vmap.alpha), float(strat.beta), seed & 0xFFFFFFFFFFFFFFFF,
                 first_call & 0xFFFFFFFF, state, buffers, result, stream_ptr(dev))
    except RuntimeError:
        if failure:
            raise failure[0]
        raise
    _vegas_finish(vmap, strat, use_records, nh, offsets)
    del keep
    return result
