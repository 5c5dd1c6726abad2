"""The U-Net's sparse convolutions on split-half activations ("gx" path, csrc/spconv_gx.cu).

Host side of `UNet5.unet` (reference models/v0/net_definitions_torch.py:535-638) for the fast path:
activations stay on the device between the ~53 convolutions as [V + 1, 2 C] float16 tensors (x = hi + lo,
last row zero), skip concatenations are channel slices of one buffer per level (torch.cat :607-631 becomes
"write into the slice"), the level-0 residual (:633-634) is fused into the up convolution's epilogue, and
every SpecialSparseConv (+ bias + ReLU + importance normalisation, models/common_torch.py:95-148) is one
`asr_gx_conv` call — two for the split first convolution of the encoder blocks (conv1a, and conv1b with the
per-row importance applied in the epilogue), one per 128 output columns for the 256-channel banks.  PyTorch only
owns the buffers and the stream.
"""
import ctypes as C
import math

import torch

from . import ops
from ._lib import check, lib

_i64 = C.c_int64
NORMALIZED_CHANNELS = 8


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class H2:
    """A [V, C] split-half activation view: rows of `buf` ([V + 1, pitch] float16), hi at columns
    [hi, hi + C), lo at [lo, lo + C)."""

    def __init__(self, buf, C_, hi, lo):
        self.buf, self.C, self.hi, self.lo = buf, C_, hi, lo
        self.V = buf.shape[0] - 1
        self.pitch = buf.shape[1]

    @staticmethod
    def empty(V, C_, device):
        buf = torch.empty((V + 1, 2 * C_), dtype=torch.float16, device=device)
        buf[V:].zero_()
        return H2(buf, C_, 0, C_)

    def slice(self, c0, C_):
        """channels [c0, c0 + C) of this view as a view of the same buffer"""
        return H2(self.buf, C_, self.hi + c0, self.lo + c0)

    def args(self):
        return (_ptr(self.buf), self.pitch, self.hi, self.lo)

    def to_f32(self):
        out = torch.empty((self.V, self.C), dtype=torch.float32, device=self.buf.device)
        check(lib().asr_gx_to_f32(_ptr(self.buf), self.V, self.C, self.pitch, self.hi, self.lo, _ptr(out), self.C,
                                  _stream()))
        return out


def from_f32(x, row_scale=None, out=None, rows=None):
    """fp32 [n, C] -> split-half.  rows (int32, with `out` given): input row i goes to row rows[i] of `out`, no other
    row of `out` is touched."""
    x = ops._cuda(x, torch.float32, "x")
    V, C_ = x.shape
    out = out or H2.empty(V, C_, x.device)
    check(lib().asr_gx_from_f32(_ptr(x), V, C_, C_, _ptr(row_scale), _ptr(rows), out.V, *out.args(), _stream()))
    return out


def scale_rows(x, row_scale, out=None):
    out = out or H2.empty(x.V, x.C, x.buf.device)
    check(lib().asr_gx_scale_rows(_ptr(x.buf), x.V, x.C, x.pitch, x.hi, x.lo, _ptr(row_scale), *out.args(), _stream()))
    return out


MODE_STATIONARY, MODE_PAIR_FINAL = 0, 1


class Plan:
    """Gather tables / pair tiles of one neighbour table (begin now, finish() after all plans were begun)."""

    def __init__(self, idx, slot, row_splits, num_in, kernel_size, mode, rows=None):
        """rows (int32, ascending): build the plan for these table rows only (this rank's rows of a sharded level);
        the convolution then writes exactly these rows of its (full-size) output."""
        self.idx = ops._cuda(idx, torch.int32, "neighbors_index")
        self.slot = ops._cuda(slot, torch.uint8, "neighbors_kernel_index")
        self.row_splits = ops._cuda(row_splits, torch.int64, "neighbors_row_splits")
        self.rows = None if rows is None else ops._cuda(rows, torch.int32, "rows")
        self.table_rows = self.row_splits.shape[0] - 1
        self.num_out = self.table_rows if rows is None else int(self.rows.shape[0])
        self.num_in = int(num_in)
        self.kernel_size = int(kernel_size)
        self.mode = mode
        self.num_rare = None
        self._h = C.c_void_p(0)
        check(lib().asr_gx_plan_begin(_ptr(self.idx), _ptr(self.slot), _ptr(self.row_splits), self.num_out, self.num_in,
                                      self.idx.shape[0], self.kernel_size, mode, _ptr(self.rows), _stream(),
                                      C.byref(self._h)))

    def finish(self):
        if self.num_rare is None:
            n = _i64(0)
            check(lib().asr_gx_plan_finish(self._h, _stream(), C.byref(n)))
            self.num_rare = n.value
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().asr_gx_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class Filters:
    """Output columns [col0, col0 + ncols) of a [K, Cin, Cout] filter bank, packed as fp16 hi/lo of
    W * 2^scale_exp (scale_exp puts max|W| near 2^8 so that the lo parts stay normal fp16 numbers)."""

    def __init__(self, W, col0=0, ncols=None, bias=None):
        W = ops._cuda(W, torch.float32, "filters")
        K, Cin, Cout = W.shape
        ncols = Cout - col0 if ncols is None else ncols
        m = float(W[:, :, col0:col0 + ncols].abs().max()) if W.numel() else 0.0
        self.scale_exp = int(8 - math.ceil(math.log2(m))) if m > 0 and math.isfinite(m) else 0
        self.K, self.Cin, self.ncols, self.col0 = K, Cin, ncols, col0
        n = int(lib().asr_gx_packed_filters_bytes(K, Cin, ncols))
        self.data = torch.empty(n, dtype=torch.uint8, device=W.device)
        check(lib().asr_gx_pack_filters(_ptr(W), K, Cin, Cout, col0, ncols, self.scale_exp, _ptr(self.data), _stream()))
        self.bias = None if bias is None else ops._cuda(bias, torch.float32, "bias")[col0:col0 + ncols].contiguous()


class Scratch:
    """Pair buffer shared by the convolutions of one U-Net pass (stream-ordered reuse)."""

    def __init__(self):
        self.buf = None

    def get(self, n, device):
        if n == 0:
            return None
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(n, dtype=torch.float32, device=device)
        return self.buf


def conv(plan, x, filt, relu=True, norm=None, res=None, out=None, out_f32=None, scratch=None, imp=None, alloc=None):
    """out[:, 0:ncols] = act(sparse_conv(x) (/ norm) + bias) (+ res); `out` an H2 view or `out_f32` a float32
    [V, ncols] tensor.  imp: importance of every INPUT row (weights each gathered row, in fp32 in the epilogue)."""
    if isinstance(filt, (list, tuple)):
        # column groups of one filter bank (banks wider than 128 columns run as 128-column passes: every pass keeps
        # its main accumulation chain split over two TMEM accumulators, see the kernel)
        total = sum(f.ncols for f in filt)
        if out is None and out_f32 is None:
            out = (alloc or H2.empty)(plan.table_rows, total, x.buf.device)
        account, ops.ACCOUNT = ops.ACCOUNT, None
        try:
            for f in filt:
                conv(plan, x, f, relu=relu, norm=norm, imp=imp, scratch=scratch,
                     res=None if res is None else res.slice(f.col0, f.ncols),
                     out=None if out is None else out.slice(f.col0, f.ncols),
                     out_f32=None if out_f32 is None else out_f32[:, f.col0:])
        finally:
            ops.ACCOUNT = account
        if ops.ACCOUNT is not None:
            ops.ACCOUNT.append({"V_in": plan.num_in, "V_out": plan.num_out, "E": plan.idx.shape[0], "K": filt[0].K,
                                "Cin": filt[0].Cin, "Cout": total, "importance": norm is not None})
        return out if out is not None else out_f32
    if x.C != filt.Cin or plan.kernel_size != filt.K or x.V != plan.num_in:
        raise ValueError("gx.conv: shapes of the input / filters / plan do not match")
    if filt.ncols > 128:
        raise ValueError("gx.conv: more than 128 output columns per call — pack the bank with gx.filter_bank")
    plan.finish()
    npad = (filt.ncols + 15) // 16 * 16
    need = plan.num_rare * npad
    pb = (scratch or Scratch()).get(need, x.buf.device)
    if out is None and out_f32 is None:
        out = (alloc or H2.empty)(plan.table_rows, filt.ncols, x.buf.device)
    rp = res.args() if res is not None else (C.c_void_p(0), 0, 0, 0)
    op = out.args() if out is not None else (C.c_void_p(0), 0, 0, 0)
    check(lib().asr_gx_conv(plan._h, _ptr(x.buf), x.C, x.pitch, x.hi, x.lo, _ptr(filt.data), filt.ncols, filt.scale_exp,
                            _ptr(filt.bias), int(bool(relu)), _ptr(norm), _ptr(imp), *rp, *op, _ptr(out_f32),
                            out_f32.stride(0) if out_f32 is not None else 0, _ptr(pb), _stream()))
    if ops.ACCOUNT is not None:
        ops.ACCOUNT.append({"V_in": plan.num_in, "V_out": plan.num_out, "E": plan.idx.shape[0], "K": filt.K,
                            "Cin": filt.Cin, "Cout": filt.ncols, "importance": norm is not None})
    return out if out is not None else out_f32


def overflow():
    f = C.c_int(0)
    check(lib().asr_gx_overflow(_stream(), C.byref(f)))
    return bool(f.value)


def filter_bank(W, bias=None, max_cols=128):
    """Filters of a whole bank; banks wider than `max_cols` as a list of column groups."""
    cout = W.shape[2]
    if cout <= max_cols:
        return Filters(W, 0, cout, bias)
    return [Filters(W, c0, min(max_cols, cout - c0), bias) for c0 in range(0, cout, max_cols)]


def _ncols(f):
    return sum(g.ncols for g in f) if isinstance(f, (list, tuple)) else f.ncols


# ------------------------------------------------------------------------------------ the U-Net on gx
def _block_filters(block):
    """Packed filter banks of a model._Block, cached on the block: [(plain, normalised or None), conv2, ...]."""
    cache = getattr(block, "_gx", None)
    if cache is not None:
        return cache
    W1, b1 = block.first_conv()
    cout = W1.shape[2]
    first = filter_bank(W1, b1)
    firstb = Filters(W1, cout - NORMALIZED_CHANNELS, NORMALIZED_CHANNELS, b1) if block.split else None
    rest = []
    for j in range(2, block.depth + 1):
        c = getattr(block, "conv%d" % j)
        rest.append(filter_bank(c.kernel, c.bias))
    block._gx = (first, firstb, rest)
    return block._gx


class LocalContext:
    """Where the U-Net's buffers live and what happens after a convolution.  Single GPU: torch allocations, nothing
    to do.  shard_gx.ShardContext allocates from a symmetric arena and pushes the halo rows to the peers."""

    def empty(self, V, C_, device):
        return H2.empty(V, C_, device)

    def done(self, view, level):
        pass

    def mark(self):
        """allocation mark / release of everything allocated after it (arena contexts; torch frees by itself)"""
        return None

    def release(self, mark):
        pass

    def plans(self, input_dict, levels):
        return build_plans(input_dict, levels)


def run_block(block, x, plan, importance, scratch, out=None, out_f32=None, res=None, ctx=None, level=None):
    """model._Block.run on H2 activations.  `out` / `out_f32` / `res` apply to the block's LAST convolution."""
    ctx = ctx or LocalContext()
    first, firstb, rest = _block_filters(block)
    out_imp = None
    if out is None and out_f32 is None:
        # the block's result is allocated before its temporaries, which are released at the end of the block
        out = ctx.empty(plan.table_rows, _ncols(rest[-1] if rest else first), x.buf.device)
    mark = ctx.mark()
    last = not rest
    y = conv(plan, x, first, out=out if last else None, out_f32=out_f32 if last else None,
             res=res if last else None, scratch=scratch, alloc=ctx.empty)
    if block.split:
        # conv1b (common_torch.py:124-142 with normalize=True): importance-weighted input, divided by the summed
        # importance of the row; lands in the last 8 channels of the block's first activation (:283,378)
        out_imp = ops.reduce_subarrays_sum(importance, plan.row_splits, index=plan.idx)
        col = _ncols(first) - NORMALIZED_CHANNELS
        if y is out_f32 and out_f32 is not None:
            raise ValueError("a split block cannot end in an fp32 output")
        conv(plan, x, firstb, norm=out_imp, imp=importance, out=y.slice(col, NORMALIZED_CHANNELS), scratch=scratch)
    if isinstance(y, H2):
        ctx.done(y, level)
    for i, f in enumerate(rest):
        last = i == len(rest) - 1
        y = conv(plan, y, f, out=out if last else None, out_f32=out_f32 if last else None, res=res if last else None,
                 scratch=scratch, alloc=ctx.empty)
        if isinstance(y, H2):
            ctx.done(y, level)
    ctx.release(mark)
    return y, out_imp


def _begin_plans(input_dict, levels):
    d = input_dict
    V = [d["neighbors_row_splits%d" % i].shape[0] - 1 for i in range(levels)]
    P = {"nb": [], "up": [], "down": [], "V": V}
    for i in range(levels):
        P["nb"].append(Plan(d["neighbors_index%d" % i], d["neighbors_kernel_index%d" % i],
                            d["neighbors_row_splits%d" % i], V[i], 55, MODE_STATIONARY))
    for i in range(levels - 1):
        ui, uk, us = d["up_neighbors_index%d" % i], d["up_neighbors_kernel_index%d" % i], d["up_neighbors_row_splits%d" % i]
        P["up"].append(Plan(ui, uk, us, V[i + 1], 9, MODE_PAIR_FINAL))
        inv = ops.invert_neighbors_list(V[i + 1], ui, us, uk)
        P["down"].append(Plan(inv.neighbors_index, inv.neighbors_attributes, inv.neighbors_row_splits, V[i], 9,
                              MODE_STATIONARY))
    return P


def _finish_plans(P):
    for group in ("nb", "up", "down"):
        for p in P[group]:
            p.finish()


def build_plans(input_dict, levels):
    """gx plans of all neighbour tables (cached in the dict); ONE host synchronisation for all of them.
    If begin_plans_async() started them on the side stream, this joins that work."""
    cache = input_dict.get("_asr_gx_plans")
    if cache is not None:
        return cache
    pending = input_dict.pop("_asr_gx_plans_pending", None)
    if pending is not None:
        finish_plans_async(input_dict, pending)
        P, side, done = pending["plans"], pending["side"], pending["done"]
        torch.cuda.current_stream().wait_event(done)
    else:
        P = _begin_plans(input_dict, levels)
        _finish_plans(P)
    input_dict["_asr_gx_plans"] = P
    return P


def begin_plans_async(input_dict, levels, side):
    """Start the plans of all tables on `side` (the tables were produced on the current stream): the many small
    kernels of the plan construction then run beside the aggregation search instead of in front of the U-Net."""
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        P = _begin_plans(input_dict, levels)
    input_dict["_asr_gx_plans_pending"] = {"plans": P, "side": side, "done": None}


def finish_plans_async(input_dict, pending=None):
    """Second phase on the side stream (call it once the host has passed a synchronisation point after
    begin_plans_async, e.g. after the search: the per-table counts are then on the host already)."""
    pending = pending or input_dict.get("_asr_gx_plans_pending")
    if pending is None or pending["done"] is not None:
        return
    with torch.cuda.stream(pending["side"]):
        _finish_plans(pending["plans"])
        pending["done"] = torch.cuda.Event()
        pending["done"].record(pending["side"])


def unet(net, feats1, input_dict, taps=None, ctx=None, x0=None, code=None):
    """UNet5.unet (:535-638) for `net.octree_levels` grids on the gx kernels.  Returns code [V0, 32] float32.
    ctx: buffer / halo context (default: single GPU); x0: the aggregated features already in split-half form."""
    from .model import enc_channels
    ctx = ctx or LocalContext()
    L = net.octree_levels
    P = ctx.plans(input_dict, L)
    V = P["V"]
    feats, imp = feats1
    dev = imp.device
    scratch = Scratch()
    x = x0 if x0 is not None else from_f32(feats)
    # encoder: the output of every level except the deepest is the second part of that level's decoder input
    # (torch.cat([up, skip]), :607-631), so it is written straight into that buffer
    cat = [None] * L
    skips = [None] * L
    for l in range(L):
        C_l = enc_channels(l)
        if l >= 1:
            x, imp = run_block(net._down(l), x, P["down"][l - 1], imp, scratch, ctx=ctx, level=l)
        block = net.sparseconv_encblock0 if l == 0 else getattr(net, "sparseconv_encblock%d" % l)
        out = None
        if 1 <= l <= L - 2:
            cat[l] = ctx.empty(V[l], 256 + C_l, dev)
            out = cat[l].slice(256, C_l)
        x, imp = run_block(block, x, P["nb"][l], imp, scratch, out=out, ctx=ctx, level=l)
        skips[l] = x
    if taps is not None:
        taps.update({"enc%d" % l: s.to_f32() for l, s in enumerate(skips)})
    for l in range(L - 2, -1, -1):
        up = getattr(net, "sparseconv_up%d" % l)
        dec = getattr(net, "sparseconv_decblock%d" % l)
        if l >= 1:
            run_block(up, x, P["up"][l], None, scratch, out=cat[l].slice(0, 256), ctx=ctx, level=l)
            x, _ = run_block(dec, cat[l], P["nb"][l], None, scratch, ctx=ctx, level=l)
            if taps is not None:
                taps["dec%d" % l] = x.to_f32()
        else:
            x, _ = run_block(up, x, P["up"][0], None, scratch, res=skips[0], ctx=ctx, level=0)  # feats20 + feats2 (:633-634)
            if code is None:
                code = torch.empty((V[0], 32), dtype=torch.float32, device=dev)
            run_block(dec, x, P["nb"][0], None, scratch, out_f32=code, ctx=ctx, level=0)
            if taps is not None:
                taps["dec0"] = code
    return code
