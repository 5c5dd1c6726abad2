"""Host-side mirror of the reference v0 network on the asr_b200 kernels.

Same method names, argument conventions and state-dict keys as
`UNet5` in the reference (models/v0/net_definitions_torch.py:390-686 configured
by models/v0/default.yaml:6-9: with_importance='all', normalized_channels=8,
residual_skip_connection=True), so a reference `state_dict` / `model.pt` loads
unchanged and scripts drive it the same way (`aggregate` -> `unet` -> `decode`,
train.py:61-86, asr.cpp:315-326).  All arithmetic runs in libasr_b200.so:

  * every SpecialSparseConv (+ gather of the importance, reduce_subarrays_sum,
    bias, ReLU; models/common_torch.py:95-148) is one fused `asr_sparse_conv`;
    the split first convolution `conv1a | conv1b` of the encoder blocks runs as a
    single convolution over the concatenated filter bank;
  * the neighbour tables are turned into slot-sorted plans once per grid and
    reused by the 4-9 convolutions that share them;
  * `levels` generalises the hard-wired 5 grids: extra levels
    re-use the 256-channel stage and `sparseconv_down3`, exactly as the reference
    does for its own fifth grid (:596-598).  levels=5 is UNet5.
"""
import torch

from . import ops

ENC_CHANNELS = [64, 128, 256, 256, 256]
NORMALIZED_CHANNELS = 8


def enc_channels(level):
    return ENC_CHANNELS[min(level, len(ENC_CHANNELS) - 1)]


class SpecialSparseConv(torch.nn.Module):
    """Parameter holder with the reference layer's names/shapes (common_torch.py:25-93)."""

    def __init__(self, in_channels, filters, kernel_size):
        super().__init__()
        if kernel_size not in (9, 55):
            raise Exception("kernel size must bei 9 or 55.")
        self.in_channels, self.filters, self.kernel_size = in_channels, filters, kernel_size
        self.kernel = torch.nn.Parameter(torch.empty(kernel_size, in_channels, filters).uniform_(-0.05, 0.05))
        self.bias = torch.nn.Parameter(torch.zeros(filters))


class _Block(torch.nn.Module):
    """SparseConvBlock (:123-302) / SparseConvTransitionBlock (:305-387) parameters.
    split=True is the normalized_channels=8 form with conv1a/conv1b."""

    def __init__(self, cin, cout, kernel_size, split, depth):
        super().__init__()
        self.kernel_size, self.split, self.depth = kernel_size, split, depth
        self.output_channels = cout
        if split:
            self.conv1a = SpecialSparseConv(cin, cout - NORMALIZED_CHANNELS, kernel_size)
            self.conv1b = SpecialSparseConv(cin, NORMALIZED_CHANNELS, kernel_size)
        else:
            self.conv1 = SpecialSparseConv(cin, cout, kernel_size)
        for j in range(2, depth + 1):
            setattr(self, "conv%d" % j, SpecialSparseConv(cout, cout, kernel_size))
        self._fused = None
        self._packed = {}
        self._gx = None

    def filters(self, j, K=ops):
        """Filter bank of conv j (1 = first conv, fused for the split form) in the
        form the active backend wants: packed hi/lo tiles for the tensor-core
        kernel (built once and cached), the fp32 tensor otherwise."""
        W = self.first_conv()[0] if j == 1 else getattr(self, "conv%d" % j).kernel
        if getattr(K, "PackedFilters", None) is None or K.SPARSE_CONV_BACKEND not in ("tensor", "gx") or W.shape[2] > 256:
            return W
        p = self._packed.get(j)
        if p is None:
            p = self._packed[j] = K.PackedFilters(W)
        return p

    def first_conv(self):
        """(kernel [K,Cin,Cout], bias [Cout]) of the first convolution; for the
        split form the two filter banks concatenated [plain | normalised]
        (channel order of torch.cat([feats1a, feats1b]), :283,378)."""
        if not self.split:
            return self.conv1.kernel, self.conv1.bias
        if self._fused is None:
            with torch.no_grad():
                self._fused = (torch.cat([self.conv1a.kernel, self.conv1b.kernel], 2).contiguous(),
                               torch.cat([self.conv1a.bias, self.conv1b.bias]).contiguous())
        return self._fused

    def run(self, x, plan, importance=None, K=ops):
        """K = kernel namespace: asr_b200.ops, or shard.ShardedOps for multi-GPU."""
        out_imp = None
        W, b = self.filters(1, K), self.first_conv()[1]
        if self.split:
            col = self.output_channels - NORMALIZED_CHANNELS
            out_imp = K.reduce_subarrays_sum(importance, plan.row_splits, index=plan.idx)
            x = K.sparse_conv(plan, W, x, inp_importance=importance, importance_col=col, normalize=True,
                              normalize_col=col, normalizer=out_imp, bias=b, relu=True)
        else:
            x = K.sparse_conv(plan, W, x, bias=b, relu=True)
        for j in range(2, self.depth + 1):
            c = getattr(self, "conv%d" % j)
            x = K.sparse_conv(plan, self.filters(j, K), x, bias=c.bias, relu=True)
        return x, out_imp


class ContinuousConv(torch.nn.Module):
    """Parameters of ml3d.layers.ContinuousConv as the reference builds it (:59-70)."""

    def __init__(self, in_channels, filters, kernel_size=(4, 4, 4)):
        super().__init__()
        self.kernel = torch.nn.Parameter(torch.empty(*kernel_size, in_channels, filters).uniform_(-0.05, 0.05))
        self.bias = torch.nn.Parameter(torch.zeros(filters))
        self.offset = torch.nn.Parameter(torch.zeros(3), requires_grad=False)


class CConvAggregationBlock(torch.nn.Module):

    def __init__(self, input_channels=4, output_channels=32):
        super().__init__()
        self.output_channels = output_channels
        self.conv1 = ContinuousConv(input_channels, output_channels)


class UNet(torch.nn.Module):
    """UNet5 generalised to `levels` grids (levels=5 == reference)."""

    def __init__(self, levels=5):
        super().__init__()
        if levels < 2:
            raise ValueError("levels must be >= 2")
        self.octree_levels = levels
        self.cconv_block_in = CConvAggregationBlock(4, 32)
        self.sparseconv_encblock0 = _Block(32, 64, 55, True, 4)
        for l in range(1, levels):
            if l <= 3:
                setattr(self, "sparseconv_down%d" % l, _Block(enc_channels(l - 1), enc_channels(l), 9, True, 1))
            setattr(self, "sparseconv_encblock%d" % l, _Block(enc_channels(l), enc_channels(l), 55, True, 4))
        prev = enc_channels(levels - 1)
        for l in range(levels - 2, -1, -1):
            up_out = 256 if l >= 1 else 64
            setattr(self, "sparseconv_up%d" % l, _Block(prev, up_out, 9, False, 1))
            if l >= 1:
                setattr(self, "sparseconv_decblock%d" % l, _Block(up_out + enc_channels(l), enc_channels(l), 55, False, 4))
                prev = enc_channels(l)
            else:
                setattr(self, "sparseconv_decblock0", _Block(up_out, 32, 55, False, 4))
        self.dense_decoder1 = torch.nn.Linear(35, 32, bias=True)
        self.dense_decoder2 = torch.nn.Linear(32, 32, bias=True)
        self.dense_decoder3 = torch.nn.Linear(32, 2, bias=False)
        self.requires_grad_(False)
        self.K = ops  # kernel namespace; shard.ShardedOps(ops, group) for the multi-GPU path

    def _down(self, level):
        return getattr(self, "sparseconv_down%d" % min(level, 3))

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        for m in self.modules():
            if isinstance(m, _Block):
                m._fused = None
                m._packed = {}
                m._gx = None
        return r

    # ------------------------------------------------------------------ reference methods
    def aggregate(self, input_dict):
        """UNet5.aggregate (:640-653) -> (feats [V0,32], per-PAIR importance [P])."""
        d, K = input_dict, self.K
        imp = K.aggregation_importance(d["aggregation_scale_compat"], d["aggregation_neighbors_dist"])
        c = self.cconv_block_in.conv1
        feats = K.continuous_conv(c.kernel, d["voxel_centers0"], d["voxel_sizes0"], c.offset, d["points"],
                                  d["feats"], None, d["aggregation_neighbors_index"], imp,
                                  d["aggregation_row_splits"], normalize=True, bias=c.bias, relu=True)
        return feats, K.pair_importance_for_unet(imp, d["voxel_centers0"].shape[0])

    def plans(self, input_dict):
        """Slot-sorted conv plans of all neighbour tables, cached in the dict."""
        cache = input_dict.get("_asr_plans")
        if cache is not None:
            return cache
        L = self.octree_levels
        d, K = input_dict, self.K
        if hasattr(K, "prepare"):  # multi-GPU: row ownership of the coarser grid levels
            K.prepare(d, L)
        P = {"nb": [], "up": [], "down": []}
        for i in range(L):
            P["nb"].append(K.ConvPlan(d["neighbors_index%d" % i], d["neighbors_kernel_index%d" % i],
                                      d["neighbors_row_splits%d" % i], 55))
        for i in range(L - 1):
            ui, uk, us = (d["up_neighbors_index%d" % i], d["up_neighbors_kernel_index%d" % i],
                          d["up_neighbors_row_splits%d" % i])
            P["up"].append(K.ConvPlan(ui, uk, us, 9))
            inv = K.invert_neighbors_list(d["voxel_centers%d" % (i + 1)].shape[0], ui, us, uk)
            P["down"].append(K.ConvPlan(inv.neighbors_index, inv.neighbors_attributes, inv.neighbors_row_splits, 9))
        input_dict["_asr_plans"] = P
        return P

    def unet(self, feats1, input_dict, taps=None):
        """UNet5.unet (:535-638).  taps: optional dict that receives the encoder output of every
        level ("enc<l>", the skip tensors) and every decoder block's output ("dec<l>") for parity tests."""
        L, K = self.octree_levels, self.K
        if K is ops and ops.SPARSE_CONV_BACKEND == "gx":
            from . import gx  # split-half activations, TMA-gather tcgen05 kernel (csrc/spconv_gx.cu)
            return gx.unet(self, feats1, input_dict, taps)
        P = self.plans(input_dict)
        x, imp = feats1
        skips = []
        x, imp = self.sparseconv_encblock0.run(x, P["nb"][0], imp, K)
        skips.append(x)
        for l in range(1, L):
            x, imp = self._down(l).run(x, P["down"][l - 1], imp, K)
            x, imp = getattr(self, "sparseconv_encblock%d" % l).run(x, P["nb"][l], imp, K)
            skips.append(x)
        if taps is not None:
            taps.update({"enc%d" % l: s for l, s in enumerate(skips)})
        for l in range(L - 2, -1, -1):
            x, _ = getattr(self, "sparseconv_up%d" % l).run(x, P["up"][l], None, K)
            x = K.cat([x, skips[l]], -1) if l >= 1 else K.add(x, skips[0])
            x, _ = getattr(self, "sparseconv_decblock%d" % l).run(x, P["nb"][l], None, K)
            if taps is not None:
                taps["dec%d" % l] = x
        return x

    def _decoder(self):
        return (self.dense_decoder1.weight, self.dense_decoder1.bias, self.dense_decoder2.weight,
                self.dense_decoder2.bias, self.dense_decoder3.weight)

    def decode(self, shifts, code, signed_scale=None):
        """UNet5.decode (:655-666); signed_scale fuses asr.cpp:334-336."""
        return self.K.decode(shifts, code, *self._decoder(), signed_scale=signed_scale)

    def decode_with_gradient(self, shifts, code):
        """UNet5.decode_with_gradient (:668-686)."""
        return self.K.decode(shifts, code, *self._decoder(), with_gradient=True)


def UNet5():
    return UNet(5)


def seeded_weights(net, seed=0, scaled=True):
    """Deterministic random weights for benchmarks/smoke runs (no checkpoint is
    available offline).  scaled=False: the reference initialisers (kernels
    U(-0.05, 0.05), zero biases, common_torch.py:57-58).  scaled=True:
    fan-in-scaled kernels and small biases so activations stay O(1) through the
    ~50 layers instead of decaying to zero."""
    import math
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("offset"):
                p.zero_()
                continue
            if name.endswith("kernel"):
                if scaled:
                    fan = p.shape[-2] * (7.7 if p.shape[0] == 55 else (16.0 if p.dim() == 5 else 1.0))
                    lim = math.sqrt(6.0 / fan)
                else:
                    lim = 0.05
            elif name.endswith("weight"):
                lim = 1.0 / math.sqrt(p.shape[1])
            else:  # biases
                lim = (1.0 / math.sqrt(35.0) if name.startswith("dense") else (0.1 if scaled else 0.0))
            p.copy_(((torch.rand(p.shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(p.dtype))
        for m in net.modules():
            if isinstance(m, _Block):
                m._fused = None
                m._packed = {}
                m._gx = None
    return net


def load_weights_file(path):
    """Reads the network weights from `path`: a TorchScript archive as written by the reference's
    models/v0/convert_tf2torchscript.py:117-122 (`model.pt`, loaded by asr.cpp:138-141 with torch::jit::load)
    or a plain `state_dict` file with the reference's key names.  Returns the state dict.

    The archive's graphs call `open3d::*` ops, so the shim that registers them
    (open3d.ml.torch.ops of this package) is imported before torch.jit.load — without it the load
    fails with "Unknown builtin op".  A file that is a zip archive with TorchScript code but fails to
    load raises that error (it is not retried as a pickle, which would hide the reason)."""
    import zipfile

    import open3d.ml.torch.ops  # noqa: F401  registers open3d::* (this package's shim, or real Open3D)
    is_script = False
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            is_script = any(n.endswith("constants.pkl") or "/code/" in n for n in z.namelist())
    if is_script:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    else:
        sd = torch.load(path, map_location="cpu")
        if not isinstance(sd, dict):
            sd = sd.state_dict()
    return {k: v for k, v in sd.items() if not k.startswith("_")}


def from_state_dict(state_dict, levels=5, device="cuda"):
    net = UNet(levels)
    net.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()})
    return net.to(device)
