"""TEST INFRASTRUCTURE ONLY — functional CPU restatement of the v0 network
(models/v0/net_definitions_torch.py:390-686 with models/v0/default.yaml:6-9:
with_importance='all', normalized_channels=8, residual_skip_connection=True).

Parity status: the *topology* is pinned — tests/test_oracle_model.py runs the
reference's own net_definitions_torch.py (imported from /root/reference in the
build container, through the oracle-backed open3d stand-in in oracle/o3d_shim)
and checks this restatement against it, and tests/golden/ holds vectors
generated that way.  The *op arithmetic* underneath (oracle/ops_cpu.py) is
unpinned, see there.

`levels` generalises the hard-wired 5-grid network (SURVEY.md §8d "N levels"):
level l >= 3 re-uses the 256-channel stage and `sparseconv_down3` (the reference
already does this for its fifth grid, net_definitions_torch.py:596-598).
levels == 5 is exactly UNet5.
"""
import math

import torch

from . import ops_cpu as ops

ENC_CHANNELS = [64, 128, 256, 256, 256]


def enc_channels(level):
    return ENC_CHANNELS[min(level, len(ENC_CHANNELS) - 1)]


def down_name(level):
    """Transition block that produces grid `level` (>=1)."""
    return "sparseconv_down%d" % min(level, 3)


def conv_specs(levels=5, normalized_channels=8):
    """[(state-dict prefix, K, Cin, Cout)] of every SpecialSparseConv, in
    execution order (shared `sparseconv_down3` listed once)."""
    specs = []
    nc = normalized_channels

    def block(prefix, cin, cout, split):
        if split:
            specs.append((prefix + ".conv1a", 55, cin, cout - nc))
            specs.append((prefix + ".conv1b", 55, cin, nc))
        else:
            specs.append((prefix + ".conv1", 55, cin, cout))
        for j in (2, 3, 4):
            specs.append((prefix + ".conv%d" % j, 55, cout, cout))

    block("sparseconv_encblock0", 32, 64, True)
    seen = set()
    for l in range(1, levels):
        cin, cout = enc_channels(l - 1), enc_channels(l)
        dn = down_name(l)
        if dn not in seen:
            seen.add(dn)
            specs.append((dn + ".conv1a", 9, cin, cout - nc))
            specs.append((dn + ".conv1b", 9, cin, nc))
        block("sparseconv_encblock%d" % l, cout, cout, True)
    prev = enc_channels(levels - 1)
    for l in range(levels - 2, -1, -1):
        up_out = 256 if l >= 1 else 64
        specs.append(("sparseconv_up%d.conv1" % l, 9, prev, up_out))
        if l >= 1:
            block("sparseconv_decblock%d" % l, up_out + enc_channels(l), enc_channels(l), False)
            prev = enc_channels(l)
        else:
            block("sparseconv_decblock0", up_out, 32, False)
            prev = 32
    return specs


def init_params(levels=5, seed=0, stress=False, dtype=torch.float32):
    """Seeded random weights with the reference initialisers: conv kernels
    U(-0.05, 0.05), biases 0 (common_torch.py:57-58), torch.nn.Linear defaults
    for the decoder (net_definitions_torch.py:503-511).  `stress=True` uses
    He-style scaling and non-zero biases so activations survive 30+ layers."""
    g = torch.Generator().manual_seed(seed)
    P = {}

    def uni(shape, lim):
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)

    P["cconv_block_in.conv1.kernel"] = uni((4, 4, 4, 4, 32), 0.5 if stress else 0.05)
    P["cconv_block_in.conv1.bias"] = uni((32,), 0.1) if stress else torch.zeros(32, dtype=dtype)
    P["cconv_block_in.conv1.offset"] = torch.zeros(3, dtype=dtype)
    for name, K, cin, cout in conv_specs(levels):
        lim = math.sqrt(6.0 / (cin * (7.7 if K == 55 else 1.0))) if stress else 0.05
        P[name + ".kernel"] = uni((K, cin, cout), lim)
        P[name + ".bias"] = uni((cout,), 0.1) if stress else torch.zeros(cout, dtype=dtype)
    for i, (fi, fo, bias) in enumerate([(35, 32, True), (32, 32, True), (32, 2, False)], 1):
        lim = 1.0 / math.sqrt(fi)
        P["dense_decoder%d.weight" % i] = uni((fo, fi), lim)
        if bias:
            P["dense_decoder%d.bias" % i] = uni((fo,), lim)
    return P


def special_sparse_conv(P, name, x, nb, importance=None, normalize=False, dtype=None):
    """SpecialSparseConv.forward (common_torch.py:95-148) with ReLU."""
    idx, kidx, rs = nb
    empty = torch.empty(0, dtype=torch.float32)
    if importance is not None:
        nimp = importance[idx.to(torch.int64)]
        out_imp = ops.reduce_subarrays_sum(nimp, rs)
    else:
        nimp, out_imp = empty, None
    y = ops.sparse_conv(P[name + ".kernel"], x, empty, idx, kidx, nimp, rs, normalize, dtype=dtype)
    y = torch.relu(y + P[name + ".bias"].to(y.dtype))
    return y, out_imp


def conv_block(P, prefix, x, nb, importance=None, split=False, dtype=None):
    """SparseConvBlock / SparseConvTransitionBlock (net_definitions_torch.py:
    123-387).  split=True is the `normalized_channels=8` encoder form: plain
    conv1a ‖ importance-normalised conv1b, concatenated [plain | normalised]."""
    out_imp = None
    if split:
        a, _ = special_sparse_conv(P, prefix + ".conv1a", x, nb, dtype=dtype)
        b, out_imp = special_sparse_conv(P, prefix + ".conv1b", x, nb, importance, True, dtype=dtype)
        y = torch.cat([a, b], -1)
    else:
        y, _ = special_sparse_conv(P, prefix + ".conv1", x, nb, dtype=dtype)
    for j in (2, 3, 4):
        if (prefix + ".conv%d.kernel" % j) in P:  # transition blocks have conv1 only
            y, _ = special_sparse_conv(P, prefix + ".conv%d" % j, y, nb, dtype=dtype)
    return y, out_imp


def aggregate(P, inp, dtype=None):
    """UNet5.aggregate -> CConvAggregationBlock (net_definitions_torch.py:
    72-120, 640-653).  Returns (feats[V0,32], per-PAIR importance[P])."""
    imp = inp["aggregation_scale_compat"] * ops.window_poly6(inp["aggregation_neighbors_dist"])
    empty = torch.empty(0, dtype=torch.float32)
    y = ops.continuous_conv(P["cconv_block_in.conv1.kernel"], inp["voxel_centers0"], inp["voxel_sizes0"],
                            P["cconv_block_in.conv1.offset"], inp["points"], inp["feats"], empty,
                            inp["aggregation_neighbors_index"], imp, inp["aggregation_row_splits"],
                            normalize=True, dtype=dtype)
    y = torch.relu(y + P["cconv_block_in.conv1.bias"].to(y.dtype))
    return y, imp


def unet(P, feats_and_importance, inp, levels=5, dtype=None, taps=None):
    """UNet5.unet (net_definitions_torch.py:535-638) for `levels` grids.  taps: optional dict that
    receives the per-level encoder outputs ("enc<l>") and decoder block outputs ("dec<l>")."""
    nb = [(inp["neighbors_index%d" % i], inp["neighbors_kernel_index%d" % i], inp["neighbors_row_splits%d" % i])
          for i in range(levels)]
    up = [(inp["up_neighbors_index%d" % i], inp["up_neighbors_kernel_index%d" % i],
           inp["up_neighbors_row_splits%d" % i]) for i in range(levels - 1)]
    down = []
    for i in range(levels - 1):
        r = ops.invert_neighbors_list(inp["voxel_centers%d" % (i + 1)].shape[0], *[up[i][0], up[i][2], up[i][1]])
        down.append((r.neighbors_index, r.neighbors_attributes, r.neighbors_row_splits))

    x, imp = feats_and_importance
    skips = []
    x, imp = conv_block(P, "sparseconv_encblock0", x, nb[0], imp, split=True, dtype=dtype)
    skips.append(x)
    for l in range(1, levels):
        x, imp = conv_block(P, down_name(l), x, down[l - 1], imp, split=True, dtype=dtype)
        x, imp = conv_block(P, "sparseconv_encblock%d" % l, x, nb[l], imp, split=True, dtype=dtype)
        skips.append(x)
    if taps is not None:
        taps.update({"enc%d" % l: s for l, s in enumerate(skips)})
    for l in range(levels - 2, -1, -1):
        x, _ = conv_block(P, "sparseconv_up%d" % l, x, up[l], dtype=dtype)
        x = torch.cat([x, skips[l]], -1) if l >= 1 else x + skips[0]
        x, _ = conv_block(P, "sparseconv_decblock%d" % l, x, nb[l], dtype=dtype)
        if taps is not None:
            taps["dec%d" % l] = x
    return x


def decode(P, shifts, code):
    """UNet5.decode (net_definitions_torch.py:655-666)."""
    dt = code.dtype
    h = torch.cat([shifts.to(dt), code], -1)
    h = torch.relu(h @ P["dense_decoder1.weight"].to(dt).T + P["dense_decoder1.bias"].to(dt))
    h = torch.relu(h @ P["dense_decoder2.weight"].to(dt).T + P["dense_decoder2.bias"].to(dt))
    return h @ P["dense_decoder3.weight"].to(dt).T


def decode_with_gradient(P, shifts, code):
    """UNet5.decode_with_gradient (:668-686): d value[:,0] / d shift."""
    dt = code.dtype
    W1, W2, W3 = (P["dense_decoder%d.weight" % i].to(dt) for i in (1, 2, 3))
    h0 = torch.cat([shifts.to(dt), code], -1)
    h1 = torch.relu(h0 @ W1.T + P["dense_decoder1.bias"].to(dt))
    h2 = torch.relu(h1 @ W2.T + P["dense_decoder2.bias"].to(dt))
    value = h2 @ W3.T
    z3 = torch.ones(shifts.shape[0], 1, dtype=dt) * W3[:1, :]
    z3[h2 <= 0] = 0
    z2 = z3 @ W2
    z2[h1 <= 0] = 0
    z1 = z2 @ W1
    return value, z1[:, :3]
