"""TEST INFRASTRUCTURE ONLY — a SECOND, structurally different CPU restatement of Open3D-ML's `sparse_conv`
and `continuous_conv`, written the way SURVEY.md §8c items 2-3 recall Open3D v0.14.1's own
CPU implementation (`_SparseConvComputeFeaturesCPU` / `_CConvComputeFeaturesCPU`): outputs are processed in
blocks of 32 voxels; per block a dense matrix B[K * Cin, 32] is FILLED (sparse conv: assignment of imp * feat
into the slot's rows; continuous conv: trilinear splat of imp * feat over the 8 taps) and ONE GEMM
filter^T[Cout, K * Cin] @ B gives the block's outputs, followed by the normalisation.

oracle/ops_cpu.py instead gathers all pairs at once and accumulates per kernel slot with index_add.  The two
share no code path beyond torch's matmul; tests/test_oracle_model.py cross-checks them on the golden cloud.
This does not pin Open3D (absent from the image) but removes the single-formulation failure mode.
PARITY UNPINNED, like ops_cpu.py.  Plain loops: use on small inputs only.
"""
import numpy as np
import torch

BLOCK = 32


def sparse_conv(filters, inp_features, inp_importance, neighbors_index, neighbors_kernel_index,
                neighbors_importance, neighbors_row_splits, normalize):
    W = filters.double()
    K, Cin, Cout = W.shape
    Wf = W.reshape(K * Cin, Cout)
    x = inp_features.double().numpy()
    idx = neighbors_index.numpy().astype(np.int64)
    kidx = neighbors_kernel_index.numpy().astype(np.int64)
    rs = neighbors_row_splits.numpy()
    nimp = neighbors_importance.double().numpy() if neighbors_importance.numel() else None
    pimp = inp_importance.double().numpy() if inp_importance.numel() else None
    V = rs.shape[0] - 1
    out = torch.zeros((V, Cout), dtype=torch.float64)
    for b0 in range(0, V, BLOCK):
        nb = min(BLOCK, V - b0)
        B = np.zeros((K * Cin, nb))
        norm = np.zeros(nb)
        for col in range(nb):
            o = b0 + col
            for e in range(rs[o], rs[o + 1]):
                imp = 1.0
                if nimp is not None:
                    imp = nimp[e]
                if pimp is not None:
                    imp = imp * pimp[idx[e]]
                # assignment, not accumulation: a slot occurs at most once per row in the reference's tables
                B[kidx[e] * Cin:(kidx[e] + 1) * Cin, col] = imp * x[idx[e]]
                norm[col] += nimp[e] if nimp is not None else 1.0
        y = Wf.T @ torch.from_numpy(B)  # [Cout, nb]
        if normalize:
            nz = norm != 0
            y[:, nz] = y[:, nz] / torch.from_numpy(norm[nz])[None, :]
        out[b0:b0 + nb] = y.T
    return out


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, normalize=True):
    W = filters.double()
    Sz, Sy, Sx, Cin, Cout = W.shape
    Wf = W.reshape(Sz * Sy * Sx * Cin, Cout)
    size = np.array([Sx, Sy, Sz], np.float64)
    P = inp_positions.double().numpy()
    Q = out_positions.double().numpy()
    F = inp_features.double().numpy()
    ext = extents.double().numpy().reshape(-1)
    off = offset.double().numpy()
    idx = neighbors_index.numpy().astype(np.int64)
    rs = neighbors_row_splits.numpy()
    nimp = neighbors_importance.double().numpy() if neighbors_importance.numel() else None
    pimp = inp_importance.double().numpy() if inp_importance.numel() else None
    V = rs.shape[0] - 1
    out = torch.zeros((V, Cout), dtype=torch.float64)
    for b0 in range(0, V, BLOCK):
        nb = min(BLOCK, V - b0)
        B = np.zeros((Sz * Sy * Sx * Cin, nb))
        norm = np.zeros(nb)
        for col in range(nb):
            o = b0 + col
            e_o = ext[o] if ext.shape[0] > 1 else ext[0]
            for e in range(rs[o], rs[o + 1]):
                n = idx[e]
                imp = nimp[e] if nimp is not None else 1.0
                norm[col] += imp
                if pimp is not None:
                    imp = imp * pimp[n]
                x = (P[n] - Q[o]) * (2.0 / e_o)
                # ball_to_cube_radial: x *= 0.5 * |x|_2 / |x|_inf  (0 if |x|_inf < 1e-8)
                ninf = np.abs(x).max()
                x = x * (0.5 * np.sqrt((x * x).sum()) / ninf) if ninf >= 1e-8 else x * 0.0
                c = (x + off + 0.5) * (size - 1)  # align_corners
                c = np.minimum(np.maximum(c, 0.0), size - 1)
                i0 = np.minimum(np.floor(c), size - 1).astype(np.int64)
                i1 = np.minimum(i0 + 1, (size - 1).astype(np.int64))
                a = np.clip(c - i0, 0.0, 1.0)
                f = imp * F[n]
                for tz, wz in ((i0[2], 1 - a[2]), (i1[2], a[2])):
                    for ty, wy in ((i0[1], 1 - a[1]), (i1[1], a[1])):
                        for tx, wx in ((i0[0], 1 - a[0]), (i1[0], a[0])):
                            cell = (tz * Sy + ty) * Sx + tx
                            B[cell * Cin:(cell + 1) * Cin, col] += (wx * wy * wz) * f
        y = Wf.T @ torch.from_numpy(B)
        if normalize:
            nz = norm != 0
            y[:, nz] = y[:, nz] / torch.from_numpy(norm[nz])[None, :]
        out[b0:b0 + nb] = y.T
    return out
