"""`open3d.ml.torch.layers.ContinuousConv` as the reference constructs and calls
it (net_definitions_torch.py:59-70,108-116): parameters `kernel`
[*kernel_size, in_channels, filters], `bias` [filters], `offset` [3];
user-supplied neighbour lists; per-output extents."""
import torch

from . import ops


class ContinuousConv(torch.nn.Module):

    def __init__(self, in_channels, filters, kernel_size, activation=None, use_bias=True,
                 kernel_initializer=lambda x: torch.nn.init.uniform_(x, -0.05, 0.05),
                 bias_initializer=torch.nn.init.zeros_, align_corners=True, coordinate_mapping="ball_to_cube_radial",
                 interpolation="linear", normalize=True, radius_search_ignore_query_points=False,
                 radius_search_metric="L2", offset=None, window_function=None, use_dense_layer_for_center=False,
                 dense_kernel_initializer=None, dense_kernel_regularizer=None, in_importance=False, **kwargs):
        super().__init__()
        if use_dense_layer_for_center or window_function is not None:
            raise NotImplementedError("not used by the reference model")
        self.in_channels = in_channels
        self.filters = filters
        self.kernel_size = list(kernel_size)
        self.activation = activation if activation is not None else (lambda x: x)
        self.use_bias = use_bias
        self.align_corners = align_corners
        self.coordinate_mapping = coordinate_mapping
        self.interpolation = interpolation
        self.normalize = normalize
        self.kernel = torch.nn.Parameter(torch.empty(*self.kernel_size, in_channels, filters))
        kernel_initializer(self.kernel)
        if use_bias:
            self.bias = torch.nn.Parameter(torch.empty(filters))
            bias_initializer(self.bias)
        off = torch.zeros(3) if offset is None else torch.as_tensor(offset, dtype=torch.float32)
        self.offset = torch.nn.Parameter(off, requires_grad=False)

    def forward(self, inp_features, inp_positions, out_positions, extents, inp_importance=None,
                fixed_radius_search_hash_table=None, user_neighbors_index=None, user_neighbors_row_splits=None,
                user_neighbors_importance=None):
        if user_neighbors_index is None or user_neighbors_row_splits is None:
            raise NotImplementedError("asr_b200 ContinuousConv needs user_neighbors_index/_row_splits "
                                      "(the reference always supplies them)")
        if not isinstance(extents, torch.Tensor):
            extents = torch.as_tensor([float(extents)], dtype=torch.float32, device=inp_features.device)
        empty = torch.empty((0,), dtype=torch.float32, device=inp_features.device)
        out = ops.continuous_conv(self.kernel, out_positions, extents, self.offset, inp_positions, inp_features,
                                  empty if inp_importance is None else inp_importance, user_neighbors_index,
                                  empty if user_neighbors_importance is None else user_neighbors_importance,
                                  user_neighbors_row_splits, align_corners=self.align_corners,
                                  coordinate_mapping=self.coordinate_mapping, normalize=self.normalize,
                                  interpolation=self.interpolation)
        if self.use_bias:
            out = out + self.bias
        return self.activation(out)
