"""mma_b200 -- B200-native (sm_100a) multi-mask aggregation hot path of asarigun/mma.

Drop-in layers (reference signatures unchanged):
    mma_b200.graph_regression.mma_conv.MMAConv            <- graph_regression/mma_conv.py
    mma_b200.graph_regression.mask_aggr.MaskAggregateLinear <- graph_regression/mask_aggr.py
    mma_b200.node_classification.layers.MMA, GraphConvolution <- node_classification/layers.py
    mma_b200.node_classification.scalers.SCALERS           <- node_classification/scalers.py
All aggregation runs in libmma_b200.so (include/mma_b200.h); there is no CPU fallback.
"""
from . import _lib
from .graph import Graph, cached_graph, clear_cache
from .functional import mmconv_aggregate, segment_sum_rows, dropout_keep_scale, scale_table
from .linear import Linear
from .graph_regression.mma_conv import MMAConv
from .graph_regression.mask_aggr import MaskAggregateLinear

__all__ = ["Graph", "cached_graph", "clear_cache", "mmconv_aggregate", "segment_sum_rows",
           "dropout_keep_scale", "scale_table", "Linear", "MMAConv", "MaskAggregateLinear"]
