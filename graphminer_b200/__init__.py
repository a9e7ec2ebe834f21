"""graphminer_b200 -- B200-native set-intersection engine for graph pattern mining.

The product is libgminer_b200.so (hand-written CUDA for sm_100a behind a C ABI, include/gminer_b200.h).
This package is the thin Python mirror: `capi` (ctypes binding) and `rmat` (synthetic inputs).
"""
from . import capi  # noqa: F401
from .capi import (DeviceGraph, GMError, device_count, host_edgelist, host_orient,  # noqa: F401
                   host_partition_part, host_shard_bounds, intersect_batch, kclique_host, motif_host,
                   read_graph, set_option, sgl_host, tc_host, write_graph)

__all__ = ["capi", "DeviceGraph", "GMError", "device_count", "host_orient", "host_edgelist",
           "host_partition_part", "host_shard_bounds", "intersect_batch", "tc_host", "kclique_host",
           "sgl_host", "motif_host", "read_graph", "write_graph", "set_option"]
