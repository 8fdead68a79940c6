"""raymarchcl_b200 -- B200-native voxel ray-march render op (drop-in for thi-ng/raymarchcl's
RenderImage/TonemapImage OpenCL path). Host-side mirror of the reference's Clojure interface over
the C-ABI library ``libraymarch_b200.so`` (include/raymarch_b200.h)."""
from .options import (OPTS_BYTES, TABLE_ENTRIES, PRESETS, render_options, encode_render_opts,
                      decode_render_opts, make_render_option_buffers, compute_eyepos)
from .generators import generate_scatter_offsets, make_gyroid_volume, make_terrain, make_blob_volume
from .volio import save_volume, load_volume
from .meshvoxel import load_mesh, mesh_scale, voxelize, voxelize_ks

__all__ = [
    "OPTS_BYTES", "TABLE_ENTRIES", "PRESETS", "render_options", "encode_render_opts",
    "decode_render_opts", "make_render_option_buffers", "compute_eyepos",
    "generate_scatter_offsets", "make_gyroid_volume", "make_terrain", "make_blob_volume",
    "save_volume", "load_volume", "load_mesh", "mesh_scale", "voxelize", "voxelize_ks",
]
