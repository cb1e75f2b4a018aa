"""svgir_b200 -- B200-native splatting + shading hot path of SVG-IR (host side).

Sub-modules: `_lib` (ctypes binding of the C ABI), `raster` (surfel rasteriser forward/backward),
`scene` (seeded synthetic clouds / cameras). The drop-in packages `svgss_rasterization` and
`rgss_rasterization` that sit next to this package re-export the reference's operator API.
"""
__version__ = "0.1.0"
