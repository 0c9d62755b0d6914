from .flosp_depth import FlospDepth, flosp_depth_conf_map  # noqa: F401
