def marching_cubes(*args, **kwargs):
    raise NotImplementedError("skimage shim: mesh extraction is outside the OccDepth forward hot path")


marching_cubes_lewiner = marching_cubes
