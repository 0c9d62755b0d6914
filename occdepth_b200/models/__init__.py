"""Drop-in counterparts of the reference's occdepth/models/*.py (forward path, CUDA only)."""
