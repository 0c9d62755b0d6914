"""Common forward plumbing of the drop-in modules: planar fp32 in/out, planned CUDA launches inside."""
import torch
import torch.nn as nn

from ..engine import CL, Plan, require_cuda


class B200Module(nn.Module):
    """nn.Module whose forward replays a cached `Plan` built by `emit(plan, x: CL) -> CL | dict`.

    Parameters/buffers are ordinary nn.Module state (so `state_dict` matches the reference); the plan snapshots
    them (BN folded, bf16 packed) on first use and is dropped whenever a state_dict is loaded.
    Forward-only: eval() mode on a CUDA device; anything else raises (there is no CPU fallback).
    """

    def _plans(self):
        d = self.__dict__.get("_plan_cache")
        if d is None:
            d = {}
            self.__dict__["_plan_cache"] = d
        return d

    def invalidate_plans(self):
        for m in self.modules():
            if isinstance(m, B200Module):
                m.__dict__["_plan_cache"] = {}

    def _load_from_state_dict(self, *a, **k):
        self.__dict__["_plan_cache"] = {}
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.__dict__["_plan_cache"] = {}
        return super()._apply(fn, *a, **k)

    def _check_mode(self, x):
        require_cuda(x, type(self).__name__ + ".forward")
        if self.training:
            raise RuntimeError("%s: occdepth_b200 implements the forward/inference path only; call .eval() "
                               "(BatchNorm is folded from running statistics)" % type(self).__name__)

    def _run_planar(self, x, squeeze_d=False):
        """x: planar fp32 [B,C,D,H,W] / [B,C,H,W]; returns emit()'s CL output(s) converted back to planar fp32."""
        self._check_mode(x)
        key = (tuple(x.shape), str(x.device))
        ent = self._plans().get(key)
        if ent is None:
            plan = Plan(x.device)
            if x.dim() == 4:
                B, C_, H, W = x.shape
                D = 1
            else:
                B, C_, D, H, W = x.shape
            xin = plan.alloc(B, D, H, W, C_)
            with torch.no_grad():
                y = self.emit(plan, xin)
            ent = (plan, xin, y)
            self._plans()[key] = ent
        plan, xin, y = ent
        CL.from_planar(x, out=xin)
        plan.run()
        return _to_planar(y, squeeze_d)


def _to_planar(y, squeeze_d):
    if isinstance(y, CL):
        return y.to_planar(squeeze_d)
    if isinstance(y, dict):
        return {k: _to_planar(v, squeeze_d) for k, v in y.items()}
    if isinstance(y, (tuple, list)):
        return type(y)(_to_planar(v, squeeze_d) for v in y)
    return y  # already a torch tensor (fp32 planar written by a kernel)
