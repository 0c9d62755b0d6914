import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from oracle import functional as OF, synth
from test_gpu_net2d import _flosp_conf, _rel
from occdepth_b200.models.flosp_depth import FlospDepth
torch.manual_seed(0)
H, W = 96, 320
conf = _flosp_conf(H, W)
m = synth.seed_weights_(FlospDepth(**conf), 5).eval()
K, Ts = synth.kitti_calib(W, H, focal=220.0)
cam_k = [torch.from_numpy(K).unsqueeze(0).repeat(2, 1, 1)]
T = [torch.stack([torch.from_numpy(t) for t in Ts])]
ida = [torch.eye(4).unsqueeze(0).repeat(2, 1, 1)]
feat = torch.randn(1, 2, 16, H // 8, W // 8)
with torch.no_grad():
    want, want_d = OF.flosp_depth({"f." + k: v.clone() for k, v in m.state_dict().items()}, "f", feat, cam_k, T, ida, conf)
    got, got_d = m.cuda()(feat.cuda(), cam_k, T, ida, None)
print("depth rel", _rel(got_d, want_d), "prior rel", _rel(got, want))
print("want_d stats", want_d.min().item(), want_d.max().item(), "got_d", got_d.min().item(), got_d.max().item())
print("want", want.abs().max().item(), (want>0).float().mean().item(), "got", got.abs().max().item(), (got>0).float().mean().item())
# sample the oracle depth with the CUDA sampler
import ctypes as C
from occdepth_b200 import _lib
cams, sps, vn = m.camera_tables(cam_k, T, ida, None, 2)
print("sps", sps.flatten(), "vn", vn)
L=_lib.lib()
d=want_d[0].contiguous().cuda(); out=torch.zeros(vn[0]*vn[1]*vn[2], device='cuda'); cc=cams[0].contiguous().cuda()
rc=L.occd_frustum_sample_fwd(d.data_ptr(), cc.data_ptr(), 2, d.shape[1], d.shape[2], d.shape[3], vn[0],vn[1],vn[2], float(W), float(H), 2.0, 18.0, 1, out.data_ptr(), 0, _lib.stream_ptr())
torch.cuda.synchronize()
print("sampler-only rel", _rel(out.view(1,1,*vn), want))
