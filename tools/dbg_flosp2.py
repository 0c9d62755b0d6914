import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, torch.nn.functional as F
from oracle import functional as OF, synth
from oracle.functional import _bn, _conv2d
from test_gpu_net2d import _flosp_conf, _rel
from occdepth_b200.models.flosp_depth import FlospDepth
from occdepth_b200.engine import CL, Plan
import occdepth_b200.engine as E
torch.manual_seed(0)
H, W = 96, 320
conf = _flosp_conf(H, W)
m = synth.seed_weights_(FlospDepth(**conf), 5).eval()
K, Ts = synth.kitti_calib(W, H, focal=220.0)
cam_k = [torch.from_numpy(K).unsqueeze(0).repeat(2, 1, 1)]
T = [torch.stack([torch.from_numpy(t) for t in Ts])]
ida = [torch.eye(4).unsqueeze(0).repeat(2, 1, 1)]
feat = torch.randn(1, 2, 16, H // 8, W // 8)
sd = {"f." + k: v.clone() for k, v in m.state_dict().items()}
p = "f.depth_net.0"
x = feat.reshape(2, 16, H//8, W//8)
K4 = torch.zeros(2,4,4); K4[:, :3,:3] = cam_k[0].float(); K4[:,3,3]=1
inv = torch.inverse(K4)
sps = torch.norm(torch.stack([inv[..., 0, 0], inv[..., 1, 1]], dim=-1), dim=-1).reshape(-1, 1) * 1000.0
x1 = F.relu(_bn(sd, p + ".reduce_conv.1", _conv2d(sd, p + ".reduce_conv.0", x, 1, 1)))
h = F.linear(F.relu(F.linear(sps, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])), sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
g = torch.sigmoid(_conv2d(sd, p + ".se.conv_expand", F.relu(_conv2d(sd, p + ".se.conv_reduce", h[..., None, None]))))
x2 = x1 * g
# product: hook Plan.conv / add to capture outputs
m = m.cuda()
caps = {}
orig_conv = Plan.conv
def conv_hook(self, *a, **k):
    out = orig_conv(self, *a, **k)
    caps[k.get("name")] = out
    return out
Plan.conv = conv_hook
with torch.no_grad():
    got, got_d = m(feat.cuda(), cam_k, T, ida, None)
xr = caps["depthnet.reduce"].to_planar(True).cpu()
print("reduce*gate rel", _rel(xr, x2), " vs ungated", _rel(xr, x1))
print("gate oracle", g.flatten()[:5], "sps", sps.flatten())
plan = list(m._plans().values())[0][0]
for op in plan.ops:
    if getattr(op, "name", "") == "depthnet.se.gate":
        print("gate cuda", op._keep[3].flatten()[:5].cpu())
    if getattr(op, "name", "") == "depthnet.mlp.fc2":
        print("fc2 cuda", op._keep[3].flatten()[:5].cpu(), "oracle", h.flatten()[:5])
    if getattr(op, "name", "") == "depthnet.mlp.fc1":
        print("fc1 in (sps)", op._keep[0].flatten().cpu())
