"""One config-2 forward for ncu (launch list + DRAM traffic per kernel): CUDA graph off, two warm-up forwards, then ONE
forward between cudaProfilerStart/Stop.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/r02_ncu_<prec>.csv python tools/ncu_forward.py <prec>"""
import os
import sys

os.environ["OCCDEPTH_CUDA_GRAPH"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
dev = torch.device("cuda", 0)
m = bench.build_model().to(dev).set_precision(prec)
img, pix, fov = bench.make_inputs(0)
b = {"img": img.to(dev), "projected_pix_2": [pix.to(dev)], "fov_mask_2": [fov.to(dev)]}
with torch.no_grad():
    for _ in range(2):
        m(b)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(b)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
