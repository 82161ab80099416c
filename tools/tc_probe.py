"""GPU probe: tcgen05 MLP engine vs the fp32 engine on the same points (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import helpers
from neumesh_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = synth.ModelConfig()
mesh = synth.icosphere_mesh(4, seed=0)
sd = synth.make_state_dict(mesh, cfg, seed=1)
dev = torch.device("cuda:0")
x, v = helpers.sample_points(n, seed=21)
x, v = x.to(dev), v.to(dev)
m32 = helpers.cuda_model(mesh, cfg, sd, "fp32")
with torch.no_grad():
    s32 = m32.forward_density_only(x)
    s32n, n32 = m32.forward_with_nablas(x)
    _, c32 = m32.forward(x, v)
torch.cuda.synchronize()
print("fp32 ok", s32[:4, 0].tolist(), flush=True)
mtc = helpers.cuda_model(mesh, cfg, sd, "tcgen05")
with torch.no_grad():
    t = time.time()
    stc = mtc.forward_density_only(x)
    torch.cuda.synchronize()
    print("tc sdf done in %.3fs" % (time.time() - t), stc[:4, 0].tolist(), "max|diff|", (stc - s32).abs().max().item(), flush=True)
    stcn, ntc = mtc.forward_with_nablas(x)
    torch.cuda.synchronize()
    print("tc jvp: sdf diff", (stcn - s32).abs().max().item(), "nabla diff", (ntc - n32).abs().max().item(), flush=True)
    _, ctc = mtc.forward(x, v)
    torch.cuda.synchronize()
    print("tc color diff", (ctc - c32).abs().max().item(), flush=True)
