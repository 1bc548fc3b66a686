import sys, os
sys.path.insert(0, "/root/repo/pytorch-a2c_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, numpy as np
import helpers as H
from a2c_b200 import models, updater, ops
dev = torch.device("cuda:0")
g = H.load("c1_fc.npz"); hyps = H.hyps_of(g)
R, T, fs, h = 4, 15, 2, 256
N = R*T
net = models.FCModel([fs,80,80],3,h_size=h); net.load_state_dict(H.params_for("FCModel", fs, h)); net.to(dev)
upd = updater.Updater(net, hyps)
shared = {k: torch.from_numpy(g["u0_"+k]).to(dev) for k in ("rewards","dones","deltas","actions")}
shared["states"] = torch.from_numpy(H.golden_states(g,0,N,fs)).float().to(dev)
res = upd.compute_gradients(shared)
tot = 0.0
for name, p in net.named_parameters():
    gr = p.grad
    off = upd.flat.offsets[[id(q) for q in upd.flat.params].index(id(p))]
    fv = upd.flat.grad[off:off+p.numel()].view(p.shape)
    same = gr.data_ptr() == fv.data_ptr()
    nrm = float((gr.double()**2).sum())
    tot += nrm
    key = "u0_grad/"+name
    print(name, "view_same", same, "maxdiff_flat", float((gr-fv).abs().max()), "sumsq", nrm, "err_vs_golden", H.compact_err(gr.cpu().numpy(), g[key]) if key in g else None)
print("sqrt tot", tot**0.5, "flat norm", float(upd.flat.grad.double().norm()), "golden", float(g["u0_GradNorm"]))
ss = ops.grad_sumsq(upd.flat.grad); print("kernel sumsq sqrt", float(ss.sqrt()))
# golden per-param norms
tg = 0.0
for name in H.O.param_shapes("FCModel", fs, h):
    v = g["u0_grad/"+name]
    tg += float(v[0])**2 if v.size > 4096 else float((v.astype(np.float64)**2).sum())
print("golden per-param total", tg**0.5)
