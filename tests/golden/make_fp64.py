"""How accurate is the fp32 reference itself at full size?  Oracle in float64 vs golden (reference fp32)."""
import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); os.chdir(ROOT)
sys.path.insert(0, "pytorch-a2c_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import helpers as H
from oracle import a2c_oracle as O
tag, model, R, T = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
g = H.load("full_%s.npz" % tag); hyps = H.hyps_of(g)
fs, h = 2, 256
traj, shared = H.full_inputs(model, R, T, fs, h, int(g["seed"]))
torch.set_default_dtype(torch.float64)
params = {k: v.double() for k, v in H.params_for(model, fs, h).items()}
p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
states = shared["states_u8"].double()
advs = O.discount(shared["deltas"], shared["dones"], hyps["gamma"]*hyps["lambda_"]).double()
rets = O.discount(shared["rewards"], shared["dones"], hyps["gamma"]).double()
if model == "GRUModel":
    vals, logits = O.bptt(model, p, states, shared["h_states"].double(), shared["dones"].double(), R, T)
else:
    vals, logits = O.model_forward(model, p, states)
a = (advs - advs.mean())/(advs.std()+1e-6)
loss, pi_l, v_l, e_l = O.a2c_loss(logits, vals, shared["actions"], a, rets, hyps)
loss.backward()
print("loss fp64", loss.item(), "golden", float(g["Loss"]))
tot = 0.0
out = {}
for k, q in p.items():
    if q.grad is None: continue
    gr = q.grad.numpy(); tot += (gr**2).sum()
    key = "grad/"+k
    scale = H.golden_grad_scale(g, "grad/")
    print("%-22s ref-vs-fp64 err %.3e" % (k, H.grad_err(gr.astype(np.float32), g[key], scale)))
    out["grad64/"+k] = H.compact(gr.astype(np.float32)) if gr.size > 4096 else gr.astype(np.float64).reshape(-1)
print("true norm", tot**0.5, "golden GradNorm", float(g["GradNorm"]))
out["norm64"] = np.float64(tot**0.5)
np.savez_compressed("tests/golden/fp64_%s.npz" % tag, **out)
