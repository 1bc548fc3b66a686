import sys, ctypes as C
sys.path.insert(0, "/root/repo/pytorch-a2c_b200")
import torch
from a2c_b200 import ops
from a2c_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda:0")
Nb = int(sys.argv[1]); mode = sys.argv[2]
Cin, Cout, H, K, S, P = 16, 24, 80, 3, 1, 1
x = torch.randn(Nb, Cin, H, H, device=dev); w = torch.randn(Cout, Cin, K, K, device=dev) * 0.1; b = torch.randn(Cout, device=dev)
s = ops.conv_shape(Cin, Cout, H, H, K, S, P)
y = torch.empty(Nb, Cout, s.OH, s.OW, device=dev); gy = torch.randn_like(y)
dW = torch.empty_like(w); db = torch.empty_like(b)
need = call("a2c_conv2d_bwd_weight_workspace", C.byref(s)); ws = ops.workspace(dev, max(need, 64 << 20))
for _ in range(2):
    if mode == "fwd":
        call("a2c_conv2d_fwd", ptr(x), None, ptr(w), ptr(b), ptr(y), Nb, C.byref(s), 1, 1, stream_ptr())
    else:
        call("a2c_conv2d_bwd_weight", ptr(x), None, ptr(gy), None, ptr(dW), ptr(db), Nb, C.byref(s), ptr(ws), ws.numel(), 1, stream_ptr())
torch.cuda.synchronize()
