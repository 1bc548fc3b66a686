import sys, time
sys.path.insert(0, "/root/repo/pytorch-a2c_b200")
import torch
from a2c_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    M, N, K = 5120, 2000, 25600
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); dY = torch.randn(M, N, device=dev)
    for prec in (0, 1, 2):
        t1 = timeit(lambda: ops.gemm(X, K, 1, W, 1, K, M, N, K, precision=prec))
        t2 = timeit(lambda: ops.gemm(dY, N, 1, W, K, 1, M, K, N, precision=prec))
        t3 = timeit(lambda: ops.gemm(dY, 1, N, X, K, 1, N, K, M, precision=prec))
        fl = 2.0 * M * N * K / 1e9
        print("gemm prec %d: fwd %.2f ms (%.0f TF)  dgrad %.2f ms (%.0f TF)  wgrad %.2f ms (%.0f TF)" % (prec, t1, fl / t1, t2, fl / t2, t3, fl / t3))
if which in ("all", "conv"):
    Nb = 5120
    for (Cin, Cout, H, K, S, P) in [(2, 16, 80, 3, 1, 1), (16, 24, 80, 3, 1, 1), (24, 32, 80, 3, 2, 1), (32, 64, 40, 3, 2, 1)]:
        x = torch.randn(Nb, Cin, H, H, device=dev)
        w = torch.randn(Cout, Cin, K, K, device=dev) * 0.1
        b = torch.randn(Cout, device=dev)
        for prec in (0, 1, 2):
            ops.set_precision(prec)
            xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
            y = ops.Conv2dFn.apply(xr, wr, br, S, P, True)
            gy = torch.randn_like(y)
            s = ops.conv_shape(Cin, Cout, H, H, K, S, P)
            import ctypes as C
            from a2c_b200._lib import call, ptr, stream_ptr
            dym = ops.relu_bwd(gy, y)
            tf = timeit(lambda: call("a2c_conv2d_fwd", ptr(x), None, ptr(w), ptr(b), ptr(y), Nb, C.byref(s), 1, prec, stream_ptr()))
            dx = torch.empty_like(x); dW = torch.empty_like(w); db = torch.empty_like(b)
            td = timeit(lambda: call("a2c_conv2d_bwd_data", ptr(dym), None, ptr(w), ptr(dx), Nb, C.byref(s), prec, stream_ptr()))
            need = call("a2c_conv2d_bwd_weight_workspace", C.byref(s)); ws = ops.workspace(dev, max(need, 64 << 20))
            tw = timeit(lambda: call("a2c_conv2d_bwd_weight", ptr(x), None, ptr(dym), None, ptr(dW), ptr(db), Nb, C.byref(s), ptr(ws), ws.numel(), prec, stream_ptr()))
            fl = 2.0 * Nb * s.OH * s.OW * Cout * Cin * K * K / 1e9
            print("conv %d>%d %dx%d s%d prec %d: fwd %.2f ms (%.1f TF) dgrad %.2f ms (%.1f TF) wgrad %.2f ms (%.1f TF)" % (Cin, Cout, H, H, S, prec, tf, fl / tf, td, fl / td, tw, fl / tw))
        del x, y, gy, dym, dx
