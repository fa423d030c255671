import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
lib = ctypes.CDLL(os.path.join(ROOT, "super_sac_b200", "libssac_b200_trace.so"))
import sys as _s
G, D, H, O, B = 10, int(_s.argv[1]) if len(_s.argv) > 1 else 256, 256, 1, 256
dev = "cuda"
trace = torch.zeros(64, dtype=torch.int64, device=dev)
lib.ssac_debug_set_trace.argtypes = [ctypes.c_void_p]
print("set_trace rc", lib.ssac_debug_set_trace(trace.data_ptr()))
W1 = torch.randn(G, H, D, device=dev); b1 = torch.randn(G, H, device=dev); W2 = torch.randn(G, H, H, device=dev); b2 = torch.randn(G, H, device=dev)
W3 = torch.randn(G, O, H, device=dev); b3 = torch.randn(G, O, device=dev)
x = torch.randn(B, D, device=dev); h1 = torch.empty(G, B, H, device=dev); h2 = torch.empty_like(h1); y = torch.empty(G, B, O, device=dev)
f = lib.ssac_mlp_forward
f.argtypes = [ctypes.c_void_p]*7 + [ctypes.c_int]*4 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int] + [ctypes.c_void_p]*2 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
lib.ssac_set_fused_forward.argtypes = [ctypes.c_int]
lib.ssac_set_fused_forward(0)   # this tool traces the layered GEMM kernel
for it in range(3):
    rc = f(W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), W3.data_ptr(), b3.data_ptr(), None, G, D, H, O, x.data_ptr(), D, 0, B, h1.data_ptr(), h2.data_ptr(), 1, y.data_ptr(), 2, None)
    torch.cuda.synchronize()
t = trace.cpu().tolist()
print("rc", rc)
base = t[0]
names = {0: "entry", 1: "setup done", 30: "loop done", 31: "mma done", 32: "tile in smem", 33: "stored", 34: "exit"}
for i in range(45):
    if t[i]:
        nm = names.get(i) or ["iter start", "raw landed", "arrived full"][(i - 2) % 3] + f" kc={(i-2)//3}"
        print(f"{i:3d} {nm:24s} {t[i]-base:8d} cycles")
