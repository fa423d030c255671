"""clock64 timeline of the single-kernel forward (CTA 0 of cluster 0).  Build the trace library first:
  nvcc <flags of __graft_entry__> -DSSAC_TRACE -o super_sac_b200/libssac_b200_trace.so super_sac_b200/csrc/*.cu
Usage: python tools/trace_fused.py [G D H O B keep]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
lib = ctypes.CDLL(os.path.join(ROOT, "super_sac_b200", "libssac_b200_trace.so"))
a = [int(v) for v in sys.argv[1:]]
G, D, H, O, B, keep = (a + [10, 23, 256, 1, 256, 0][len(a):])[:6]
dev = "cuda"
trace = torch.zeros(256, dtype=torch.int64, device=dev)
lib.ssac_debug_set_trace_fz.argtypes = [ctypes.c_void_p]
print("set_trace rc", lib.ssac_debug_set_trace_fz(trace.data_ptr()))
W1 = torch.randn(G, H, D, device=dev); b1 = torch.randn(G, H, device=dev); W2 = torch.randn(G, H, H, device=dev); b2 = torch.randn(G, H, device=dev)
W3 = torch.randn(G, O, H, device=dev); b3 = torch.randn(G, O, device=dev)
x = torch.randn(B, D, device=dev); h1 = torch.empty(G, B, H, device=dev); h2 = torch.empty_like(h1); y = torch.empty(G, B, O, device=dev)
f = lib.ssac_mlp_forward
f.argtypes = [ctypes.c_void_p]*7 + [ctypes.c_int]*4 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int] + [ctypes.c_void_p]*2 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
for it in range(3):
    rc = f(W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), W3.data_ptr(), b3.data_ptr(), None, G, D, H, O, x.data_ptr(), D, 0, B, h1.data_ptr(), h2.data_ptr(), keep, y.data_ptr(), 2, None)
    torch.cuda.synchronize()
t = trace.cpu().tolist()
print("rc", rc, "shape", (G, D, H, O, B), "keep", keep)
names = {0: "entry", 1: "setup done", 2: "L1 operands staged", 3: "L1 accumulator ready", 30: "L2 accumulator ready", 31: "L3 partials done",
         32: "h2 stored / workers done", 33: "cluster barrier passed", 34: "head epilogue done", 35: "exit"}
names[36] = "L3 tmem loads issued"
names.update({37: "L3 grp0 relu", 38: "L3 grp1 relu", 39: "L3 grp2 relu", 40: "L3 grp3 relu", 41: "L3 fma done"})
for i in sorted(range(45), key=lambda i: t[i]):
    if t[i]:
        nm = names.get(i) or ["tmem->regs done", "W2 chunk landed", "arrived full"][(i - 4) % 3] + f" q={(i-4)//3}"
        r1 = t[128 + i] - t[128] if t[128 + i] else -1
        print(f"{i:3d} {nm:28s} rank0 {t[i]-t[0]:8d}   rank1 {r1:8d} cycles")
for r in range(2):
    b0 = t[64]
    print(f"rank {r} (ns, common clock): entry {t[64+8*r]-b0}  before cluster barrier {t[65+8*r]-b0}  after {t[66+8*r]-b0}  exit {t[67+8*r]-b0}")
