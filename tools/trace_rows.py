"""clock64 timeline of the row-local chain kernel (CTA 0).  Builds a trace library on the fly:
  nvcc ... -DSSAC_TRACE -o super_sac_b200/libssac_b200_trace.so super_sac_b200/csrc/*.cu   (done by tools/gpu_trace.sh)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
lib = ctypes.CDLL(os.path.join(ROOT, "super_sac_b200", "libssac_b200_trace.so"))
dev = "cuda"
trace = torch.zeros(64, dtype=torch.int64, device=dev)
lib.ssac_debug_set_trace_rows.argtypes = [ctypes.c_void_p]
print("set", lib.ssac_debug_set_trace_rows(trace.data_ptr()))
S, A, H, N, M, B = 17, 6, 256, 10, 2, 256
g = torch.Generator(device=dev).manual_seed(0)
def P(*s): return torch.randn(*s, device=dev, generator=g) * 0.05
aW = [P(1, H, S), P(1, H), P(1, H, H), P(1, H), P(1, 2 * A, H), P(1, 2 * A)]
cW = [P(N, H, S + A), P(N, H), P(N, H, H), P(N, H), P(N, 1, H), P(N, 1)]
X = P(B, S + A); eps = P(B, A); logp = torch.empty(B, device=dev); qt = torch.empty(M, B, 1, device=dev)
ni = torch.tensor([3, 7], dtype=torch.int32, device=dev)
f = lib.ssac_target_chain
f.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_float] * 4 + [ctypes.c_void_p] * 3
for it in range(3):
    rc = f(*[w.data_ptr() for w in aW], S, H, A, 0, *[w.data_ptr() for w in cW], ni.data_ptr(), M, X.data_ptr(), S + A, B, eps.data_ptr(), None, 0.0, 0.0, -5.0, 2.0, logp.data_ptr(), qt.data_ptr(), None)
    torch.cuda.synchronize()
t = trace.cpu().tolist()
print("rc", rc)
names = ["net start", "small operands in smem", "layer 1 done", "h1 broadcast issued", "cluster barrier 1", "layer 2 partials", "layer 2 reduced", "layer 3 + partial exchange issued", "cluster barrier 2", "outputs summed"]
print("entry -> after pdl:", t[1] - t[0], " total:", t[2] - t[0])
for base, nm in ((10, "actor"), (20, "critic 0"), (30, "critic 1")):
    for i in range(10):
        if t[base + i]:
            print(f"{nm:9s} {names[i]:36s} +{t[base+i]-t[0]:7d}  (d {t[base+i]-(t[base+i-1] if i else t[base]):6d})")
