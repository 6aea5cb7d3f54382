import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "ergm_b200", "build", "libtrace.so")
import torch
L = ctypes.CDLL(so)
B, T, H, nh = 32, 256, 768, 12
M = B * T
qkv = torch.randn(M, 3 * H, device="cuda").bfloat16()
out = torch.zeros(M, H, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B, nh, T, device="cuda")
L.ergm_attn_fwd.restype = ctypes.c_int
args = [ctypes.c_void_p(qkv.data_ptr()), ctypes.c_int64(3 * H), ctypes.c_int(0), ctypes.c_void_p(qkv.data_ptr()), ctypes.c_int64(3 * H), ctypes.c_int(H),
        ctypes.c_void_p(qkv.data_ptr()), ctypes.c_int64(3 * H), ctypes.c_int(2 * H), ctypes.c_void_p(out.data_ptr()), ctypes.c_int64(H), None,
        ctypes.c_void_p(lse.data_ptr()), None, ctypes.c_int(B), ctypes.c_int(nh), ctypes.c_int(T), ctypes.c_int(T), ctypes.c_int(64),
        ctypes.c_int(1), ctypes.c_int(0), ctypes.c_float(0.0), ctypes.c_uint64(0), ctypes.c_uint64(0), None]
for _ in range(3):
    rc = L.ergm_attn_fwd(*args)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
L.ergm_debug_attn_trace(buf)
t0 = buf[0]
names = {0: "start", 1: "setup done", 40: "roles done", 41: "after final sync"}
for j in range(2):
    for k, n in enumerate(["wait_s", "got_s", "pass1 done", "xch done", "pass2 done", "arrived p", "got o", "o accumulated"]):
        names[2 + 8 * j + k] = "it%d %s" % (j, n)
for i in sorted(names):
    if buf[i]:
        print("%-22s %8d ns" % (names[i], buf[i] - t0))
