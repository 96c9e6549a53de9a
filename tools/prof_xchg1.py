"""merge_xchg_kernel on ONE GPU (an exchange group of one rank): python tools/prof_xchg1.py [rows] [B]
For ncu source-level profiles of the merge + exchange kernel without a second GPU."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import persian_rag_system_b200 as P
from persian_rag_system_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
d, k = 768, 10
dev = torch.device("cuda", 0)
L = _lib.lib()
idx = P.IndexFlatIP(d, storage="fp16"); idx.reserve(n)
g = torch.Generator(device=dev).manual_seed(1)
for o in range(0, n, 250_000):
    xb = torch.randn(min(250_000, n - o), d, generator=g, device=dev); xb /= xb.norm(dim=1, keepdim=True); idx.add(xb.half())
x = ctypes.c_void_p()
_lib.check(L.prs_xchg_create(0, 1, 0, 1024, 16, ctypes.byref(x)))
hb = int(L.prs_xchg_handle_bytes())
mine = (ctypes.c_ubyte * hb)()
_lib.check(L.prs_xchg_get_handle(x, mine))
_lib.check(L.prs_xchg_open_peers(x, bytes(mine)))
q = torch.randn(B, d, generator=g, device=dev)
D = torch.empty(B, k, dtype=torch.float32, device=dev); I = torch.empty(B, k, dtype=torch.int64, device=dev)
Dw, Iw = idx.search(q, k)
idx.set_timing(True)
st = torch.cuda.current_stream(dev).cuda_stream


def step():
    _lib.check(L.prs_index_search_sharded_device(idx._h, x, ctypes.c_void_p(q.data_ptr()), _lib.F32, B, k, ctypes.c_void_p(D.data_ptr()),
                                                 ctypes.c_void_p(I.data_ptr()), ctypes.c_void_p(st)))


for _ in range(5): step()
torch.cuda.synchronize()
assert torch.equal(I, Iw) and torch.equal(D, Dw)
idx.scan_time(); idx.phase_times()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): step()
e1.record(); torch.cuda.synchronize()
scan_ms, nl = idx.scan_time()
prep_ms, merge_ms = idx.phase_times()
print(f"G=1 exchange, {n} x {d}, B={B}: step {e0.elapsed_time(e1)/50*1e3:.1f} us, scan {scan_ms/50*1e3:.1f} us, prep {prep_ms/50*1e3:.1f} us, merge+exchange {merge_ms/50*1e3:.1f} us")
