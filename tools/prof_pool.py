"""Pooling epilogue for ncu / timing: python tools/prof_pool.py B T H dtype [full|random] [sweep]
`sweep` (experiments build only: PRS_LIB_PATH=.../libprs_x.so) walks cluster size x ring depth in one process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import persian_rag_system_b200 as P
B, T, H = (int(v) for v in sys.argv[1:4])
dt = getattr(torch, sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "random"
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(B + T)
hid = torch.randn(B, T, H, generator=g, device=dev).to(dt)
lens = torch.randint(1, T + 1, (B,), generator=g, device=dev) if mode == "random" else torch.full((B,), T, device=dev)
mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.int64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
me = mask.unsqueeze(-1).float()
want = torch.nn.functional.normalize((hid.float() * me).sum(1) / me.sum(1).clamp(min=1e-9), p=2, dim=1)
read = int(lens.sum().item()) * H * hid.element_size()


import ctypes
from persian_rag_system_b200 import _lib
from persian_rag_system_b200.flat import _torch_dtype_code
hid2 = hid.clone()                       # second copy: back-to-back calls never find their input in L2
out_buf = torch.empty(B, H, dtype=torch.float32, device=dev)


def raw(h):                              # the C-ABI call alone (no torch allocation between the events)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().prs_pool_norm(ctypes.c_void_p(h.data_ptr()), _torch_dtype_code(h), ctypes.c_void_p(mask.data_ptr()), B, T, H, 1,
                                        ctypes.c_void_p(out_buf.data_ptr()), 0, ctypes.c_void_p(st)))


def run(tag=""):
    for _ in range(3):
        out = P.mean_pool_normalize(hid, mask, True)
    err = float((out - want).abs().max())
    ts = []
    for i in range(20):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); raw(hid); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        raw(hid2 if i & 1 else hid)
    e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 20
    nb = B * T * H * hid.element_size()
    print(f"{tag}B={B} T={T} H={H} {sys.argv[4]} mask={mode}: single call median {ts[10]*1e3:.1f} us ({nb/ts[10]/1e6:.0f} GB/s all-token, {read/ts[10]/1e6:.0f} unmasked); "
          f"back to back {b2b*1e3:.1f} us ({nb/b2b/1e6:.0f} / {read/b2b/1e6:.0f} GB/s); {nb/1e6:.0f} MB, unmasked {read/1e6:.0f} MB, max |err| {err:.1e}", flush=True)


if len(sys.argv) > 6 and sys.argv[6] == "sweep":
    for s in (1, 2, 4):
        for nst in (2, 3, 4, 6):
            os.environ["PRS_POOL_S"], os.environ["PRS_POOL_NST"] = str(s), str(nst)
            run(f"S={s} NST={nst} ")
else:
    run()
