# qm31_inv / cm31_inv throughput, 2^26 elements, device resident
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import stark_symphony_b200 as S
ver = S.Verifier(0); lib = S.load()
n = 1 << 26
for name, w in (("qm31_inv", 4), ("cm31_inv", 2)):
    a = torch.randint(0, 2**31 - 1, (n * w,), dtype=torch.int32, device="cuda")
    out = torch.empty_like(a); fail = torch.empty(n, dtype=torch.uint8, device="cuda")
    fn = getattr(lib, "ssym_" + name)
    run = lambda: fn(ver.h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(fail.data_ptr()), n, 0)
    for _ in range(3): assert run() == 0
    ver.synchronize(); t0 = time.perf_counter()
    for _ in range(10): run()
    ver.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(name, f"{n/dt/1e9:.1f} G elem/s {n*(8*w+1)/dt/1e9:.0f} GB/s")
