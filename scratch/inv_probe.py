# m31_inv throughput for the K knob (SSYM_M31_INV_K), 2^28 elements, device resident
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stark_symphony_b200 as S
ver = S.Verifier(0)
n = 1 << 28
a = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
out = torch.empty_like(a); fail = torch.empty(n, dtype=torch.uint8, device="cuda")
import ctypes as C
lib = S.load()
def run():
    assert 0 == (lib.ssym_m31_inv(ver.h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(fail.data_ptr()), n, 0))
for _ in range(3): run()
ver.synchronize()
t0 = time.perf_counter()
for _ in range(10): run()
ver.synchronize()
dt = (time.perf_counter() - t0) / 10
print(os.environ.get("SSYM_M31_INV_K"), f"{n/dt/1e9:.1f} G elem/s {n*9/dt/1e9:.0f} GB/s")
