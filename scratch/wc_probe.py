# H2D copy rate from default pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), 64 MiB copies
import ctypes as C, time, glob, os
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(cands[0])
n = 64 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, flags in (("default", 0), ("write-combined", 4)):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    for _ in range(3):
        rt.cudaMemcpy(C.c_void_p(dev.data_ptr()), p, C.c_size_t(n), C.c_int(1))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(0))
    rt.cudaDeviceSynchronize(); dt = time.perf_counter() - t0
    print(name, f"{20 * n / dt / 1e9:.2f} GB/s")
    rt.cudaFreeHost(p)
