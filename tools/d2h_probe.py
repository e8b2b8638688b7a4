"""Developer probe: device-to-host copy rate into ordinary pinned memory vs write-combined pinned memory."""
import ctypes as C, time, torch
rt = C.CDLL("libcudart.so.12")
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, device="cuda")
src.fill_(3)
stream = torch.cuda.current_stream().cuda_stream
for name, flag in (("default", 0), ("portable", 1), ("write-combined", 4)):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flag)) == 0
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(4):
            assert rt.cudaMemcpyAsync(p, C.c_void_p(src.data_ptr()), C.c_size_t(n), C.c_int(2), C.c_void_p(stream)) == 0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"{name:15s} {4 * n / dt / 1e9:6.1f} GB/s")
    rt.cudaFreeHost(p)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): h.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"{'torch pinned':15s} {4 * n / dt / 1e9:6.1f} GB/s")
# many smaller copies, as the pipeline issues them (28 per chunk, ~40 MB each)
m = 40 << 20
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(25): h[i * m:(i + 1) * m].copy_(src[i * m:(i + 1) * m], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"{'25 x 40 MB':15s} {25 * m / dt / 1e9:6.1f} GB/s")
