"""H2D copy bandwidth from pinned host memory: default pinned vs write-combined, one stream vs two concurrent halves.
Run on the GPU box; prints one line per case.  (Decides how bench.py / callers should allocate their ingest buffers.)"""
import ctypes as C
import time

rt = C.CDLL("libcudart.so")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]


def main():
    assert rt.cudaSetDevice(0) == 0
    s1, s2 = C.c_void_p(), C.c_void_p()
    rt.cudaStreamCreate(C.byref(s1)); rt.cudaStreamCreate(C.byref(s2))
    for mb in (64, 256, 576):
        nbytes = mb << 20
        d = C.c_void_p(); assert rt.cudaMalloc(C.byref(d), nbytes) == 0
        for name, flag in (("pinned", 0), ("write-combined", 4)):
            h = C.c_void_p(); assert rt.cudaHostAlloc(C.byref(h), nbytes, flag) == 0
            C.memset(h, 1, nbytes)
            for streams in (1, 2):
                best = 1e9
                for rep in range(6):
                    t0 = time.perf_counter()
                    if streams == 1:
                        rt.cudaMemcpyAsync(d, h, nbytes, 1, s1)
                    else:
                        half = nbytes // 2
                        rt.cudaMemcpyAsync(d, h, half, 1, s1)
                        rt.cudaMemcpyAsync(C.c_void_p(d.value + half), C.c_void_p(h.value + half), half, 1, s2)
                    rt.cudaStreamSynchronize(s1); rt.cudaStreamSynchronize(s2)
                    best = min(best, time.perf_counter() - t0)
                print(f"h2d {mb:4d} MB {name:15s} streams={streams}: {nbytes / best / 1e9:6.1f} GB/s")
            rt.cudaFreeHost(h)
        rt.cudaFree(d)


if __name__ == "__main__":
    main()
