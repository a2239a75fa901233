"""H2D bandwidth of one frame's particles (12 MB) from plain pinned memory and from write-combined pinned memory,
alone and with a concurrent 8.3 MB D2H on a second stream.    python tools/wc_probe.py
"""
import ctypes as C
import glob
import os
import time

import torch

lib = None
for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*"):
    try:
        lib = C.CDLL(p)
        break
    except OSError:
        pass
assert lib is not None
torch.cuda.init()
torch.zeros(1, device="cuda")
N_IN, N_OUT = 11_959_344, 8_294_400
vp = C.c_void_p


def host(nbytes, flags):
    p = vp()
    assert lib.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    C.memset(p, 1, nbytes)
    return p


def dev(nbytes):
    p = vp()
    assert lib.cudaMalloc(C.byref(p), C.c_size_t(nbytes)) == 0
    return p


def stream():
    s = vp()
    assert lib.cudaStreamCreateWithFlags(C.byref(s), C.c_uint(1)) == 0
    return s


d_in, d_out = dev(N_IN), dev(N_OUT)
h_out = host(N_OUT, 0)
s1, s2 = stream(), stream()
for name, flags in (("pinned", 0), ("write-combined", 4), ("pinned", 0), ("write-combined", 4)):
    h_in = host(N_IN, flags)
    for both in (False, True):
        best = 1e9
        for rnd in range(5):
            lib.cudaDeviceSynchronize()
            t0 = time.perf_counter()
            for _ in range(40):
                lib.cudaMemcpyAsync(d_in, h_in, C.c_size_t(N_IN), C.c_int(1), s1)
                if both:
                    lib.cudaMemcpyAsync(h_out, d_out, C.c_size_t(N_OUT), C.c_int(2), s2)
            lib.cudaDeviceSynchronize()
            best = min(best, (time.perf_counter() - t0) / 40)
        print(f"{name:15s} {'H2D + D2H' if both else 'H2D only '}: {best * 1e3:.4f} ms per frame, H2D {N_IN / best / 1e9:.1f} GB/s", flush=True)
    lib.cudaFreeHost(h_in)
