import sys, time, ctypes as C
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes, luts, _lib
H = codes.regular_random(8000, 3, 6, seed=20181001)
T, imax, B = 16, 50, 65536
tb = luts.minsum_like_tables(T, 6, 3, imax)
dec = pkg.Discrete_LDPC_Decoder_class(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
dec.init_OpenCL_decoding(B); dec.early_termination = False; dec.host_output_dtype = np.uint8
rng = np.random.default_rng(0)
host_in = pkg.pinned_empty((8000, B), np.uint8); host_in[:] = rng.integers(0, 16, size=(8000, B), dtype=np.uint8)
h = dec._ensure_handle()
# raw PCIe
d = torch.empty((8000, B), dtype=torch.uint8, device='cuda'); hin = torch.from_numpy(host_in)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(hin, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s", 5 * host_in.nbytes / (time.perf_counter() - t0) / 1e9)
for chunk in (4096, 8192, 16384, 32768, 65536):
    _lib.check(_lib.lib().ibldpc_set_host_chunk(h, chunk))
    for _ in range(2): dec.decode_OpenCL(host_in, buffer_in=False, return_buffer=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): dec.decode_OpenCL(host_in, buffer_in=False, return_buffer=False)
    dt = (time.perf_counter() - t0) / 5
    print("chunk", chunk, "ms", dt * 1e3, "Gbit/s", 4000 * B / dt / 1e9)
