import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes, luts
H = codes.regular_random(8000, 3, 6, seed=20181001)
T, imax = 16, 50
tb = luts.minsum_like_tables(T, 6, 3, imax)
rng = np.random.default_rng(0)
def mk(B):
    dec = pkg.Discrete_LDPC_Decoder_class(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    dec.init_OpenCL_decoding(B); dec.early_termination = False
    ch = torch.from_numpy(rng.integers(0, 16, size=(8000, B), dtype=np.uint8)).cuda()
    return dec, ch
for B, ns in ((16384, 1), (8192, 2), (4096, 4), (2048, 4), (2048, 8), (1024, 8), (1024, 16), (512, 16)):
    decs = [mk(B) for _ in range(ns)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    reps = max(2, 65536 // (B * ns))
    def run():
        for r in range(reps):
            for (dec, ch), st in zip(decs, streams):
                with torch.cuda.stream(st):
                    dec.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    frames = B * ns * reps
    print(f"B={B} streams={ns} reps={reps}: {frames/dt/1e3:.1f} kframes/s = {4000*frames/dt/1e9:.3f} Gbit/s")
    del decs
