import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
def run(name, H, B, imax=50):
    N = H.shape[1]; K = N - H.shape[0]
    q = pkg.AWGN_Channel_Quantizer(10 ** (-1.6 / 10), 3, 16, 2000)
    q.init_OpenCL_quanti(N, B, return_buffer_only=True)
    for cls, meth, algo in ((pkg.Min_Sum_Decoder_class_irregular, "decode_OpenCL_min_sum", "min-sum"),
                            (pkg.BeliefPropagationDecoderClassIrregular, "decode_OpenCL_belief_propagation", "BP")):
        for dt in (np.float32, np.float64):
            q.llr_dtype = dt
            llr = q.quantize_direct_OpenCL_LLR(N, B)
            dec = cls(H, imax, 16, B); dec.init_OpenCL_decoding(B); dec.early_termination = False
            for _ in range(2): out = getattr(dec, meth)(llr, buffer_in=True, return_buffer=True)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): out = getattr(dec, meth)(llr, buffer_in=True, return_buffer=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            E = H.nnz; sz = 4 if dt == np.float32 else 8
            bytes_frame = (imax - 1) * (4 * E + N) * sz
            print(f"{name} {algo} {dt.__name__}: {ms:.2f} ms/decode B={B} -> {K*B/ms/1e6:.3f} Gbit/s info, "
                  f"{bytes_frame*B/ms/1e6:.0f} GB/s algorithmic ({bytes_frame*B/ms/1e6/6459.9:.2f} of HBM peak), errors {dec.return_errors_all_zero(out)}")
            del dec, llr, out
            torch.cuda.empty_cache()
run("c1", codes.regular_random(8000, 3, 6), 4096)
run("wlan1296", codes.wlan_80211n(54), 32768)
