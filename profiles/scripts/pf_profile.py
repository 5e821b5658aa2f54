#!/usr/bin/env python3
"""One per-frame-early-termination decode of the C1 workload (for an ncu launch list / timing breakdown)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import informationbottleneckdecodingldpc_b200 as pkg

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
mode = sys.argv[2] if len(sys.argv) > 2 else "frame"
wl = bench.workload(name)
B = int(sys.argv[3]) if len(sys.argv) > 3 else wl["B"]
t, tb, quanti, decodi = bench.build_ib(pkg, wl, B, 0)
ch = quanti.quantize_direct_OpenCL(t.n_var, B)
decodi.early_termination = {"frame": "frame", "batch": True, "fixed": False}[mode]
for _ in range(3):
    decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
e1.record()
torch.cuda.synchronize()
print(name, mode, "B", B, "ms", e0.elapsed_time(e1), "launches", decodi.info()[1], "i_num", decodi.last_i_num)
if mode != "frame":
    import ctypes as C
    from informationbottleneckdecodingldpc_b200 import _lib
    h = decodi._ensure_handle()
    L = _lib.lib()
    _lib.check(L.ibldpc_set_profiling(h, 1))
    decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
    ms3 = (C.c_float * 3)()
    n3 = (C.c_int32 * 3)()
    _lib.check(L.ibldpc_phase_times(h, ms3, n3))
    print("phase avg ms: cn", ms3[0] / max(n3[0], 1), "vn", ms3[1] / max(n3[1], 1), "other total", ms3[2])
if mode == "frame":
    inum = decodi.last_i_num_per_frame.tensor.float()
    print("mean i_num", float(inum.mean()), "hist", torch.bincount(decodi.last_i_num_per_frame.tensor.long()).tolist())
