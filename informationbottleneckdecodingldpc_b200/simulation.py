"""Monte-Carlo BER loop of the reference drivers (``while errors < min_errors`` over batches of
``msg_at_time`` all-zero codewords, Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:81-136),
with frames sharded over the ranks of a torchrun job and the error counters all-reduced per batch.

The loop is pipelined: batch k+1 is sampled and decoded while the counters of batch k are still in
flight; the stop decision uses counters that are one batch old on every rank alike, so all ranks leave
the loop in the same iteration (a few extra frames are decoded, never fewer than min_errors asks for).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .engine import count_errors_async


def ber_point(decoder, quantizer, msg_at_time: int, min_errors: int = 7000, max_frames: int = 10 ** 9,
              llr: bool = False, count_all_rows: bool | None = None):
    """Simulate one Eb/N0 point.  ``decoder`` / ``quantizer`` are initialised objects of this package
    (``init_OpenCL_decoding`` / ``init_OpenCL_quanti(..., return_buffer_only=True)`` done).
    Returns dict(bit_errors, frame_errors, frames, ber, fer, seconds, info_bit_rate)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    N = decoder.N_v
    irregular = hasattr(decoder, "data_len") and getattr(decoder, "_irregular", True)
    rows = N if (count_all_rows or (count_all_rows is None and not irregular)) else int(decoder.data_len)
    thr = int(getattr(decoder, "cardinality_T_decoder_ops", 2) / 2)
    totals = torch.zeros(4, dtype=torch.int64, device="cuda")
    base = torch.tensor([0, 0, msg_at_time, 0], dtype=torch.int64, device="cuda")
    if world > 1:
        # every rank draws its own channel realisations: sub-stream = rank unless the caller chose one
        if getattr(quantizer, "_stream", 0) is None:
            quantizer.set_stream(dist.get_rank())
        from .parallel import counter_allreduce_fn
        allreduce = counter_allreduce_fn(decoder)
    pending = None          # (event, pinned host copy) of the previous batch's running totals
    host = torch.zeros(4, dtype=torch.int64).pin_memory()
    side = torch.cuda.Stream()
    t0 = time.time()
    known = np.zeros(4, dtype=np.int64)
    while known[0] < min_errors and known[2] < max_frames:
        if llr:
            rec = quantizer.quantize_direct_OpenCL_LLR(N, msg_at_time)
            out = decoder.decode(rec, buffer_in=True, return_buffer=True)
        else:
            rec = quantizer.quantize_direct_OpenCL(N, msg_at_time)
            out = decoder.decode_OpenCL(rec, buffer_in=True, return_buffer=True)
        c = base.clone()                         # {0, 0, msg_at_time, 0}
        count_errors_async(out, rows, thr, c)    # library kernel for cluster and LLR buffers alike
        if world > 1:
            allreduce(c)
        totals.add_(c)
        # read back the totals of the previous batch (already complete) without stalling this one
        if pending is not None:
            pending.synchronize()
            known = host.numpy().copy()
        ev = torch.cuda.Event()
        snap = totals.clone()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host.copy_(snap, non_blocking=True)
            ev.record(side)
        pending = ev
    torch.cuda.synchronize()
    tot = totals.cpu().numpy()
    dt = time.time() - t0
    frames = int(tot[2])
    return dict(bit_errors=int(tot[0]), frame_errors=int(tot[1]), frames=frames,
                ber=float(tot[0]) / max(frames * rows, 1), fer=float(tot[1]) / max(frames, 1), seconds=dt,
                info_bit_rate=float(getattr(decoder, "R_c", (N - decoder.N_c) / N)) * frames * N / dt)
