#!/usr/bin/env python3
"""Headline benchmark: decoded info Gbit/s of the IB decoder at fixed i_max (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own kernels on the host cores

Workload (BASELINE.json configs[0], the configuration the metric is quoted on): regular (3,6)
LDPC code n=8000, R=1/2, BPSK/AWGN, IB decoder |T|=16, i_max=50, early termination off,
B=65536 frames per GPU and step, channel cluster indices drawn by the inversion method from the
|T|=16 quantizer at Eb/N0 = 1.6 dB (all-zero codeword), exactly like quantize_direct_OpenCL;
IB tables designed at 1.2 dB by the in-repo discrete density evolution.  The decoder runs the
packed-nibble kernel family (four-bit messages); the roofline object reports both SURVEY 8(d)'s
algorithmic (uint8) bytes and the bytes really stored.
A "step" = decode one batch + count bit/frame errors (+ all-reduce of the 4 counters for N>1).
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded info Gbit/s @ fixed i_max"
N_VAR, D_V, D_C, T, IMAX = 8000, 3, 6, 16, 50
EBN0_DB = 1.2
SEED = 20181001


def workload(name):
    from informationbottleneckdecodingldpc_b200 import codes
    if name == "c1":
        return dict(name="(3,6) n=8000 R=0.5 IB |T|=16 i_max=50 ET off", H=codes.regular_random(8000, 3, 6, seed=SEED),
                    irregular=False, B=65536, ebn0=1.6, design_ebn0=1.2)
    if name == "wlan":
        return dict(name="802.11n n=1296 R=0.5 IB |T|=16 i_max=50 ET off, message alignment", H=codes.wlan_80211n(54),
                    irregular=True, B=100096, ebn0=2.0, design_ebn0=1.0)
    if name == "wlan1944":
        return dict(name="802.11n n=1944 R=0.5 IB |T|=16 i_max=50 ET off, message alignment", H=codes.wlan_80211n(81),
                    irregular=True, B=65536, ebn0=2.0, design_ebn0=1.0)
    if name == "dvbs2":
        return dict(name="DVB-S2-like n=64800 R=0.5 IB |T|=16 i_max=50 ET off, message alignment",
                    H=codes.dvbs2_like_half_rate(), irregular=True, B=8192, ebn0=1.6, design_ebn0=1.0)
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes_per_frame(N, E, imax):
    """uint8 messages, two-phase flooding, syndrome fused:
    (imax-1)(4E+N) + 2E + 3N bytes per frame."""
    return (imax - 1) * (4 * E + N) + 2 * E + 3 * N


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_from_phase_times(ms3, n3, N, E, B, imax, packed, workload, family=None):
    """The `roofline` object of the JSON line from the per-phase CUDA-event times of one profiled decode
    (ibldpc_phase_times: ms3 = [check-node launches, variable-node launches, iteration 0 + output + packing],
    n3 = launch counts).  Algorithmic bytes per launch as SURVEY.md 8(d) defines them (uint8 messages): CN 2E,
    VN 2E+N per frame; the packed-nibble kernels store two frames per byte, so the bytes they really move
    ("stored") are half of that -- both figures are reported, and `frac` (algorithmic / peak) may exceed 1 for that
    reason; `frac_stored` is the physical fraction of the HBM peak.  `bound` names the real limiter: with four-bit
    storage the kernels are NOT HBM-bound but sit on the issue slots / the shared-memory look-up pipe (ncu,
    profiles/README.md); the HBM peak stays the denominator SURVEY 8(d) prescribes.  When the whole decode ran as one
    cooperative launch (small batches) there are no per-phase times: the dominant "kernel" is then the whole decode
    with SURVEY's bytes per frame."""
    peak, peak_src = measured_peak_gbs()
    if family is None:
        family = 2 if packed else 1
    stored_div = 2 if packed else 1
    fam = {0: "generic", 1: "fast", 2: "n4", 3: "t32"}.get(family, "n4")
    bytes_frame = algorithmic_bytes_per_frame(N, E, imax)
    decode_ms = ms3[0] + ms3[1] + ms3[2]
    per_phase = n3[0] > 0 and n3[1] > 0 and ms3[0] > 0 and ms3[1] > 0
    cn_bytes, vn_bytes = 2 * E * B, (2 * E + N) * B
    if per_phase:
        # average duration of one check-node / variable-node PHASE (all degree classes of the phase; one launch when
        # the fused per-phase kernels run, one launch per class otherwise); n3 counts phases
        cn_ms, vn_ms = ms3[0] / n3[0], ms3[1] / n3[1]
        if ms3[0] >= ms3[1]:
            dom, dom_ms, dom_bytes = f"ib_cn_{fam} (check-node update + syndrome)", cn_ms, cn_bytes
        else:
            dom, dom_ms, dom_bytes = f"ib_vn_{fam} (variable-node update)", vn_ms, vn_bytes
    else:
        cn_ms = vn_ms = None
        dom, dom_ms, dom_bytes = "ib_decode_coop_kernel (whole decode in one cooperative launch)", decode_ms, bytes_frame * B
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_note = None, "profiles/traffic.json has no capture for this workload"
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tfile):
        try:
            tj = json.load(open(tfile)).get(workload)
            if tj and tj.get("kernel", "")[:5] == dom[:5]:
                if tj.get("source_hash") == source_hash():
                    # measured at tj["frames_per_launch"]; DRAM traffic of these kernels is linear in B
                    traffic = tj["bytes"] * B / tj["frames_per_launch"]
                    traffic_note = "ncu dram__bytes_read.sum + dram__bytes_write.sum, " + tj.get("source", "")
                else:
                    traffic_note = ("stale: captured on kernel sources %s, current sources are %s -- re-run "
                                    "profiles/capture_traffic.py" % (tj.get("source_hash"), source_hash()))
        except Exception as e:   # noqa: BLE001
            traffic_note = f"profiles/traffic.json unreadable: {e}"
    if family in (1, 2):
        if not packed:
            bound = "hbm / shared-memory look-up pipe"
        elif workload == "c1":
            bound = ("issue slots + shared-memory look-up pipe (ncu, profiles/r02_ncu_full_c1_final_summary.csv: check-node kernel issue 82 %, "
                     "l1tex data-pipe 82 %, DRAM 43 %; variable-node kernel issue 79 %, data-pipe 75 %, DRAM 63 % = 0.80 of the measured copy "
                     "bandwidth on the bytes really moved); HBM peak is the roofline denominator of SURVEY 8(d)")
        else:
            bound = ("shared-memory look-up pipe (ncu, profiles/r02_ncu_full_phase_{wlan,dvbs2}_summary.csv: l1tex data-pipe 82-89 %, issue 76-87 %, "
                     "DRAM 28-37 %); HBM peak is the roofline denominator of SURVEY 8(d)")
    elif family == 3:
        bound = "shared-memory look-up pipe (unstriped 32x32 byte tables, bank conflicts)"
    else:
        bound = "L1/L2 table gathers (generic path: tables in global memory)"
    return {
        "bound": bound, "denominator": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "frac_stored": achieved / stored_div / peak,
        "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
        "message_storage": "packed nibbles (2 frames per byte)" if packed else "uint8",
        "stored_bytes_per_launch": dom_bytes // stored_div, "achieved_stored": achieved / stored_div,
        "avg_launch_ms": dom_ms, "cn_avg_ms": cn_ms, "vn_avg_ms": vn_ms,
        "cn_frac": cn_bytes / (cn_ms * 1e-3) / 1e9 / peak if per_phase else None,
        "vn_frac": vn_bytes / (vn_ms * 1e-3) / 1e9 / peak if per_phase else None,
        "cn_share_of_decode": ms3[0] / decode_ms, "vn_share_of_decode": ms3[1] / decode_ms,
        "whole_decode": {"bytes_per_frame": bytes_frame, "achieved_gbs": bytes_frame * B / (decode_ms * 1e-3) / 1e9,
                         "frac": bytes_frame * B / (decode_ms * 1e-3) / 1e9 / peak,
                         "frac_stored": bytes_frame * B / (decode_ms * 1e-3) / 1e9 / peak / stored_div},
    }


class ClockSampler:
    """SM clock, power and throttle reasons sampled every 100 ms during the timed region, in-process through NVML
    (pynvml).  A spawned `nvidia-smi -lms` does the same job but its queries serialise with kernel launches in the
    driver and were seen to stretch launch-heavy steps (irregular codes: 1500 launches per step) by 10-25 %; it is
    only the fallback when NVML cannot be loaded."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = int(gpu_index)
        self.rows = []          # (sm_mhz, sm_max_mhz, power_w, reason names)
        self.proc = None
        self.nvml = None
        self.stop_flag = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid_order = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if uuid_order:
                ent = uuid_order.split(",")[self.gpu].strip()
                handle = pynvml.nvmlDeviceGetHandleByUUID(ent) if ent.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(ent))
            else:
                handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = (pynvml, handle)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv, h = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(smax), power, [n for bit, n in self.REASONS if mask & bit]))
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def _read_smi(self):
        names = [n for _, n in self.REASONS]
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((float(r[1]), float(r[2]), float(r[3]),
                                  [nm for nm, val in zip(names, r[4:8]) if val.lower() == "active"]))
            except Exception:
                continue

    def stop(self):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        if self.nvml is not None:
            self.stop_flag.wait(0.12)
            self.stop_flag.set()
            self.thread.join(timeout=2)
        else:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for n in r[3]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if sm else None,
                "power_w_max": max(r[2] for r in self.rows) if sm else None, "samples": len(sm), "reasons": reasons,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


_TABLES = {}   # (workload, T) -> designed tables: the legs of one run share them


def make_tables(wl, T_=T):
    """IB tables designed by discrete density evolution (decoder_config_generation.py, the in-repo
    stand-in for the reference's design chain); irregular codes with message alignment."""
    from informationbottleneckdecodingldpc_b200 import graph, luts
    key = (wl["name"], T_)
    if key in _TABLES:
        wl["tables"] = _TABLES[key][2]
        return _TABLES[key][0], _TABLES[key][1]
    t = graph.edge_tables(wl["H"])
    if not wl["irregular"]:
        from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_regular_config
        tb, _ = generate_regular_config(wl["design_ebn0"], t.d_v_max, t.d_c_max, T_, IMAX)
        wl["tables"] = "IB tables, discrete density evolution at Eb/N0 = %.1f dB (in-repo design)" % wl["design_ebn0"]
    else:
        from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
        tb, _ = generate_irregular_config(wl["design_ebn0"], wl["H"], T_, IMAX)
        wl["tables"] = ("IB tables + message alignment, degree-mixed density evolution at Eb/N0 = %.1f dB "
                        "(in-repo design)" % wl["design_ebn0"])
    _TABLES[key] = (t, tb, wl["tables"])
    return t, tb


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms are meant to use every host core."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)


def cpu_reference_run(wl, frames, seed=SEED, T_=T):
    """The reference's own kernels (oracle/_ref) -- or the C port when that library is absent -- on
    `frames` frames of the workload, all host threads.  Returns (seconds, kind, cores)."""
    from oracle import oracle
    from informationbottleneckdecodingldpc_b200 import AWGN_Channel_Quantizer
    t, tb = make_tables(wl, T_)
    R = 1 - t.n_chk / t.n_var
    q = AWGN_Channel_Quantizer(10 ** (-wl["ebn0"] / 10) / (2 * R), 3, T_, 2000)
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(size=(t.n_var, frames))
    ch = ((u[:, :, None] - q.cdf_t_given_x_equals_zero) > 0).sum(2) - 1
    kind = "reference" if oracle.ref_available() else "port"
    cores = os.cpu_count() or 1
    if kind == "reference":
        try:
            cores = int(oracle.ref_lib().ref_num_threads())
        except Exception:
            pass
    kw = dict(T=T_, imax=IMAX, cn_lut=tb.Trellis_checknodevector_a, vn_lut=tb.Trellis_varnodevector_a,
              cn_match=tb.matching_vector_checknode, vn_match=tb.matching_vector_varnode, early=False)
    t0 = time.perf_counter()
    try:
        oracle.ib_decode(t, ch, backend=kind, irregular=wl["irregular"], **kw)
    except KeyError:
        kind = "port"
        t0 = time.perf_counter()
        oracle.ib_decode(t, ch, backend="port", **kw)
    return time.perf_counter() - t0, kind, cores, t


def size_cpu_sample(wl, target_s, T_=T):
    """Calibrate on a small batch, then pick a frame count worth about `target_s` seconds."""
    cores = os.cpu_count() or 1
    cal = max(2, min(2 * cores, 64))
    dt, _, _, _ = cpu_reference_run(wl, cal, T_=T_)
    return int(max(cal, min(16384, cal * target_s / max(dt, 1e-3)))) // 1 or cal


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_threads()
    wl = workload(args.workload)
    frames = size_cpu_sample(wl, 8.0)
    times = []
    kind = cores = t = None
    for i in range(args.warmup + args.steps):
        dt, kind, cores, t = cpu_reference_run(wl, frames, seed=SEED + i)
        if i >= args.warmup:
            times.append(dt)
    K = t.n_var - t.n_chk
    total = sum(times)
    value = K * frames * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": wl["name"], "frames_per_step": frames, "i_max": IMAX, "early_termination": False,
                   "note": "reference OpenCL kernels compiled for the host CPU (oracle/_ref), OpenMP over all host threads"
                   if kind == "reference" else "C port of the reference kernels (oracle/ldpc_oracle.c), OpenMP"},
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": kind,
                         "sample": f"{frames} frames x {len(times)} steps of the workload, i_max={IMAX}, ET off"},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# translation units and headers the per-class packed-nibble kernels (the dominant kernels of the headline workload) are
# compiled from; profiles/traffic.json is tied to exactly these
TRAFFIC_SOURCES = ("ib_kernels.cuh", "ib_kernels_n4.cuh", "ib_triple_n4.cuh", "ib_n4_vn3.cu", "ib_n4_cn_pair.cu", "ib_n4_cn_v2.cu",
                   "ib_n4_vn_v2.cu", "ib_n4_vn_v4.cu", "ib_n4_vn_pair.cu", "kernel_tables.h")


def source_hash():
    """sha256 (16 hex digits) over the sources of the per-class packed-nibble kernels: ties profiles/traffic.json to the
    kernels it was captured on."""
    import hashlib
    from informationbottleneckdecodingldpc_b200 import _lib
    h = hashlib.sha256()
    for f in sorted(_lib.SOURCES):
        if os.path.basename(f) in TRAFFIC_SOURCES and os.path.exists(f):
            h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def sample_columns(B, n=32, seed=7):
    """Frame columns of a batch for the oracle parity sample: first, middle and last 8 + 8 random ones."""
    n = min(n, B)
    if B <= n:
        return np.arange(B)
    rng = np.random.Generator(np.random.PCG64(seed))
    q = n // 4
    cols = np.r_[0:q, B // 2 - q // 2:B // 2 - q // 2 + q, B - q:B, rng.integers(0, B, n - 3 * q)]
    return np.unique(np.clip(cols, 0, B - 1))


def parity_sample_ib(decodi, ch, out, t, tb, wl, T_):
    """Decoded columns of the TIMED batch geometry against the CPU oracle on the same inputs (bit-exact bar)."""
    from oracle import oracle
    import torch
    cols = sample_columns(ch.shape[1])
    idx = torch.from_numpy(cols).to(ch.tensor.device)
    ch_s = ch.tensor.index_select(1, idx).cpu().numpy()
    got = out.tensor.index_select(1, idx).cpu().numpy()
    t0 = time.perf_counter()
    ref, _ = oracle.ib_decode(t, ch_s, T=T_, imax=IMAX, cn_lut=tb.Trellis_checknodevector_a, vn_lut=tb.Trellis_varnodevector_a,
                              cn_match=tb.matching_vector_checknode, vn_match=tb.matching_vector_varnode, early=False)
    eq = bool(np.array_equal(got, ref.astype(np.uint8)))
    return {"frames": int(cols.size), "equal": eq, "checker": "oracle/ldpc_oracle.c (pinned to the reference kernels)",
            "mismatching_symbols": int((got != ref.astype(np.uint8)).sum()), "oracle_s": round(time.perf_counter() - t0, 2),
            "batch": int(ch.shape[1])}


def build_ib(pkg, wl, B, rank, T_=T):
    t, tb = make_tables(wl, T_)
    N, M = t.n_var, t.n_chk
    R = (N - M) / N
    quanti = pkg.AWGN_Channel_Quantizer(10 ** (-wl["ebn0"] / 10) / (2 * R), 3, T_, 2000)
    quanti.seed = SEED
    quanti.set_stream(rank)                      # independent Philox sub-stream per rank
    quanti.init_OpenCL_quanti(N, B, return_buffer_only=True)
    if wl["irregular"]:
        decodi = pkg.Discrete_LDPC_Decoder_class_irregular(wl["H"], IMAX, T_, T_, tb.Trellis_checknodevector_a,
                                                           tb.Trellis_varnodevector_a, tb.matching_vector_checknode,
                                                           tb.matching_vector_varnode, B)
    else:
        decodi = pkg.Discrete_LDPC_Decoder_class(wl["H"], IMAX, T_, T_, tb.Trellis_checknodevector_a,
                                                 tb.Trellis_varnodevector_a, B)
    decodi.init_OpenCL_decoding(B, quanti.context)
    decodi.early_termination = False
    return t, tb, quanti, decodi


def phase_roofline(decodi, ch, N, E, B, workload_name):
    """Per-launch CUDA-event times of one profiled decode -> the `roofline` object."""
    import ctypes as C
    from informationbottleneckdecodingldpc_b200 import _lib
    h = decodi._ensure_handle()
    L = _lib.lib()
    _lib.check(L.ibldpc_set_profiling(h, 1))
    decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
    ms3 = (C.c_float * 3)()
    n3 = (C.c_int32 * 3)()
    _lib.check(L.ibldpc_phase_times(h, ms3, n3))
    _lib.check(L.ibldpc_set_profiling(h, 0))
    fam = decodi.info()[0]
    return roofline_from_phase_times(list(ms3), list(n3), N, E, B, IMAX, fam == 2, workload_name, family=fam)


LAST_TIMING = {}   # filled by time_steps: repeats, clocks under load (the legs copy it into their result)


def _sm_clock_now(sampler):
    try:
        nv, h = sampler.nvml
        return float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
    except Exception:
        return None, None


def time_steps(step, steps, warmup, min_warm_s=0.5, max_warm_s=8.0, repeats=3):
    """`warmup` untimed steps -- at least `min_warm_s` seconds of them, until two consecutive steps agree within 2 % AND
    NVML reports the SM clock back at >= 97 % of its maximum (at most `max_warm_s`): a leg starts after seconds of host-side
    table design during which the idle GPU drops its clocks, and some boxes take seconds under load to come back -- then
    `repeats` timed regions of `steps` steps each (CUDA events on the current stream).  Returned: the FASTEST region (the
    others are kept in LAST_TIMING["ms_per_step_repeats"]; a region slowed by a clock ramp or a host hiccup is not the
    kernel's speed), with the clocks sampled over all timed regions."""
    import torch
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    t0 = time.perf_counter()
    n, prev, stable = 0, None, False
    while True:
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        step()
        w1.record()
        torch.cuda.synchronize()
        cur = w0.elapsed_time(w1)
        stable = prev is not None and abs(cur - prev) <= 0.02 * max(cur, prev)
        prev = cur
        n += 1
        el = time.perf_counter() - t0
        sm, smax = _sm_clock_now(sampler) if sampler.nvml is not None else (None, None)
        clock_ok = sm is None or smax is None or sm >= 0.97 * smax
        if (n >= max(warmup, 2) and el >= min_warm_s and stable and clock_ok) or el > max_warm_s:
            break
    sampler.rows.clear()
    torch.cuda.synchronize()
    regions = []
    last = None
    for _ in range(max(1, repeats)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            last = step()
        e1.record()
        torch.cuda.synchronize()
        regions.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    LAST_TIMING.clear()
    LAST_TIMING.update({"ms_per_step_repeats": [round(r / steps, 4) for r in regions], "warmup_steps": n,
                        "warmup_s": round(el, 2), "clocks": clocks})
    return min(regions), last


def leg_ib(pkg, name, steps, rank, T_=T, frames=0):
    """Short leg of another BASELINE configuration: value, roofline fractions and an oracle parity sample."""
    import torch
    wl = workload(name)
    B = frames or wl["B"]
    t, tb, quanti, decodi = build_ib(pkg, wl, B, rank, T_)
    N, M, E = t.n_var, t.n_chk, t.n_edge
    ch = quanti.quantize_direct_OpenCL(N, B)
    ms, out = time_steps(lambda: decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True), steps, 1)
    launches = decodi.info()[1]
    roof = phase_roofline(decodi, ch, N, E, B, name)
    par = parity_sample_ib(decodi, ch, out, t, tb, wl, T_)
    fam = decodi.info()[0]
    res = {"workload": wl["name"] if T_ == T else wl["name"].replace("|T|=16", f"|T|={T_}"), "frames_per_step": B, "steps": steps,
           "value": (N - M) * B * steps / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms / steps,
           "kernel_family": {0: "generic (tables in global memory)", 1: "uint8 shared-memory", 2: "packed-nibble (n4)",
                             3: "|T|<=32 shared-memory bytes (t32)"}.get(fam, str(fam)),
           "gpu_launches_per_step": launches, "tables": wl.get("tables"),
           "roofline": {k: roof[k] for k in ("bound", "kernel", "frac", "frac_stored", "avg_launch_ms", "cn_frac", "vn_frac")}
           | {"whole_decode_frac": roof["whole_decode"]["frac"], "whole_decode_frac_stored": roof["whole_decode"]["frac_stored"]},
           "parity_sample": par, "timing": dict(LAST_TIMING)}
    del decodi, ch, out
    torch.cuda.empty_cache()
    return res


def leg_reference_batches(pkg, rank, decodes=20):
    """The reference drivers' own batch sizes (msg_at_time = 100 / 2000 / 2: Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:51,
    Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL.py:77, .../DVB-S2/BER_simulation_OpenCL.py:71): time per decode call with
    device buffers, fixed i_max, and the WHOLE batch (at most 32 frames of it) against the oracle."""
    import torch
    res = {}
    for name, B, T_ in (("c1", 100, T), ("wlan", 2000, T), ("dvbs2", 2, T), ("wlan", 2000, 32)):   # 32: the reference's WLAN cardinality
        wl = workload(name)
        t, tb, quanti, decodi = build_ib(pkg, wl, B, rank, T_)
        ch = quanti.quantize_direct_OpenCL(t.n_var, B)
        ms, out = time_steps(lambda: decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True), decodes, 3, min_warm_s=0.2)
        res[f"{name}_B{B}" + ("" if T_ == T else f"_T{T_}")] = {"workload": wl["name"] if T_ == T else wl["name"].replace("|T|=16", f"|T|={T_}"), "frames_per_decode": B, "ms_per_decode": ms / decodes,
                               "value": (t.n_var - t.n_chk) * B * decodes / (ms * 1e-3) / 1e9, "unit": "Gbit/s",
                               "gpu_launches_per_decode": decodi.info()[1], "parity_sample": parity_sample_ib(decodi, ch, out, t, tb, wl, T_)}
        del decodi, ch, out
        torch.cuda.empty_cache()
    return res


def leg_llr(pkg, algo, steps, rank, B=16384):
    """min-sum / BP benchmark decoders on the (3,6) n=8000 graph in float64 (the reference's arithmetic; the only
    precision that meets the >= 99.99 % identical-frames bar, tests/test_gpu_parity.py)."""
    import torch
    from oracle import oracle
    from informationbottleneckdecodingldpc_b200 import graph
    wl = workload("c1")
    t = graph.edge_tables(wl["H"])
    N, M, E = t.n_var, t.n_chk, t.n_edge
    R = (N - M) / N
    quanti = pkg.AWGN_Channel_Quantizer(10 ** (-wl["ebn0"] / 10) / (2 * R), 3, T, 2000)
    quanti.seed = SEED
    quanti.set_stream(rank)
    quanti.llr_dtype = np.float64
    quanti.init_OpenCL_quanti(N, B, return_buffer_only=True)
    cls = pkg.Min_Sum_Decoder_class_irregular if algo == "minsum" else pkg.BeliefPropagationDecoderClassIrregular
    decodi = cls(wl["H"], IMAX, T, B)
    decodi.init_OpenCL_decoding(B, quanti.context)
    decodi.early_termination = False
    ch = quanti.quantize_direct_OpenCL_LLR(N, B)
    ms, out = time_steps(lambda: decodi.decode(ch, buffer_in=True, return_buffer=True), steps, 1)
    launches = decodi.info()[1]
    cols = sample_columns(B, 16)
    idx = torch.from_numpy(cols).to(ch.tensor.device)
    ch_s = ch.tensor.index_select(1, idx).cpu().numpy()
    got = out.tensor.index_select(1, idx).cpu().numpy()
    ref, _ = oracle.llr_decode(t, ch_s, algo=algo, imax=IMAX, early=False)
    same = ((got < 0) == (ref < 0)).all(axis=0)
    peak, _src = measured_peak_gbs()
    bytes_frame = (IMAX - 1) * (32 * E + 8 * N) + 8 * (2 * E + 3 * N)
    res = {"workload": f"(3,6) n=8000 {'min-sum' if algo == 'minsum' else 'belief propagation'} float64 i_max=50 ET off",
           "frames_per_step": B, "steps": steps, "value": (N - M) * B * steps / (ms * 1e-3) / 1e9, "unit": "Gbit/s",
           "ms_per_step": ms / steps, "dtype": "f64", "gpu_launches_per_step": launches,
           "roofline": {"bound": "hbm" if algo == "minsum" else "fp64 transcendental (exp/log) pipe",
                        "bytes_per_frame": bytes_frame,
                        "whole_decode_frac": bytes_frame * B * steps / (ms * 1e-3) / 1e9 / peak},
           "parity_sample": {"frames": int(cols.size), "identical_hard_decision_frames": int(same.sum()),
                             "equal": bool(same.all()), "max_abs_llr_diff": float(np.abs(got - ref).max()),
                             "tolerance": "identical hard decisions on >= 99.99 % of frames (north_star); sample must be all-equal"},
           "timing": dict(LAST_TIMING)}
    del decodi, ch, out
    torch.cuda.empty_cache()
    return res


def leg_llr_schedules(pkg, steps, rank, B=16384, ebn0=2.4):
    """The opt-in layered schedule (csrc/llr_layered.cu) next to the reference's flooding schedule, float64 min-sum on
    the (3,6) n=8000 graph with early termination ON (the BER drivers' mode): ms per batch and passes until the batch
    stops.  No reference counterpart for the layered schedule (parity: its numpy restatement, tests/test_gpu_round2.py)."""
    import torch
    wl = workload("c1")
    N, M = wl["H"].shape[1], wl["H"].shape[0]
    quanti = pkg.AWGN_Channel_Quantizer(10 ** (-ebn0 / 10) / (2 * (N - M) / N), 3, T, 2000)
    quanti.seed = SEED
    quanti.set_stream(rank)
    quanti.llr_dtype = np.float64
    quanti.init_OpenCL_quanti(N, B, return_buffer_only=True)
    ch = quanti.quantize_direct_OpenCL_LLR(N, B)
    res = {"workload": "(3,6) n=8000 min-sum float64 i_max=50 ET on", "frames_per_step": B, "EbN0_dB": ebn0, "steps": steps}
    for sched in ("flooding", "layered"):
        decodi = pkg.Min_Sum_Decoder_class_irregular(wl["H"], IMAX, T, B)
        decodi.init_OpenCL_decoding(B, quanti.context)
        decodi.schedule = sched
        ms, out = time_steps(lambda: decodi.decode(ch, buffer_in=True, return_buffer=True), steps, 1)
        res[sched] = {"value": (N - M) * B * steps / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms / steps,
                      "i_num": int(decodi.last_i_num), "gpu_launches_per_step": decodi.info()[1],
                      "bit_errors": int(decodi.return_errors_all_zero(out))}
        if sched == "layered":
            res[sched]["layers"] = decodi.layer_count()
        del decodi, out
        torch.cuda.empty_cache()
    return res


def leg_early_termination(pkg, steps, rank, name="c1"):
    """The BER drivers' operating mode: early termination on.  Same workload and inputs three ways -- fixed i_max (the
    headline), the reference's batch-granular stop (discrete_LDPC_decoder.py:233,273) and the opt-in per-frame stop
    with frame compaction (ibldpc_decode_ib_perframe) -- with the iteration statistics that explain the difference."""
    import torch
    wl = workload(name)
    B = wl["B"]
    t, tb, quanti, decodi = build_ib(pkg, wl, B, rank)
    N, M = t.n_var, t.n_chk
    ch = quanti.quantize_direct_OpenCL(N, B)
    res = {"workload": wl["name"].replace("ET off", "ET on"), "frames_per_step": B, "EbN0_dB": wl["ebn0"], "steps": steps}
    outs = {}
    for mode, et in (("fixed_imax", False), ("batch_stop", True), ("per_frame_stop", "frame")):
        decodi.early_termination = et
        ms, out = time_steps(lambda: decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True), steps, 1)
        outs[mode] = out.tensor.clone()
        entry = {"value": (N - M) * B * steps / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms / steps,
                 "i_num": int(decodi.last_i_num), "gpu_launches_per_step": decodi.info()[1],
                 "ms_per_step_repeats": LAST_TIMING.get("ms_per_step_repeats")}
        if et == "frame":
            inum = decodi.last_i_num_per_frame.tensor.to(torch.float32)
            entry.update({"mean_i_num": float(inum.mean()), "frames_at_imax": int((inum >= IMAX).sum()),
                          "median_i_num": float(inum.median())})
        res[mode] = entry
    # a frame that stopped (in either mode) is a codeword; the per-frame result of a frame that ran to i_max equals the
    # fixed-i_max result
    inum = decodi.last_i_num_per_frame.tensor
    full = inum >= IMAX
    res["per_frame_equals_fixed_imax_on_unconverged_frames"] = bool(torch.equal(outs["per_frame_stop"][:, full], outs["fixed_imax"][:, full]))
    res["bit_errors"] = {m: int((o[: N if not wl["irregular"] else int(decodi.data_len)] < T // 2).sum()) for m, o in outs.items()}
    # oracle parity sample of the per-frame mode: every sampled frame decoded on its own (msg_at_time = 1, early
    # termination on) must give the same output column AND the same i_num
    from oracle import oracle
    cols = sample_columns(B, 16)
    idx = torch.from_numpy(cols).to(ch.tensor.device)
    ch_s = ch.tensor.index_select(1, idx).cpu().numpy()
    got = outs["per_frame_stop"].index_select(1, idx).cpu().numpy()
    got_inum = inum.index_select(0, idx).cpu().numpy()
    ok = True
    for j in range(cols.size):
        ref, ref_inum = oracle.ib_decode(t, np.ascontiguousarray(ch_s[:, j:j + 1]), T=T, imax=IMAX, cn_lut=tb.Trellis_checknodevector_a,
                                         vn_lut=tb.Trellis_varnodevector_a, cn_match=tb.matching_vector_checknode,
                                         vn_match=tb.matching_vector_varnode, early=True)
        ok &= bool(np.array_equal(got[:, j], ref[:, 0].astype(np.uint8))) and int(got_inum[j]) == int(ref_inum)
    res["per_frame_parity_sample"] = {"frames": int(cols.size), "equal": bool(ok),
                                      "checker": "oracle/ldpc_oracle.c, one frame per call (the reference with msg_at_time = 1)"}
    del decodi, ch, outs
    torch.cuda.empty_cache()
    return res


def e2e_contracts(pkg, decodi, ch, N, K_info, B, T_, world, steps, rows_counted, dist):
    """The same metric through the class API with HOST buffers, copies inside the timed region, three contracts:
      packed        decode_packed: pinned nibble-packed cluster indices in, bit-packed hard decisions of the counted rows out
      uint8_pinned  decode_OpenCL(pinned uint8 numpy) -> uint8 numpy (round 1's e2e)
      int32_numpy   decode_OpenCL(int32 numpy) -> int32 numpy: the reference's own contract
                    (discrete_LDPC_decoder.py:207-209, :292-295), pageable memory, host-side casts included."""
    import torch
    host_u8 = pkg.pinned_empty((N, B), np.uint8)
    host_u8[:] = ch.get()
    res = {}

    def timed(fn, reps):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return K_info * B * world * reps / float(tm.item()) / 1e9

    # packed
    host_p = pkg.pinned_empty((N, (B + 1) // 2), np.uint8)
    decodi.pack_channel_values(host_u8, out=host_p)
    sink = [0]

    def run_packed():
        bits = decodi.decode_packed(host_p, B, rows=rows_counted)
        sink[0] += int(bits[0, 0])               # touch the result on the host

    v = timed(run_packed, steps)
    res["packed"] = {"value": v, "h2d_bytes_per_step": int(host_p.size), "d2h_bytes_per_step": int(rows_counted) * ((B + 7) // 8),
                     "api": "decode_packed(pinned nibble-packed uint8) -> bit-packed hard decisions -> ibldpc_decode_ib_host_packed"}
    # check the packed result against the device-buffer decode of the same inputs
    bits = decodi.unpack_bits(decodi.decode_packed(host_p, B, rows=rows_counted), B)
    dev_out = decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True).tensor[:rows_counted]
    res["packed"]["matches_device_path"] = bool(np.array_equal(bits, (dev_out < T_ // 2).cpu().numpy().astype(np.uint8)))

    decodi.host_output_dtype = np.uint8

    def run_u8():
        r = decodi.decode_OpenCL(host_u8, buffer_in=False, return_buffer=False)
        sink[0] += int(r[0, 0])

    v = timed(run_u8, steps)
    res["uint8_pinned"] = {"value": v, "h2d_bytes_per_step": int(N) * int(B), "d2h_bytes_per_step": int(N) * int(B),
                           "api": "decode_OpenCL(pinned uint8 numpy, buffer_in=False, return_buffer=False) -> ibldpc_decode_ib_host"}
    decodi.host_output_dtype = np.int32
    host_i32 = host_u8.astype(np.int32)

    def run_i32():
        r = decodi.decode_OpenCL(host_i32, buffer_in=False, return_buffer=False)
        sink[0] += int(r[0, 0])

    v = timed(run_i32, max(2, steps // 3))
    res["int32_numpy"] = {"value": v, "h2d_bytes_per_step": int(N) * int(B), "d2h_bytes_per_step": int(N) * int(B),
                          "host_bytes_touched_per_step": 8 * int(N) * int(B),
                          "api": "decode_OpenCL(int32 numpy, pageable) -> int32 numpy (class default, the reference's contract) -> ibldpc_decode_ib_host_i32"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "wlan", "wlan1944", "dvbs2"])
    ap.add_argument("--card", type=int, default=T, help="|T| of channel and decoder (16 = BASELINE; 32 = the reference's WLAN design)")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU and step (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the short legs of the other BASELINE configurations")
    ap.add_argument("--leg-steps", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / warnings never land on stdout
    import torch
    import torch.distributed as dist
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.parallel import bind_to_gpu_numa_node, counter_allreduce_fn, init_distributed

    rank, world, local = init_distributed(args.gpus)
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None   # host buffers of the e2e leg local to the GPU
    wl = workload(args.workload)
    B = args.frames or wl["B"]
    T_ = args.card
    # --- objects the BER drivers build (Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:76-91)
    t, tb, quanti, decodi = build_ib(pkg, wl, B, rank, T_)
    N, M, E = t.n_var, t.n_chk, t.n_edge
    K_info = N - M
    ch = quanti.quantize_direct_OpenCL(N, B)      # resident in HBM before the timed region
    torch.cuda.synchronize()

    # Per step: decode, count bit/frame errors on the device, all-reduce the 4 counters over the ranks
    # (NCCL, on the stream) and accumulate.  Nothing is read back inside the loop: a BER driver looks at
    # the totals one batch late, so the GPU never idles on the host.
    from informationbottleneckdecodingldpc_b200.engine import count_errors_async
    totals = torch.zeros(4, dtype=torch.int64, device="cuda")
    base = torch.tensor([0, 0, B, IMAX * B], dtype=torch.int64, device="cuda")
    rows_counted = N if not wl["irregular"] else int(decodi.data_len)
    # the one collective of the path goes through the library's own NCCL binding (C ABI: ibldpc_allreduce_counters);
    # torch.distributed stays the fallback when libnccl cannot be bound at run time
    allreduce = None
    if world > 1:
        os.environ.setdefault("IBLDPC_ABI_ALLREDUCE", "1")
        try:
            allreduce = counter_allreduce_fn(decodi)
        except Exception as e:   # noqa: BLE001
            print(f"[bench] C-ABI all-reduce unavailable ({e}); using torch.distributed", file=sys.stderr)
            os.environ["IBLDPC_ABI_ALLREDUCE"] = "0"
            allreduce = counter_allreduce_fn(decodi)
    abi_allreduce = bool(getattr(decodi, "_nccl_ready", False))

    def step():
        out = decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
        c = base.clone()
        count_errors_async(out, rows_counted, T_ // 2, c)
        if world > 1:
            allreduce(c)
        totals.add_(c)
        return out

    # Warm-up: W steps with the allocation pattern of the timed loop (the previous output stays alive while the next
    # one is produced, so the caching allocator holds both buffers before the clock starts) and, beyond W, until two
    # consecutive steps agree within 2 % and NVML shows the SM clock back at its maximum (at most 5 s more): the timed
    # region starts on a GPU in its steady state, then EXACTLY `steps` steps are timed.
    out_last = None
    warm_sampler = ClockSampler(local)
    warm_sampler.start()
    t_warm, n_warm, prev_ms = time.perf_counter(), 0, None
    while True:
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        out_last = step()
        w1.record()
        torch.cuda.synchronize()
        cur_ms = w0.elapsed_time(w1)
        n_warm += 1
        stable = prev_ms is not None and abs(cur_ms - prev_ms) <= 0.02 * max(cur_ms, prev_ms)
        prev_ms = cur_ms
        sm_now, sm_max = _sm_clock_now(warm_sampler) if warm_sampler.nvml is not None else (None, None)
        clock_ok = sm_now is None or sm_max is None or sm_now >= 0.97 * sm_max
        ready = n_warm >= max(args.warmup, 1) and ((stable and clock_ok) or time.perf_counter() - t_warm > 5.0 + 0.001 * cur_ms * args.warmup)
        if world > 1:   # step() holds a collective: every rank leaves the warm-up after the same step
            flag = torch.tensor([1 if ready else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ready = bool(flag.item())
        if ready:
            break
    warm_sampler.stop()
    launches_per_step = decodi.info()[1] + 2      # decode kernels + the two error-count kernels (torch/NCCL kernels not counted)
    totals.zero_()
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    # EXACTLY `steps` steps between two events; an event after every step as well, so that a stall of the box inside the
    # region (seen on this pool: 50-110 ms of a 436 ms region, kernel times unchanged, clocks at maximum) is visible.
    # If one step takes more than 1.25x the median step, the region is timed ONCE more and the second region is reported
    # ("timed_regions" keeps both).
    timed_regions = []
    for attempt in range(2):
        totals.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            out_last = step()
            evs[i + 1].record()
        torch.cuda.synchronize()
        ms = evs[0].elapsed_time(evs[-1])
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
        timed_regions.append({"ms_per_step": ms / args.steps, "max_step_ms": max(per_step), "median_step_ms": statistics.median(per_step)})
        outlier = max(per_step) > 1.25 * statistics.median(per_step)
        if world > 1:
            flag = torch.tensor([1 if outlier else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            outlier = bool(flag.item())
        if not outlier:
            break
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    rank_ms = [None] * world
    if world > 1:
        dist.all_gather_object(rank_ms, ms)
    else:
        rank_ms = [ms]
    ms = float(max(rank_ms))
    frames_total = B * world * args.steps
    value = K_info * frames_total / (ms * 1e-3) / 1e9

    # --- parity sample at the timed geometry (rank 0): 32 decoded columns of the last timed batch vs the CPU oracle
    parity = parity_sample_ib(decodi, ch, out_last, t, tb, wl, T_) if rank == 0 else None

    # --- roofline of the dominant kernel: per-launch CUDA-event times on the launch stream
    fam = decodi.info()[0]
    packed = fam == 2
    stored_div = 2 if packed else 1
    roofline = phase_roofline(decodi, ch, N, E, B, args.workload)

    # --- end to end through the class API with HOST buffers (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        contracts = e2e_contracts(pkg, decodi, ch, N, K_info, B, T_, world, args.steps, rows_counted, dist)
        head = contracts["packed"] if T_ <= 16 else contracts["uint8_pinned"]
        e2e = {"value": head["value"], "unit": "Gbit/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"],
               "d2h_bytes_per_step": head["d2h_bytes_per_step"], "api": head["api"],
               "contract": "packed" if T_ <= 16 else "uint8_pinned", "contracts": contracts}

    legs = None
    cpu = None
    if rank == 0 and world == 1:
        del out_last
        if not args.no_legs and args.workload == "c1" and T_ == T:
            del ch, decodi
            torch.cuda.empty_cache()
            legs = {}
            for name in ("wlan", "wlan1944", "dvbs2"):
                legs[name] = leg_ib(pkg, name, args.leg_steps, rank)
            legs["wlan_T32"] = leg_ib(pkg, "wlan", args.leg_steps, rank, T_=32)
            legs["c1_early_termination"] = leg_early_termination(pkg, args.leg_steps, rank)
            legs["wlan_early_termination"] = leg_early_termination(pkg, args.leg_steps, rank, name="wlan")
            legs["minsum_f64"] = leg_llr(pkg, "minsum", args.leg_steps, rank)
            legs["minsum_f64_schedules_early_termination"] = leg_llr_schedules(pkg, args.leg_steps, rank)
            legs["bp_f64"] = leg_llr(pkg, "bp", args.leg_steps, rank)
            legs["reference_batch_sizes"] = leg_reference_batches(pkg, rank)
        if not args.no_cpu_baseline:
            use_all_host_threads()
            frames = size_cpu_sample(wl, 12.0, T_)
            dtc, kind, cores, _ = cpu_reference_run(wl, frames, T_=T_)
            cpu = {"value": K_info * frames / dtc / 1e9, "unit": "Gbit/s", "cores": cores, "kind": kind,
                   "sample": f"{frames} frames of the workload (same tables, i_max={IMAX}, ET off) in {dtc:.1f} s"}

    if rank == 0:
        tot = totals.tolist()
        line = {
            "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_steps_run": n_warm, "timed_regions": timed_regions, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u4" if packed else "u8", "data": "synthetic",
            "config": {"workload": wl["name"] if T_ == T else wl["name"].replace("|T|=16", f"|T|={T_}"),
                       "frames_per_gpu_per_step": B, "n_var": N, "n_chk": M, "n_edge": E,
                       "info_bits": K_info, "i_max": IMAX, "EbN0_dB": wl["ebn0"], "tables": wl.get("tables"),
                       "l2": "inputs_larger_than_L2 (message array %.0f MB)" % (E * B / stored_div / 1e6), "parallelism": f"frames sharded x{world}", "rank0_numa_node": numa_node,
                       "fast_path": bool(fam), "kernel_family": {0: "generic", 1: "uint8", 2: "packed-nibble (n4)", 3: "t32"}.get(fam),
                       "counter_allreduce": "ibldpc_allreduce_counters (C ABI, NCCL)" if abi_allreduce else ("torch.distributed NCCL" if world > 1 else None)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "parity_sample": parity, "cpu_baseline": cpu, "workloads": legs,
            "errors": {"bit": tot[0], "frame": tot[1], "frames": tot[2]},
            "frames_per_s": frames_total / (ms * 1e-3), "ms_per_step_by_rank": [m / args.steps for m in rank_ms],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
