"""A/B of kernel options through environment switches, per-phase times (scratch)."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def run(env, label, wl="c1", B=32768):
    e = dict(os.environ); e.update(env)
    cmd = [sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--workload", wl, "--frames", str(B)]
    p = subprocess.run(cmd, env=e, capture_output=True, text=True)
    for l in p.stdout.splitlines():
        if l.startswith("{"):
            d = json.loads(l); r = d["roofline"]
            print(f"{label:28s} {wl:6s} B={B} {d['value']:.3f} Gbit/s cn_ms {r.get('cn_avg_ms',0):.4f} vn_ms {r.get('vn_avg_ms',0):.4f}", flush=True)
            return
    print(label, "FAILED", p.stderr[-400:])
if __name__ == "__main__":
    for wl, B in (("wlan", 65536), ("dvbs2", 4096), ("c1", 65536)):
        run({"IBLDPC_VN_PAIR_MIN_DEGREE": "99"}, "no vn pair", wl, B)
        for md in (3, 4, 5):
            run({"IBLDPC_VN_PAIR_MIN_DEGREE": str(md)}, f"vn pair d>={md} 512thr", wl, B)
        run({"IBLDPC_VN_PAIR_MIN_DEGREE": "5", "IBLDPC_VN_PAIR_THREADS": "256"}, "vn pair d>=5 256thr", wl, B)
