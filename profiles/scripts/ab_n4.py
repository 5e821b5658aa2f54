"""Default configuration on all workloads (scratch)."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def run(env, label, wl="c1", B=0):
    e = dict(os.environ); e.update(env)
    cmd = [sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--workload", wl] + (["--frames", str(B)] if B else [])
    p = subprocess.run(cmd, env=e, capture_output=True, text=True)
    for l in p.stdout.splitlines():
        if l.startswith("{"):
            d = json.loads(l); r = d["roofline"]
            print(f"{label:14s} {wl:8s} B={d['config']['frames_per_gpu_per_step']} {d['value']:.3f} Gbit/s cn_ms {r.get('cn_avg_ms',0):.4f} vn_ms {r.get('vn_avg_ms',0):.4f}", flush=True)
            return
    print(label, "FAILED", p.stderr[-400:])
if __name__ == "__main__":
    for wl, B in (("c1", 0), ("c1", 5000), ("c1", 20000), ("wlan", 0), ("wlan1944", 0), ("dvbs2", 0), ("dvbs2", 2048)):
        run({}, "default", wl, B)
    run({"IBLDPC_CN_THREADS": "512", "IBLDPC_VN_THREADS": "256"}, "small CTAs", "c1", 5000)
    run({"IBLDPC_CN_THREADS": "512", "IBLDPC_VN_THREADS": "256"}, "small CTAs", "c1", 20000)
