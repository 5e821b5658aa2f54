"""Stage the reference's BER driver scripts for the replay tests (TEST INFRASTRUCTURE ONLY).

The GPU box has no /root/reference, and reference sources must not enter the repository, so -- exactly like
build_ref.py does for the kernel sources -- the unmodified driver scripts are copied from where they lie under the
reference tree into the git-ignored ``oracle/_ref/drivers/`` (same relative paths), which travels to the GPU box
with the snapshot.  ``tests/test_driver_replay.py`` executes them there through
``informationbottleneckdecodingldpc_b200.run_driver``.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_REF = "/root/reference"
DEST = os.path.join(HERE, "_ref", "drivers")
DRIVERS = [
    "Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py",
    "Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL_min_sum.py",
    "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL.py",
    "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_enc.py",
    "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_min_sum.py",
    "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_quant_BP.py",
    "Irregular_LDPC_Decoding/DVB-S2/BER_simulation_OpenCL.py",
]


def stage(ref: str = DEFAULT_REF) -> list[str]:
    out = []
    for rel in DRIVERS:
        src = os.path.join(ref, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        out.append(dst)
    return out


def staged(rel: str) -> str | None:
    p = os.path.join(DEST, rel)
    return p if os.path.exists(p) else None


if __name__ == "__main__":
    for p in stage():
        print("staged", p)
