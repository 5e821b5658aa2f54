import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "H_indptr" in d:
        n = d["H_indices"].size
        d["H"] = sp.csr_matrix((np.ones(n, dtype=np.int8), d["H_indices"], d["H_indptr"]), shape=tuple(d["H_shape"]))
    return d


IB_CASES = sorted(f[:-4] for f in os.listdir(GOLD) if f.startswith("ib_"))
LLR_CASES = sorted(f[:-4] for f in os.listdir(GOLD) if f.startswith("llr_"))
ENC_CASES = sorted(f[:-4] for f in os.listdir(GOLD) if f.startswith("enc_"))


@pytest.fixture(scope="session")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
