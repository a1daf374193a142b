import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_TAGS = ("conmodC_P100", "conmodA_E1m2", "f7_chr1to10")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_batch():
    z = load_npz("reads_batch.npz")
    names = [str(x) for x in z.pop("contig_names")]
    lens = z.pop("contig_len")
    return z, names, lens


@pytest.fixture(params=MODEL_TAGS)
def model_tag(request):
    return request.param


def golden_model(tag):
    return load_npz("model_%s.npz" % tag)


def golden_windows(tag):
    return load_npz("windows_%s.npz" % tag)


def golden_reads(tag):
    z = load_npz("reads_%s.npz" % tag)
    z["bed"] = {str(k): str(t) for k, t in zip(z["bed_keys"], z["bed_text"])}
    z["base"] = str(z["base"])
    return z


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
