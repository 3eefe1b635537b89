import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/models")
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present"))
        if "gpu" in it.keywords and not have_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        # The reference's fp32 row reductions (LayerNorm sum of squares, softmax sum) are split over the intra-op threads, so its
        # very logits depend on the thread count (ViT-L, mixed config: one of two images flips between 4 and 8 threads).  The golden
        # files record the count they were generated with; comparisons against them run the CPU oracle with the same count.
        if "meta.threads" in g.files:
            import torch
            torch.set_num_threads(int(g["meta.threads"]))
        return g

    return load
