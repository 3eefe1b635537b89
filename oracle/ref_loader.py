"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference (jiho264/P2-ViT) from
/root/reference so that golden vectors can be generated in the build container.

Nothing in the product (`p2vit_b200/`), in the `-m gpu` tests, in `smoke()` or in
`bench.py` imports this module: /root/reference does not exist on the GPU box.  Only
`oracle/gen_golden.py` and the container-only `tests/test_oracle_vs_reference.py`
(skipped automatically when /root/reference is absent) use it.

The reference cannot be imported as shipped; the shims installed
here do not edit it:
  * matplotlib / timm are absent            -> stub modules in sys.modules
    (models/plot_distrib.py:1-4, utils/build_model.py:5-8)
  * hard-coded `.cuda()` calls              -> identity when no CUDA device
    (quantizer/uniform.py:83,125; observer/minmax.py:53-61,146-150; ptf.py:55-63,73)
  * omse observer rejects the kwargs QAct passes (layers.py:251-253 vs omse.py:30)
    -> kwargs-tolerant subclass registered under the same name
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("P2VIT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "ptq", "layers.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_loaded = None


def load_reference():
    """Returns (models_module, Config_class) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.gridspec = _stub("matplotlib.gridspec")
    mpl.collections = _stub("matplotlib.collections", PolyCollection=object)
    _stub("timm")
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the repo under test may shadow top-level names; make sure `models`/`config`
    # resolve to the reference
    for name in ("models", "config"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    import models as ref_models  # noqa: E402
    from config import Config as RefConfig  # noqa: E402

    # Q3: kwargs-tolerant omse
    from models.ptq.observer import build as obuild
    from models.ptq.observer.omse import OmseObserver

    class _OmseTolerant(OmseObserver):
        def get_quantization_params(self, inputs, *args, **kwargs):
            return OmseObserver.get_quantization_params(self, inputs)

    obuild.str2observer["omse"] = _OmseTolerant
    _loaded = (ref_models, RefConfig)
    return _loaded
