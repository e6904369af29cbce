"""Shim loader that imports the UNMODIFIED reference from /root/reference on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (fixture generation) and by
`tests/test_oracle_pins_reference.py` (skipped when /root/reference is absent, e.g. on the GPU box).
Nothing under `sttode_b200/` may import this file.

The five shims of SURVEY.md §8(c); none alters the reference's arithmetic except #3, which restates the
one third-party routine the path calls (torchdiffeq==0.2.3, requirement.txt:195, not installed here):

 1. stub modules `matplotlib`, `matplotlib.pyplot`, `matplotlib.image`, `glob2`
    (hypertransformer.py:3, model/utils.py:8 import them; never used on the path);
 2. `torch.nn.modules.linear._LinearWithBias` alias (hypertransformer.py:11; imported, never used);
 3. `torchdiffeq.odeint(func, y0, t, method='euler')`: fixed-grid explicit Euler over the user grid `t`
    (FixedGridODESolver with step_size=None uses `t` itself as the grid; Euler step = dt * f(t0, y0));
    returns the stacked solution so `[1]` is the state at t[-1]  (ode_demo.py:188);
 4. `Tensor.cuda()` -> identity and `device='cuda'` factory kwargs -> cpu
    (hypertransformer.py:69, model/STTODE.py:333-334 hard-code cuda);
 5. `/root/reference` prepended to sys.path (flat top-level imports).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("STTODE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hyptransformerlib.py"))


def _odeint(func, y0, t, method="euler", **kw):          # shim 3 (<= 10 lines, auditable)
    assert method == "euler"
    ys, y = [y0], y0
    for i in range(len(t) - 1):
        y = y + (t[i + 1] - t[i]) * func(t[i], y)
        ys.append(y)
    return torch.stack(ys)


_installed = False


def install():
    """Install the shims (idempotent) and return a namespace with the reference modules."""
    global _installed
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if not _installed:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "glob2"):      # shim 1
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
        import torch.nn.modules.linear as _lin                                              # shim 2
        if not hasattr(_lin, "_LinearWithBias"):
            _lin._LinearWithBias = _lin.NonDynamicallyQuantizableLinear
        tde = types.ModuleType("torchdiffeq")                                               # shim 3
        tde.odeint = _odeint
        tde.odeint_adjoint = _odeint
        sys.modules.setdefault("torchdiffeq", tde)
        if not torch.cuda.is_available() or os.environ.get("STTODE_REF_FORCE_CPU", "1") == "1":   # shim 4
            torch.Tensor.cuda = lambda self, *a, **k: self
            for fname in ("zeros", "ones", "empty", "randn", "rand", "full", "tensor"):
                _orig = getattr(torch, fname)

                def _wrap(*a, __orig=_orig, **k):
                    if str(k.get("device", "")).startswith("cuda"):
                        k["device"] = "cpu"
                    return __orig(*a, **k)
                setattr(torch, fname, _wrap)
        if REFERENCE_ROOT not in sys.path:                                                  # shim 5
            sys.path.insert(0, REFERENCE_ROOT)
        _installed = True
    import hyptransformerlib, hypertransformer, ode_demo            # noqa: E401
    import core.manifolds.oblique as oblique
    import hyptorch.pmath as pmath
    import model.STTODE as sttode
    import utils.metrics as metrics
    return types.SimpleNamespace(hyptransformerlib=hyptransformerlib, hypertransformer=hypertransformer,
                                 ode_demo=ode_demo, oblique=oblique, pmath=pmath, sttode=sttode,
                                 metrics=metrics)


from sttode_b200.synth import make_args          # noqa: E402,F401  (the args Namespace STTODENet reads)
