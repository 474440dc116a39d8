"""Handle lifetime of the C ABI: a garbage-collected binding finalises a model and its context in any order.
Destroying a context destroys the models that run on it, destroying a stale handle is a no-op and using one is an
error (never a crash)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from shamrock_b200 import _capi  # noqa: E402
from tests import scenarios as S  # noqa: E402


def test_context_destroyed_before_its_model():
    sc = S.periodic_box(1500, "M4", "constant", jitter=0.1)
    ctx = _capi.Context(0)
    m = S.make_cuda(sc, ctx=ctx)
    m.evolve_once()
    h_model, h_ctx = m.h, ctx.h
    L = _capi.lib()
    assert L.shamb200_ctx_destroy(h_ctx) == 0  # takes the model with it
    assert m.patch_count == 0 and m.patch_size(0) == 0
    with pytest.raises(_capi.ShamB200Error):
        m.evolve_once()
    with pytest.raises(_capi.ShamB200Error):
        m.get(0, "xyz")
    assert L.shamb200_model_destroy(h_model) == 0  # stale: no-op
    assert L.shamb200_ctx_destroy(h_ctx) == 0      # stale: no-op
    out = C.c_void_p()
    assert L.shamb200_model_create(h_ctx, C.byref(_capi.default_config()), C.byref(out)) != 0  # stale context
    m.h, ctx.h = None, None


def test_model_destroyed_before_its_context_and_twice():
    sc = S.periodic_box(1500, "M4", "constant", jitter=0.1)
    ctx = _capi.Context(0)
    m = S.make_cuda(sc, ctx=ctx)
    m.evolve_once()
    h = m.h
    m.close()
    assert _capi.lib().shamb200_model_destroy(h) == 0
    m2 = S.make_cuda(sc, ctx=ctx)  # the context is still usable
    st = m2.evolve_once()
    assert st["npart"] == len(sc["xyz"])
    ctx.close()  # closes m2 first
    assert m2.h is None
    m2.close()
    ctx.close()
