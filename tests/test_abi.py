"""CPU-side checks of the drop-in boundary: libbmpc.so loads, exports every symbol include/bmpc.h
declares, and refuses to run without a CUDA device (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mpc_b200
from mpc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bm(?:pc|he)_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 23
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/bmpc.h but not exported by libbmpc.so"
    assert sorted(_lib.SYMBOLS) == syms
    assert L.bmpc_version() >= 100


def test_struct_layouts_match_header():
    # bmpc_dims: 12 int32 + 1 double; bmpc_step_io: 12 pointers + 2 int32; bmpc_info: 6 pointers
    assert C.sizeof(_lib.Dims) == 12 * 4 + 8
    assert C.sizeof(_lib.StepIO) == 12 * 8 + 16 + 3 * 8  # 12 pointers + device_ptrs, sync, resident, host_mapped + y0m, Yhat_s, kkt
    assert C.sizeof(_lib.Info) == 6 * 8
    assert C.sizeof(_lib.Softness) == 8 * 8
    assert C.sizeof(_lib.MheDims) == 12 * 4 + 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(mpc_b200.BmpcError) as e:
        mpc_b200.BatchLinMPC(4, 2, 2, 6, 20, 5)
    assert e.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_validation_without_device():
    L = _lib.lib()
    h = C.c_void_p()
    dims = _lib.Dims(N=4, nu=2, ny=2, nd=0, nxhat=6, Hp=20, Hc=5, neps=1)
    nb = (C.c_int32 * 5)(1, 1, 1, 1, 15)  # sum != Hp
    assert L.bmpc_create(C.byref(h), C.byref(dims), nb) == _lib.ERR_ARG
    assert b"move_blocking" in L.bmpc_last_error()
    dims.Hc = 25
    assert L.bmpc_create(C.byref(h), C.byref(dims), nb) == _lib.ERR_ARG
    assert L.bmpc_step(None, None) == _lib.ERR_ARG
    assert L.bmpc_destroy(None) == _lib.OK
    # entry points added for the prediction-form MHE and the time-varying KalmanFilter: null handles are refused
    assert L.bmhe_update_solve(None, None, None, None, None, None, None, None, None, None, None) == _lib.ERR_ARG
    assert L.bmpc_set_estimator_cov(None, None, None, None) == _lib.ERR_ARG
    assert L.bmpc_get_cov(None, None) == _lib.ERR_ARG
    mh = C.c_void_p()
    md = _lib.MheDims(N=2, nu=1, nym=1, nd=0, nxhat=40, He=3, neps=0, direct=0)
    assert L.bmhe_create(C.byref(mh), C.byref(md)) == _lib.ERR_UNSUPPORTED  # nxhat > 32
    md.nxhat, md.He = 2, 0
    assert L.bmhe_create(C.byref(mh), C.byref(md)) == _lib.ERR_ARG


def test_host_mirror_constructors():
    """move_blocking / augment_model / batched Kalman gain of the host mirror against the oracle."""
    from oracle.linmpc import LinModel as OL, SteadyKalmanFilter as OS, move_blocking as omb
    from mpc_b200 import workloads
    for Hp, Hc in [(10, 2), (20, 5), (10, [1, 2, 3]), (10, [1, 2, 3, 6, 7])]:
        assert mpc_b200.move_blocking(Hp, Hc) == omb(Hp, Hc)
    m, _ = workloads.random_plants(6, 4, 2, 2, 1)
    skf = mpc_b200.SteadyKalmanFilter(m)
    for i in range(6):
        o = OS(OL(m.A[i], m.Bu[i], m.C[i]))
        assert np.allclose(o.Ahat, skf.Ahat[i]) and np.allclose(o.Chat, skf.Chat[i])
        assert np.allclose(o.Khat, skf.Khat[i], rtol=1e-9, atol=1e-11)
    sp = workloads.setpoints(np.random.default_rng(0), 5, 2, 60)
    assert sp.shape == (60, 5, 2) and set(np.unique(sp)) <= {-1.0, 1.0}
    assert (sp[0] == sp[24]).all()
