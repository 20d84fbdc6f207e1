"""The C-ABI library builds, loads, and exports every symbol include/avc_b200.h declares.
No compute calls here (no GPU): avc_create must fail loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from attack_vc_b200 import _lib
    from attack_vc_b200.build import build_library
    build_library()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "avc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from attack_vc_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/avc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_struct_layouts_match_header():
    from attack_vc_b200 import _lib
    # sizes computed from the C declarations (natural alignment)
    assert C.sizeof(_lib.EncoderDesc) == 8 * 4 + 8 * 4 + 4
    assert C.sizeof(_lib.DecoderDesc) == 6 * 4 + 8 * 4 + 4
    assert C.sizeof(_lib.WeightView) == 8 + 8 + 8 + 32
    assert C.sizeof(_lib.AttackArgs) % 8 == 0
    assert _lib.AttackArgs.inv_norm.offset % 8 == 0


def test_version_and_no_cpu_fallback(lib):
    import torch
    assert b"sm_100a" in lib.avc_version()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    from attack_vc_b200 import _lib
    h = C.c_void_p()
    d = _lib.ModelDesc()
    rc = lib.avc_create(C.byref(h), C.byref(d), 0)
    assert rc == -2 and b"no CPU fallback" in lib.avc_last_error(None)


def test_engine_refuses_cpu(oracle, cpu_model):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from attack_vc_b200 import AvcError, Engine
    with pytest.raises(AvcError):
        Engine(cpu_model)


def test_round2_struct_layouts_match_header():
    """ctypes mirrors of the structs added for the VSMask trainer: sizes / offsets computed from the C declarations
    (natural alignment; include/avc_b200.h avc_spk_grad_args, avc_pm_trainer_args)."""
    from attack_vc_b200 import _lib
    # avc_spk_grad_args: ptr, i64[3], i32 B, i32 T | ptr, i64[3] | ptr, i64[3], i32 T_tgt (+4 pad) | ptr, i64[3] | ptr | f32 (+4 pad) | f64 | i32 (+4 pad)
    assert C.sizeof(_lib.SpkGradArgs) == (8 + 24 + 8) + (8 + 24) + (8 + 24 + 8) + (8 + 24) + 8 + 8 + 8 + 8
    assert _lib.SpkGradArgs.inv_norm.offset % 8 == 0 and _lib.SpkGradArgs.grad_out.offset % 8 == 0
    # avc_pm_trainer_args: 4 x i32, 7 x f32 (+4 pad), f64
    assert C.sizeof(_lib.PmTrainerArgs) == 16 + 28 + 4 + 8
    assert _lib.PmTrainerArgs.inv_norm.offset == 48


def test_predictive_engine_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from attack_vc_b200 import AvcError
    from attack_vc_b200.predictive import PredictiveEngine
    from attack_vc_b200.synthetic import pm_make_state_dict
    with pytest.raises(AvcError):
        PredictiveEngine(pm_make_state_dict(0))


def test_audio_engine_refuses_cpu_and_struct_layout():
    import torch
    from attack_vc_b200 import _lib
    assert C.sizeof(_lib.AudioDesc) == 5 * 4 + 3 * 4          # avc_audio_desc: five int32, three float
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from attack_vc_b200 import AvcError
    from attack_vc_b200.audio import AudioEngine
    with pytest.raises(AvcError):
        AudioEngine()
