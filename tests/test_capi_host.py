"""CPU: the C-ABI library loads, exports every symbol include/nfp_b200.h declares, and its host-only
entry points (shape / workspace / path queries, argument validation) behave.  No compute calls."""
import ctypes
import os
import re

import pytest

from neighbour_feature_pooling_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _desc(**kw):
    base = dict(dtype=_capi.F32, B=2, C=8, H=7, W=7, R=1, stride=1, padding=1, dilation=1,
                padding_mode="reflect", measure="cosine", similarity=True, difference_taps=False,
                eps=1e-6, p=1, q_scs=1e-6, path="auto")
    base.update(kw)
    return _capi.make_desc(**base)


def test_header_symbols_are_exported():
    with open(os.path.join(ROOT, "include", "nfp_b200.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(nfpb200_[a-z_]+)\s*\(", header)))
    assert declared, "no entry points found in the header"
    lib = _capi.load()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/nfp_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == declared
    assert lib.nfpb200_abi_version() == _capi.ABI_VERSION == int(re.search(r"NFPB200_ABI_VERSION (\d+)", header).group(1))


def test_desc_struct_matches_header_layout():
    # ABI v3: 18 x 4-byte fields, layout + inner_R (2 x 4), two 8-byte batch strides; no padding
    assert ctypes.sizeof(_capi.Desc) == 96
    assert _capi.Desc.layout.offset == 72 and _capi.Desc.inner_R.offset == 76 and _capi.Desc.x_batch_stride.offset == 80


def test_multi_radius_descriptor():
    """desc.inner_R: y / gy = [radius-r map | radius-R map]; the fused kernels' map modes or nothing."""
    lib = _capi.load()
    n = ctypes.c_size_t()

    def d(**kw):
        inner = kw.pop("inner_R", 1)
        dd = _desc(**{"B": 4, "C": 64, "R": 2, "padding": 2, **kw})
        dd.inner_R = inner
        return dd
    assert _capi.describe_path(d(), _capi.OP_FORWARD) == "fused/stream_7x7_r2"
    assert _capi.describe_path(d(H=14, W=14, dtype=_capi.BF16), _capi.OP_BACKWARD) == "fused/stream_14x14_r2"
    assert _capi.workspace_bytes(d(), _capi.OP_BACKWARD) == 0 and _capi.launch_count(d(), _capi.OP_BACKWARD) == 1
    tok = d(dtype=_capi.BF16)
    tok.layout = _capi.LAYOUT_NHWC
    assert _capi.describe_path(tok, _capi.OP_BACKWARD) == "fused/token_7x7_r2"
    # the inner radius must be smaller than R; pooled / head modes, generic and split paths do not carry it
    for bad in (d(inner_R=2), d(inner_R=-1), d(R=1, padding=1, inner_R=1)):
        assert lib.nfpb200_workspace_bytes(ctypes.byref(bad), _capi.OP_FORWARD, ctypes.byref(n)) == -1
    assert lib.nfpb200_workspace_bytes(ctypes.byref(d()), _capi.OP_POOL_FORWARD, ctypes.byref(n)) == -5
    assert lib.nfpb200_head_supported(ctypes.byref(d())) == -5
    for path in ("generic", "split"):
        assert lib.nfpb200_workspace_bytes(ctypes.byref(d(path=path)), _capi.OP_FORWARD, ctypes.byref(n)) == -5
    # maps / geometries outside the fused kernels: refused (the caller launches once per radius)
    # other map sizes with 16-byte aligned planes: the planar row-band kernels, still one launch and no workspace
    assert _capi.describe_path(d(H=28, W=28), _capi.OP_BACKWARD) == "planar/band"
    assert _capi.workspace_bytes(d(H=28, W=28), _capi.OP_BACKWARD) == 0 and _capi.launch_count(d(H=28, W=28), _capi.OP_FORWARD) == 1
    assert lib.nfpb200_workspace_bytes(ctypes.byref(d(H=28, W=28, path="fused")), _capi.OP_FORWARD, ctypes.byref(n)) == -5
    for bad in (d(H=9, W=9), d(stride=2), d(measure="dot"), d(padding_mode="circular")):
        assert lib.nfpb200_workspace_bytes(ctypes.byref(bad), _capi.OP_FORWARD, ctypes.byref(n)) == -5


def test_channels_last_layout_selection():
    """NHWC descriptors: served by the tensor-core token kernels for bf16 / cosine / pad = R, refused otherwise
    (the caller repacks to NCHW); batch strides must be multiples of 8 elements (16-byte copies)."""
    def d(**kw):
        dd = _desc(**{k: v for k, v in kw.items() if k not in ("dtype", "layout", "xbs")})
        dd.dtype = kw.get("dtype", _capi.BF16)
        dd.layout = kw.get("layout", _capi.LAYOUT_NHWC)
        dd.x_batch_stride = kw.get("xbs", 0)
        return dd
    assert _capi.describe_path(d(B=4, C=512), _capi.OP_FORWARD) == "fused/token_7x7_r1"
    assert _capi.describe_path(d(B=4, C=192, H=14, W=14, xbs=197 * 192), _capi.OP_BACKWARD) == "fused/token_14x14_r1"
    assert _capi.describe_path(d(B=4, C=960), _capi.OP_POOL_BACKWARD) == "fused/token_7x7_r1"
    assert _capi.workspace_bytes(d(B=4, C=512), _capi.OP_BACKWARD) == 0
    lib = _capi.load()
    n = ctypes.c_size_t()
    for bad in (d(B=4, C=512, dtype=_capi.F32), d(B=4, C=100), d(B=4, C=64, H=9, W=9), d(B=4, C=512, measure="dot")):
        assert lib.nfpb200_workspace_bytes(ctypes.byref(bad), _capi.OP_FORWARD, ctypes.byref(n)) == -5
    assert lib.nfpb200_workspace_bytes(ctypes.byref(d(B=4, C=512, xbs=49 * 512 + 4)), _capi.OP_FORWARD, ctypes.byref(n)) == -7
    assert lib.nfpb200_workspace_bytes(ctypes.byref(d(B=4, C=512, xbs=8)), _capi.OP_FORWARD, ctypes.byref(n)) == -1


def test_output_shape_rule():
    assert _capi.output_shape(_desc()) == (7, 7)
    assert _capi.output_shape(_desc(H=14, W=9, R=2, padding=2)) == (14, 9)
    assert _capi.output_shape(_desc(H=8, W=7, stride=2)) == (4, 4)
    assert _capi.output_shape(_desc(H=9, W=9, dilation=2, padding=2)) == (9, 9)
    assert _capi.output_shape(_desc(H=5, W=5, padding=0)) == (3, 3)


def test_argument_errors():
    lib = _capi.load()
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    d = _desc(H=2, W=2, padding=2)  # reflect pad >= dim
    assert lib.nfpb200_output_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == -2
    assert "Padding size should be less" in _capi.status_string(-2)
    d = _desc(H=2, W=2, padding=0)  # window larger than input
    assert lib.nfpb200_output_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == -3
    d = _desc()
    d.struct_bytes = 10
    assert lib.nfpb200_output_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == -1
    d = _desc()
    d.measure = 99
    assert lib.nfpb200_output_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == -1
    # null data pointers are rejected before anything touches the device
    assert lib.nfpb200_forward(ctypes.byref(_desc()), None, None, None, 0, None) == -1
    # misaligned tensor pointers are refused before any launch (the fused kernels use 16-byte TMA copies)
    assert lib.nfpb200_forward(ctypes.byref(_desc()), 0x1004, 0x2000, None, 0, None) == -7
    assert "16-byte" in _capi.status_string(-7)
    with pytest.raises(RuntimeError, match="invalid argument"):
        _capi.check(-1, "x")


def test_path_selection_and_workspace():
    fused = _desc(B=256, C=512)
    assert _capi.describe_path(fused, _capi.OP_FORWARD).startswith("fused/")
    assert _capi.workspace_bytes(fused, _capi.OP_BACKWARD) == 0
    assert _capi.launch_count(fused, _capi.OP_FORWARD) == 1
    assert _capi.launch_count(fused, _capi.OP_BACKWARD) == 1
    generic = _desc(B=4, C=8, H=9, W=9, stride=2)
    assert _capi.describe_path(generic, _capi.OP_FORWARD) == "generic/pairs"
    assert _capi.workspace_bytes(generic, _capi.OP_BACKWARD) > 0
    forced = _desc(B=4, C=8, H=9, W=9, stride=2, path="fused")
    n = ctypes.c_size_t()
    assert _capi.load().nfpb200_workspace_bytes(ctypes.byref(forced), _capi.OP_FORWARD, ctypes.byref(n)) == -5
    # maps outside the streaming kernels' shape list (multi-stage heads): planar kernels.  Aligned planes: fused row-band
    # kernels, ONE launch each way and no workspace; otherwise table / coefficient passes through the caller's workspace
    planar = _desc(B=2, C=16, H=112, W=112)
    assert _capi.describe_path(planar, _capi.OP_FORWARD) == "planar/band"
    assert _capi.describe_path(planar, _capi.OP_BACKWARD) == "planar/band"
    assert _capi.workspace_bytes(planar, _capi.OP_FORWARD) == 0 and _capi.workspace_bytes(planar, _capi.OP_BACKWARD) == 0
    assert _capi.launch_count(planar, _capi.OP_FORWARD) == 1 and _capi.launch_count(planar, _capi.OP_BACKWARD) == 1
    scalar = _desc(B=2, C=7, H=9, W=13)   # unaligned planes
    assert _capi.describe_path(scalar, _capi.OP_BACKWARD) == "planar/table"
    P = 9 * 13
    assert _capi.workspace_bytes(scalar, _capi.OP_FORWARD) >= 2 * 5 * P * 4
    assert _capi.workspace_bytes(scalar, _capi.OP_BACKWARD) >= 2 * (5 + 9) * P * 4
    assert _capi.launch_count(scalar, _capi.OP_FORWARD) == 2 and _capi.launch_count(scalar, _capi.OP_BACKWARD) == 3
    assert _capi.describe_path(planar, _capi.OP_POOL_FORWARD) == "generic/pairs"  # pooled mode: generic kernels
    assert _capi.describe_path(_desc(B=2, C=16, H=112, W=112, path="generic"), _capi.OP_FORWARD) == "generic/pairs"
    # the x-stable hint rides on `path` and changes neither the path choice nor the workspace; other bits are refused
    hinted = _desc(B=256, C=512)
    hinted.path |= _capi.HINT_X_STABLE
    assert _capi.describe_path(hinted, _capi.OP_BACKWARD) == _capi.describe_path(fused, _capi.OP_BACKWARD)
    assert _capi.workspace_bytes(hinted, _capi.OP_BACKWARD) == 0
    yf32 = _desc(B=256, C=512)
    yf32.path |= _capi.FLAG_Y_F32          # fp32 similarity map from bf16 x (autocast): a flag like the hint
    assert _capi.describe_path(yf32, _capi.OP_FORWARD) == _capi.describe_path(fused, _capi.OP_FORWARD)
    bad = _desc(B=256, C=512)
    bad.path |= 0x400
    assert _capi.load().nfpb200_workspace_bytes(ctypes.byref(bad), _capi.OP_BACKWARD, ctypes.byref(n)) == -1
    # every measure has a generic path
    for m in _capi.MEASURES:
        assert _capi.describe_path(_desc(measure=m, path="generic"), _capi.OP_BACKWARD) == "generic/pairs"
