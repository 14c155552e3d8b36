"""Host-side tiling logic of the tensor-core convolution (csrc/conv_plan.h) through pg_conv_tc_plan: no GPU.
Pins the plan of every layer at the BASELINE shapes (inference C=512/T=696 and training C=1024/T=128)."""
import ctypes

import pytest

KEYS = ("n_tile", "n_ntiles", "nb", "strip_rows", "pair", "merged", "mgroups", "acc_stages", "n_chunks", "n_cotiles",
        "OS", "IS", "taps0", "taps1", "groups0", "groups1")
GEOM = {"d1": (0, 32, 2, 16, 1, 2), "d2": (0, 8, 1, 2, 2, 2), "d3": (0, 8, 2, 1, 2, 2), "d4": (0, 4, 2, 1, 2, 4),
        "u4": (1, 5, 2, 1, 4, 2), "u3": (1, 8, 2, 1, 4, 2), "u2": (1, 8, 1, 2, 4, 2), "u1": (1, 32, 2, 16, 4, 2)}


def lengths(T):
    L1 = T // 2 + 1; L2 = L1 - 3; L3 = L2 // 2 - 2; L4 = (L3 - 1) // 2
    return {"d1": T, "d2": L1, "d3": L2, "d4": L3, "u4": L4, "u3": L3, "u2": L2, "u1": L1}


def plan(layer, C, T, B, prec, **kw):
    from phasegen import _lib, ops
    kind, k, s, p, cim, com = GEOM[layer]
    L_in = lengths(T)[layer]
    rows = (L_in + 7) // 8 * 8
    d = ops.conv_desc(kind, B, C * cim, C * com, L_in, k, s, p, rows, C * cim, _lib.PRECISIONS[prec], **kw)
    out = (ctypes.c_int * 17)()
    rc = _lib.load().pg_conv_tc_plan(ctypes.byref(d), out, 17)
    assert rc == 0, _lib.last_error()
    return dict(zip(KEYS + ("whole_clip",), out)), d


def test_inference_shapes_use_pairs_and_full_width_tiles():
    for layer in GEOM:
        pl, d = plan(layer, 512, 696, 256, "bf16x3")
        assert pl["pair"] == 1 and pl["n_chunks"] == d.C_in // 64 and pl["n_cotiles"] == d.C_out // 128
        assert pl["n_tile"] % 16 == 0 and pl["n_tile"] <= 256
        assert pl["n_tile"] * pl["n_ntiles"] >= (d.L_out + pl["OS"] - 1) // pl["OS"]
    big, _ = plan("u1", 512, 696, 256, "bf16x3")
    assert (big["n_tile"], big["n_ntiles"], big["OS"], big["nb"], big["acc_stages"]) == (176, 2, 2, 1, 2)
    assert big["taps0"] == 16 and big["taps1"] == 16 and big["groups0"] == 1     # 16 taps share one activation strip
    assert big["strip_rows"] == 104                                                # half strip per CTA: 88 + 15 -> 104
    # two-product precision: the weight stream is the bound, two clips share every weight tile (one TMEM stage)
    fast, _ = plan("u1", 512, 696, 256, "f16x2")
    assert (fast["nb"], fast["acc_stages"], fast["merged"]) == (2, 1, 0)
    # short axis of the inference net (L = 85): merged clips, one MMA over two strips
    u4, _ = plan("u4", 512, 696, 256, "bf16x3")
    assert u4["merged"] == 1 and u4["nb"] == 2 and u4["mgroups"] == 1 and u4["nb"] * u4["strip_rows"] <= 256


def test_training_shapes_merge_clips():
    for layer in GEOM:
        pl, d = plan(layer, 1024, 128, 32, "bf16")
        assert pl["merged"] == 1, layer
        per_mma = pl["nb"] // pl["mgroups"]
        assert per_mma % 2 == 0                                   # a CTA pair splits a merged MMA by clips
        assert per_mma * pl["strip_rows"] <= 256 and pl["nb"] * pl["strip_rows"] <= 512
        assert pl["strip_rows"] % 8 == 0 and pl["strip_rows"] >= pl["n_tile"]
        assert pl["acc_stages"] == (2 if pl["nb"] * pl["strip_rows"] <= 256 else 1)
    u1, _ = plan("u1", 1024, 128, 32, "bf16")
    assert u1["mgroups"] == 2 and u1["nb"] == 4                   # two MMAs per weight tile for the weight-heaviest layer
    d4, _ = plan("d4", 1024, 128, 32, "bf16")
    assert d4["n_tile"] == 16 and d4["nb"] >= 4


def test_plan_options_and_errors():
    from phasegen import _lib
    single, _ = plan("u1", 512, 696, 256, "bf16x3", cta_pair=1)
    assert single["pair"] == 0 and single["strip_rows"] == 192    # full strip: 176 + 15 -> 192
    nomerge, _ = plan("d3", 1024, 128, 32, "bf16", max_clips_per_tile=1)
    assert nomerge["merged"] == 0 and nomerge["nb"] == 1
    tpg1, _ = plan("u1", 512, 696, 256, "bf16x3", taps_per_group=1)
    assert tpg1["groups0"] == 16
    # C_out not a multiple of 256: pairs are refused when required, dropped when automatic
    from phasegen import ops
    d = ops.conv_desc(0, 4, 64, 128, 64, 8, 1, 2, 64, 64, _lib.PG_PREC_BF16X3, cta_pair=2)
    out = (ctypes.c_int * 16)()
    assert _lib.load().pg_conv_tc_plan(ctypes.byref(d), out, 16) < 0 and "256" in _lib.last_error()
    d = ops.conv_desc(0, 4, 64, 128, 64, 8, 1, 2, 64, 64, _lib.PG_PREC_BF16X3)
    assert _lib.load().pg_conv_tc_plan(ctypes.byref(d), out, 16) == 0 and out[4] == 0
    assert _lib.load().pg_conv_tc_plan(ctypes.byref(d), out, 8) < 0


def test_whole_clip_tiles_for_the_fused_norm_epilogue():
    """BASELINE inference shape (C 512, T 696), bench precisions: which layers hold a whole clip per tile."""
    from phasegen import _lib, ops
    d2, _ = plan("d2", 512, 696, 256, "f16x3", whole_clip=1)          # 346 positions = 2 x 176 columns, one TMEM stage
    assert (d2["whole_clip"], d2["n_ntiles"], d2["OS"], d2["nb"], d2["acc_stages"]) == (1, 2, 1, 1, 1)
    u2, _ = plan("u2", 512, 696, 256, "f16x2", whole_clip=1)          # same column budget as its two-clip tiles
    assert (u2["whole_clip"], u2["n_ntiles"], u2["nb"], u2["acc_stages"]) == (1, 2, 1, 1)
    u3, _ = plan("u3", 512, 696, 256, "f16x3", whole_clip=1)          # both output phases side by side
    assert (u3["whole_clip"], u3["OS"], u3["n_ntiles"], u3["n_tile"], u3["acc_stages"]) == (1, 2, 1, 176, 1)
    _, du4 = plan("u4", 512, 696, 256, "f16x3")                       # short two-phase layer: stays merged, two-pass norm
    assert _lib.load().pg_conv_epilogue_supported(ctypes.byref(du4), _lib.PG_EPI_NORM_ACT) == 0
    d3, d = plan("d3", 512, 696, 256, "f16x3", whole_clip=1)          # one part: the ordinary tile already is a whole clip
    assert (d3["whole_clip"], d3["n_ntiles"], d3["OS"]) == (0, 1, 1)
    lib = _lib.load()
    assert lib.pg_conv_epilogue_supported(ctypes.byref(d), _lib.PG_EPI_NORM_ACT) == 1
    _, du1 = plan("u1", 512, 696, 256, "f16x2")
    assert lib.pg_conv_epilogue_supported(ctypes.byref(du1), _lib.PG_EPI_NORM_ACT) == 0      # 704 columns
    assert lib.pg_conv_epilogue_supported(ctypes.byref(du1), _lib.PG_EPI_ACT) == 1
    # training shape: merged single-phase tiles already hold whole clips; two-phase layers are un-merged on request
    d3t, _ = plan("d3", 1024, 128, 32, "bf16", whole_clip=1)
    assert d3t["merged"] == 1 and d3t["whole_clip"] == 0
    _, du3t = plan("u3", 1024, 128, 32, "bf16")
    assert lib.pg_conv_epilogue_supported(ctypes.byref(du3t), _lib.PG_EPI_NORM_ACT) == 0
