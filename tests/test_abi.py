"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/phasegen.h declares; the drop-in modules keep the reference's call surface; the product
path refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "phasegen.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from phasegen import _lib
    lib = ctypes.CDLL(os.path.normpath(_lib.LIB_PATH))
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/phasegen.h but not exported"
    assert set(_lib.EXPORTS) == set(syms), "ctypes binding and header disagree"
    assert _lib.load().pg_abi_version() == _lib.ABI_VERSION == 4


def test_struct_layouts_match_header():
    from phasegen import _lib
    assert ctypes.sizeof(_lib.ConvDesc) == 21 * 4
    assert ctypes.sizeof(_lib.ActDst) == 8 + 8 + 8 + 4 * 4 + 8          # + range_flag pointer
    assert ctypes.sizeof(_lib.ConvEpilogue) == 8 + 8 + 8 + 8 + 2 * 48 + 8   # mode(+pad), gamma, beta, eps(+pad), dst0, dst1, scale_shift


def test_error_reporting_without_gpu_call():
    from phasegen import _lib
    lib = _lib.load()
    d = _lib.ConvDesc(0, 1, 64, 128, 32, 999, 8, 1, 2, 32, 64, 999, 128, 1, 0, 0, 0, 0, 0, 0)
    assert lib.pg_conv_stat_parts(ctypes.byref(d)) < 0          # L_out inconsistent with geometry
    assert "L_out" in _lib.last_error()
    d.L_out = 29
    assert lib.pg_conv_stat_parts(ctypes.byref(d)) == 1
    assert lib.pg_stft_num_frames(177920, 256) == 696


def test_model_call_surface_and_checkpoint_keys(golden_dir, tmp_path):
    import model
    z = np.load(os.path.join(golden_dir, "unet_c8_t24.npz"))
    ref = {k[4:]: z[k] for k in z.files if k.startswith("sd::")}
    net = model.UNetModel(8, 16)                                 # default norm_layer, like train.py:15
    sd = net.model.state_dict()
    assert list(sd) == list(ref) or set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert hasattr(net, "gpu_ids") and isinstance(net.model, model.UNetBlock)
    # a reference checkpoint (inner-block state_dict, model.py:45-48) loads; save round-trips
    path = str(tmp_path / "ckpt")
    torch.save({k: torch.from_numpy(v) for k, v in ref.items()}, path)
    net.load(path)
    assert torch.equal(net.model.state_dict()["model.0.weight"], torch.from_numpy(ref["model.0.weight"]).float())
    net.save(path)
    again = torch.load(path)
    assert set(again) == set(ref)
    n_params = sum(p.numel() for p in model.UNetModel(1024, 2048).parameters())
    assert n_params == 612_392_960


def test_no_cpu_fallback():
    import model
    import utils
    net = model.UNetModel(8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.forward(torch.zeros(1, 8, 24))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            utils.generate_audio(np.zeros((256, 24), np.complex64), 16000, 128, is_stft=True)


def test_reference_utility_names_importable():
    import utils
    for name in ("Pool", "GANLoss", "View", "EnergyLoss", "Transpose", "Flatten", "generate_spec_img",
                 "generate_audio", "griffin_lim"):
        assert hasattr(utils, name)
    assert utils.Flatten()(torch.zeros(2, 3, 4)).shape == (2, 12)
    assert utils.Transpose(1, 2)(torch.zeros(2, 3, 4)).shape == (2, 4, 3)
    a = torch.randn(2, 2, 5)
    assert float(utils.EnergyLoss()(a, a)) == 0.0
    assert abs(float(utils.GANLoss()(torch.ones(3), True))) < 1e-12


def test_synthetic_shapes():
    from phasegen import synth
    assert synth.frames_for(4.0, 44100, 256) == 696
    assert synth.clip_samples(4.0, 44100, 256) == 177920
    w = synth.synthetic_waves(2, 1000, seed=3)
    assert w.shape == (2, 1000) and float(w.abs().max()) <= 1.0
    assert torch.equal(w, synth.synthetic_waves(2, 1000, seed=3))


def test_logger_drop_in_on_stock_tensorboard(tmp_path):
    """logger.Logger (reference logger.py:6-49) without tensorboardX: scalars / audio / images are accepted and
    write() exports log.json."""
    import json
    from collections import OrderedDict
    import logger
    lg = logger.Logger(str(tmp_path / "run"))
    lg.log(1, OrderedDict([("MSE", 0.5), ("NOPMSE", np.float32(0.25))]))
    lg.log(2, OrderedDict([("MSE", 0.4)]), text=True)
    lg.log(2, OrderedDict([("wav", np.zeros(800, np.float32))]), log_type="audio", sr=8000)
    lg.log(2, OrderedDict([("img", np.zeros((20, 30, 3), np.uint8))]), log_type="image")
    with pytest.raises(ValueError):
        lg.log(3, {}, log_type="video")
    with pytest.raises(ValueError):
        lg.log(3, {"a": np.zeros(4)}, log_type="audio")
    lg.write(); lg.flush(); lg.close()
    data = json.load(open(tmp_path / "run" / "log.json"))
    key = [k for k in data if k.endswith("scalar/MSE")][0]
    assert [row[1:] for row in data[key]] == [[1, 0.5], [2, 0.4]]
