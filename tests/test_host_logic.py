"""CPU: host-side logic of the product (no kernels are launched here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from cases import CALIB, GOLDEN
from attention_based_tbn_b200 import TBNModel, baseline_config, build_model, default_cfg, get_modality
from attention_based_tbn_b200 import arch, sampling
from attention_based_tbn_b200.synthetic import make_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arch_table_flops_match_survey():
    assert len(list(arch.all_conv_specs(3))) == 69
    assert abs(arch.backbone_flops(3, 224, 224) / 1e9 - 4.0632) < 1e-3
    assert abs(arch.backbone_flops(10, 224, 224) / 1e9 - 4.6139) < 1e-3
    assert abs(arch.backbone_flops(1, 256, 421) / 1e9 - 8.3070) < 1e-3
    assert abs(arch.backbone_flops(1, 256, 256) / 1e9 - 5.1015) < 1e-3
    assert arch.spatial_trace(256, 421)["5b"] == (8, 13)
    assert arch.spatial_trace(224, 224)["5b"] == (7, 7)


def test_segment_indices_bit_exact_vs_reference_vectors():
    from cases import sampling_golden

    z = sampling_golden()
    assert len(z["rows"]) == 28472  # all rows of the reference's annotations/epic_train_val.csv
    mods = ["RGB", "Flow", "Audio"]
    for mode, nseg in (("test", 25), ("val", 3), ("train", 3)):
        for samp in ("sync", "async"):
            np.random.seed(0)  # same global-RNG call order as the reference (main.py:20)
            for i, (s, e) in enumerate(zip(z["start_frame"], z["stop_frame"])):
                rec = sampling.SegmentRecord(int(s), int(e))
                idx = sampling.sample_indices(rec, mods, nseg, "train" if mode == "train" else "test", samp, 5)
                for m in mods:
                    ref = z[f"{mode}{nseg}_{samp}_{m}"][i]
                    assert idx[m].dtype == np.int64
                    assert np.array_equal(idx[m], ref), (mode, samp, m, i)


def test_audio_window_start_matches_reference_function():
    """tests/golden/audio_window.npz: the reference's own ``_get_audio_segment`` (dataset.py:421-459) run on a ramp
    signal, so the window it returned names the start index it chose (oracle/make_golden.py::make_audio_windows).
    Includes clips shorter than the window (zero padded; the reference then slices with a NEGATIVE start)."""
    z = np.load(f"{GOLDEN}/audio_window.npz")
    fps, sr = int(z["vid_fps"]), int(z["sampling_rate"])
    short = 0
    for audio_length, frame_idx, n, win_len, first, last in z["cases"]:
        frame_idx, n, win_len = int(frame_idx), int(n), int(win_len)
        min_len = int(audio_length * sr)
        start = sampling.audio_window_start(frame_idx, n, fps, float(audio_length), sr)
        ramp = np.arange(n, dtype=np.float64)
        if n < min_len:
            ramp = np.pad(ramp, (0, min_len - n))
            short += 1
        win = ramp[start:start + min_len]  # same slicing as the reference, negative start included
        assert len(win) == win_len, (audio_length, frame_idx, n, start)
        if win_len:
            assert win[0] == first and win[-1] == last, (audio_length, frame_idx, n, start)
    assert short > 50 and len(z["cases"]) == 800


def test_sampling_edge_cases():
    rec = sampling.SegmentRecord(10, 14)  # 3 frames < 25 segments -> all indices = start
    idx = sampling.segment_offsets(rec, "RGB", 25, "test")
    assert idx.shape == (25,) and (idx == 9).all()
    fl = sampling.flow_frame_indices(np.array([3, 10], np.int64), 5, 2)
    assert fl.tolist() == [3, 4, 5, 6, 7, 10, 11, 12, 13, 14]
    assert sampling.audio_window_start(0, 24000 * 10) == 0
    assert sampling.audio_window_start(600, 24000 * 10) == int((600 / 60 - 1.05) * 24000) - (
        int((600 / 60 - 1.05) * 24000) + 50400 - 240000)
    u = sampling.attention_prior("uniform", 2.1)
    assert u.shape == (13, 1) and u.dtype == np.float32 and np.allclose(u, 1 / 13)
    g = sampling.attention_prior("gaussian", 2.1)
    assert abs(g.sum() - 1) < 1e-6 and g.argmax() == 6
    assert [len(sampling.shard_clips(10, 4, r)) for r in range(4)] == [3, 3, 2, 2]
    assert list(sampling.shard_clips(10, 4, 1)) == [3, 4, 5]


def test_state_dict_names_match_reference_contract():
    cfg = baseline_config(5)
    mod = get_modality(cfg)
    model = TBNModel(cfg, mod, "cpu")
    own = model.state_dict()
    assert len(own) == 1464  # probed on the reference
    for k, shape in (("Base_RGB.conv1_7x7_s2.weight", (64, 3, 7, 7)), ("Base_Flow.conv1_7x7_s2.weight", (64, 10, 7, 7)),
                     ("Base_Audio.conv1_7x7_s2.weight", (64, 1, 7, 7)), ("pe.0.pe", (1, 10, 13)),
                     ("pe.1.weight", (1024, 1034, 1)), ("pe.2.weight", (1024,)),
                     ("attention_layer.attention_layer.in_proj_weight", (3072, 1024)),
                     ("attention_layer.attention_layer.out_proj.bias", (1024,)),
                     ("fusion.fusion_layer.0.weight", (512, 3072)), ("classifier.verb.weight", (125, 512)),
                     ("classifier.noun.bias", (352,)), ("Base_RGB.inception_5b_pool_proj_bn.running_var", (128,))):
        assert tuple(own[k].shape) == shape, k
    model.load_state_dict(make_state_dict(cfg, mod, 0, calib=CALIB), strict=True)
    # partial-BN freeze (reference model.py:164-176): only the first BN keeps trainable affine
    assert model.Base_RGB.conv1_7x7_s2_bn.weight.requires_grad
    assert not model.Base_RGB.conv2_3x3_bn.weight.requires_grad


def test_model_contract_errors():
    cfg = baseline_config(2)
    mod = get_modality(cfg)
    model, criterion, _ = build_model(cfg, mod, torch.device("cpu"))
    assert "crossentropy" in criterion
    x = {"RGB": torch.zeros(1, 3, 3, 32, 32), "Audio": torch.zeros(1, 3, 1, 32, 32)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x)  # training mode: the training forward runs on the device as well
    model.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x)
    bad = default_cfg(**{"model.arch": "resnet"})
    with pytest.raises(AssertionError):
        build_model(bad, mod, torch.device("cpu"))
    nope = default_cfg(**{"model.attention.use_pe": False, "data.flow.enable": False})
    m2 = TBNModel(nope, ["RGB", "Audio"], "cpu")
    assert not hasattr(m2, "pe")  # reference quirk: mha without use_pe has no pe module


def test_pretrained_loading_follows_reference_factory(tmp_path):
    """bn_inception.py:58-101 on synthesised Cadene-style files: RGB strict, Flow from the 10-channel Kinetics file,
    Audio = ImageNet file with conv1 averaged over RGB and keys reconciled; a missing file raises like torch.load."""
    from attention_based_tbn_b200.model import BNInceptionParams

    g = torch.Generator().manual_seed(5)
    files = {}
    for fname, cin, ncls in (("imagenet_bninception_rgb.pth", 3, 1000), ("kinetics_bninception_flow.pth", 10, 400)):
        sd = {k: torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone()
              for k, v in BNInceptionParams(cin).state_dict().items()}
        sd["last_linear.weight"], sd["last_linear.bias"] = torch.randn(ncls, 1024, generator=g), torch.randn(ncls, generator=g)
        torch.save(sd, tmp_path / fname)
        files[fname] = sd
    cfg = baseline_config(5)
    cfg.model["weights_dir"] = str(tmp_path)
    model = TBNModel(cfg, get_modality(cfg), "cpu")
    img, kin = files["imagenet_bninception_rgb.pth"], files["kinetics_bninception_flow.pth"]
    assert torch.equal(model.Base_RGB.conv1_7x7_s2.weight, img["conv1_7x7_s2.weight"])
    assert torch.equal(model.Base_Flow.conv1_7x7_s2.weight, kin["conv1_7x7_s2.weight"])
    assert torch.equal(model.Base_Audio.conv1_7x7_s2.weight, img["conv1_7x7_s2.weight"].mean(dim=1).unsqueeze(1))
    assert torch.equal(model.Base_Audio.inception_5b_pool_proj.weight, img["inception_5b_pool_proj.weight"])
    assert torch.equal(model.Base_Flow.inception_4c_3x3_bn.running_var, kin["inception_4c_3x3_bn.running_var"])
    assert not hasattr(model.Base_RGB, "last_linear")
    os.remove(tmp_path / "kinetics_bninception_flow.pth")
    with pytest.raises(FileNotFoundError):
        TBNModel(cfg, get_modality(cfg), "cpu")
    bad = dict(img)
    del bad["inception_3a_1x1.weight"]
    torch.save(bad, tmp_path / "imagenet_bninception_rgb.pth")
    cfg2 = baseline_config(2)
    cfg2.model["weights_dir"] = str(tmp_path)
    with pytest.raises(RuntimeError, match="inception_3a_1x1.weight"):
        TBNModel(cfg2, ["RGB"], "cpu")  # strict for RGB, as in the reference


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "tbn_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:int|const char\*)\s+(tbn_\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        protos[m.group(1)] = args
    return protos


def test_c_abi_exports_every_declared_symbol(lib):
    protos = _header_prototypes()
    assert len(protos) >= 17
    handle = lib.load()
    for name, args in protos.items():
        assert hasattr(handle, name), f"{name} declared in include/tbn_sm100.h but not exported"
        if name == "tbn_last_error":
            continue
        sig = lib.SIGNATURES[name]
        assert len(sig) == len(args), (name, len(sig), len(args))
        for ct, decl in zip(sig, args):
            if decl.startswith("const char*"):
                assert ct is ctypes.c_char_p, (name, decl)
            elif "*" in decl:
                assert ct is ctypes.c_void_p, (name, decl)
            elif decl.startswith("long long"):
                assert ct is ctypes.c_longlong, (name, decl)
            elif decl.startswith("float"):
                assert ct is ctypes.c_float, (name, decl)
            else:
                assert ct is ctypes.c_int, (name, decl)
    assert handle.tbn_version() >= 100
    assert set(lib.SIGNATURES) | {"tbn_last_error"} == set(protos)


def test_argument_errors_do_not_throw_across_the_abi(lib):
    # argument validation happens before any CUDA call, so this runs without a GPU
    with pytest.raises(lib.TbnKernelError, match="k must be 1 or 3"):
        lib.call("tbn_conv_tc", 16, 16, 16, 16, 1, 8, 8, 64, 64, 0, 64, 64, 0, 5, 1, 1, 1, 0)
    with pytest.raises(lib.TbnKernelError, match="multiples of 8"):
        lib.call("tbn_pool3s1", 16, 16, 1, 8, 8, 12, 12, 0, 12, 0, 0, 0, 0)
    assert "multiples of 8" in lib.last_error()


def test_training_gradient_names_are_model_parameters():
    """The names the training step returns gradients for (head_train.head_parameter_names + TrainBackbone.parameter_names)
    are exactly parameters of the model, for every attention variant; with freeze_mode "partialbn" the trainable set is
    all convs + the first BatchNorm of each backbone + the heads, as the reference leaves it (model.py:164-176)."""
    from attention_based_tbn_b200 import TBNModel
    from attention_based_tbn_b200.backbone_train import TrainBackbone
    from attention_based_tbn_b200.config import baseline_config, default_cfg, get_modality
    from attention_based_tbn_b200.head_train import head_parameter_names

    cfgs = [baseline_config(i) for i in (1, 2, 3, 4, 5)] + [default_cfg(**{"data.flow.enable": False, "model.attention.type": t})
                                                            for t in ("unimodal", "proto")]
    for cfg in cfgs:
        mod = get_modality(cfg)
        model = TBNModel(cfg, mod, "cpu", precision="fp32")
        params = dict(model.named_parameters())
        names = list(head_parameter_names(cfg, mod))
        for m in mod:
            names += TrainBackbone.parameter_names(_bb_stub(m))
        assert len(set(names)) == len(names)
        missing = [n for n in names if n not in params]
        assert not missing, missing[:5]
        trainable = {n for n, p in params.items() if p.requires_grad}
        unused_pe = {n for n in trainable if n.startswith("pe.")} if cfg.model.attention.type != "mha" else set()
        assert trainable - unused_pe <= set(names)
        for m in mod:
            bn = [n for n in trainable if n.startswith(f"Base_{m}.") and "_bn." in n]
            assert sorted(bn) == [f"Base_{m}.conv1_7x7_s2_bn.bias", f"Base_{m}.conv1_7x7_s2_bn.weight"], (m, bn)


def _bb_stub(m):
    from attention_based_tbn_b200.arch import block_specs, stem_specs
    from attention_based_tbn_b200.backbone_train import TrainBackbone

    bb = TrainBackbone.__new__(TrainBackbone)
    bb.prefix, bb.stem, bb.blocks = f"Base_{m}.", stem_specs({"RGB": 3, "Flow": 10, "Audio": 1}[m]), block_specs()
    return bb
