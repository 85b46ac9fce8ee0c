"""Forward throughput of the five BASELINE configurations on one GPU (inputs resident, CUDA-graph forward): clips/s and
achieved backbone TFLOP/s.  The headline bench (bench.py) is configuration 2; this table shows the others."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import TBNModel, baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.arch import backbone_flops  # noqa: E402
from attention_based_tbn_b200.synthetic import audio_frames, make_inputs, make_state_dict  # noqa: E402

CALIB = os.path.join(ROOT, "tests", "golden", "bn_calib.npz")
dev = torch.device("cuda", 0)
for idx, B, N in ((1, 64, 3), (2, 64, 3), (3, 64, 3), (4, 64, 3), (5, 8, 25)):
    cfg = baseline_config(idx)
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    model = TBNModel(cfg, mod, dev, precision="bf16", use_cuda_graph=True, chunk=32)
    model.load_state_dict(sd)
    model.eval()
    inp = {k: v.to(dev) for k, v in make_inputs(cfg, mod, B, N, seed=1).items()}
    with torch.no_grad():
        for _ in range(4):
            model(inp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 20
        for _ in range(iters):
            model(inp)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 0.0
    for m in mod:
        fl += backbone_flops(3, 224, 224) if m == "RGB" else backbone_flops(10, 224, 224) if m == "Flow" else backbone_flops(
            1, 256, audio_frames(cfg.data.audio.audio_length))
    print(f"cfg{idx} {'+'.join(mod):14s} B={B:3d} N={N:2d}: {ms:7.2f} ms/step  {B / ms * 1e3:8.0f} clips/s  "
          f"{fl * B * N / ms / 1e9:6.0f} TFLOP/s (backbone)", flush=True)
    del model
    torch.cuda.empty_cache()
