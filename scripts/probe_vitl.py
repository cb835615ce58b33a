"""ViT-L/14 (the reference's default tower) throughput probe: images/s and model TFLOP/s per chunk."""
import dataclasses, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth

eng = HmsgEngine(0)
sh = synth.VIT_L14
eng.encoder_load(synth.make_vit_weights(sh, seed=3), **dataclasses.asdict(sh))
T = (sh.image // sh.patch) ** 2 + 1
W = sh.width
# 2 flop/MAC: patch embed + layers * (qkv + attention + out-proj + mlp) + projection
flop = 2 * ((T - 1) * 3 * sh.patch ** 2 * W + sh.layers * (T * W * 3 * W + 2 * sh.heads * T * T * 64 + T * W * W + 2 * T * W * sh.mlp) + W * sh.out_dim)
res = {"gflop_per_image": flop / 1e9}
for B in (64, 520):
    x = torch.randn(B, 3, 224, 224, device="cuda"); torch.cuda.synchronize()
    eng.encode_images(x); eng.sync()
    eng.prof_enable("attn", "gemm", "eltwise")
    for c in ("attn", "gemm", "eltwise"): eng.prof_read(c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.encode_images(x)
    e1.record(); torch.cuda.synchronize()
    r = {c: eng.prof_read(c) for c in ("attn", "gemm", "eltwise")}
    eng.prof_enable()
    ms = e0.elapsed_time(e1) / 3
    res[f"B{B}"] = {"ms": ms, "img_per_s": B / ms * 1e3, "model_tflops": B * flop / ms / 1e9, "attn_ms": r["attn"]["ms"] / 3, "gemm_ms": r["gemm"]["ms"] / 3,
                    "gemm_tflops": r["gemm"]["work"] / r["gemm"]["ms"] / 1e9, "eltwise_ms": r["eltwise"]["ms"] / 3}
    print(B, res[f"B{B}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_vitl.json", "w"), indent=1)
