"""First-contact GPU probe: correctness spot checks + raw kernel timings (not the bench)."""
import json, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth

eng = HmsgEngine(0)
ts = eng.torch_stream()
res = {}

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    eng.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(reps): fn()
    e1.record(ts); e1.synchronize()
    return e0.elapsed_time(e1) / reps

what = sys.argv[1:] or ["gemm", "knn", "enc", "geom"]
if "gemm" in what:
  for two in (0, 1):
    eng.set_option("gemm_2sm", two)
    for (M, N, K) in [(52000, 2304, 768), (52000, 768, 768), (52000, 3072, 768), (52000, 768, 3072), (8192, 8192, 8192 // 8 * 8)]:
        A = (torch.randn(M, K, device="cuda") * 0.5).half(); W = (torch.randn(N, K, device="cuda") * 0.05).half()
        C = torch.zeros(M, N, device="cuda"); torch.cuda.synchronize()
        ms = timed(lambda: eng.gemm_debug(A, W, C, M, N, K))
        ref = (A[:256].float() @ W.float().T)
        err = (C[:256] - ref).abs().max().item() / ref.abs().max().item()
        res[f"gemm{two}_{M}x{N}x{K}"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "relerr": err}
        print(two, res[f"gemm{two}_{M}x{N}x{K}"], flush=True)
        del A, W, C
  eng.set_option("gemm_2sm", 0)
if "knn" in what:
    N, d = 1_000_000, 512
    E, Q = synth.make_knn_tables(N, 64, d, device="cuda")
    torch.cuda.synchronize()
    eng.index_set(E, borrow=True)
    for bq in (1, 2, 4, 8, 16):
        os.environ["HMSG_KNN_BQ"] = str(bq)
        eng.index_set(E, borrow=True)
        q = Q[:bq].contiguous(); torch.cuda.synchronize()
        ms = timed(lambda: eng.query_topk(q, 5), reps=10)
        res[f"knn_bq{bq}"] = {"ms_per_pass": ms, "GBs": N * d * 4 / ms / 1e6, "qps": bq / ms * 1e3}
        print(bq, res[f"knn_bq{bq}"], flush=True)
    del E, Q
if "enc" in what:
    sd = synth.make_vit_weights()
    eng.encoder_load(sd)
    for two in (0, 1):
      eng.set_option("gemm_2sm", two)
      for B in (65, 1040, 2080):
        x = torch.randn(B, 3, 224, 224, device="cuda"); torch.cuda.synchronize()
        ms = timed(lambda: eng.encode_images(x), reps=3, warm=1)
        res[f"enc{two}_B{B}"] = {"ms": ms, "img_per_s": B / ms * 1e3, "tflops": B * 8.82e9 / ms / 1e9}
        print(two, B, res[f"enc{two}_B{B}"], flush=True)
    eng.set_option("gemm_2sm", 0)
if "attn" in what:
    sd = synth.make_vit_weights()
    eng.encoder_load(sd)
    x = torch.randn(2080, 3, 224, 224, device="cuda"); torch.cuda.synchronize()
    for var in (0, 4, 3, 1):
        eng.set_option("attn_variant", var)
        eng.encode_images(x); eng.sync()
        eng.prof_enable("attn", "gemm", "eltwise")
        for c in ("attn", "gemm", "eltwise"): eng.prof_read(c)
        for _ in range(3): eng.encode_images(x)
        r = {c: eng.prof_read(c) for c in ("attn", "gemm", "eltwise")}
        eng.prof_enable()
        res[f"attn_var{var}"] = {"attn_ms_per_fwd": r["attn"]["ms"] / 3, "gemm_ms": r["gemm"]["ms"] / 3, "eltwise_ms": r["eltwise"]["ms"] / 3,
                                 "gemm_tflops": r["gemm"]["work"] / r["gemm"]["ms"] / 1e9}
        print("attn", var, res[f"attn_var{var}"], flush=True)
    eng.set_option("attn_variant", 0)
if "geom" in what:
    F, H, W = 200, 480, 640
    d, c, T, K = synth.make_frames(np.arange(F), H, W, device="cuda")
    eng.scene_begin(H, W, K, 1000.0, 0.05, F)
    eng.add_frames(d.view(torch.int16), c, torch.from_numpy(T).cuda().reshape(F, 16)); eng.sync()
    t = time.time(); nv, mb = eng.voxel_build(); res["voxel_build_s"] = time.time() - t; res["n_voxels"] = nv
    t = time.time(); nn = eng.radius_filter(1000, 1.0); res["radius_s"] = time.time() - t; res["n_nodes"] = nn
    print(res, flush=True)
    eng.features_begin(512)
    M = 32
    boxes = torch.from_numpy(np.stack([synth.make_mask_boxes(i, H, W, M) for i in range(F)])).cuda()
    feats = torch.nn.functional.normalize(torch.randn(64, 2 * M + 1, 512, device="cuda"), dim=-1); torch.cuda.synchronize()
    def step():
        for b0 in range(0, 192, 64):
            eng.masks_boxes(b0, boxes[b0:b0 + 64])
            eng.fuse_scatter(b0, 64, M, feats, 0.4418)
    ms = timed(step, reps=3, warm=1)
    res["fuse_scatter_ms_per_frame"] = ms / 192
    print(res, flush=True)
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
