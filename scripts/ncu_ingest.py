"""Small ingest job (64 frames 640x480, M=32) for ncu captures of the non-GEMM kernels."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth, ingest
F, H, W, M = 64, 480, 640, 32
eng = HmsgEngine(0)
d, c, T, K = synth.make_frames(np.arange(F), H, W, device="cuda")
eng.scene_begin(H, W, K, 1000.0, 0.05, F)
eng.add_frames(d.view(torch.int16), c, torch.from_numpy(T.reshape(F, 16)).cuda()); eng.sync()
boxes = torch.from_numpy(np.stack([synth.make_mask_boxes(i, H, W, M) for i in range(F)])).cuda()
eng.encoder_load(synth.make_vit_weights())
job = ingest.IngestJob(eng, F, 32, M, 512, boxes)
for _ in range(2):
    job.step_device()
eng.sync()
print("ok", eng.n_nodes)
