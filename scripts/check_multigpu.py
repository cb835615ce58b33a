"""torchrun --nproc-per-node N scripts/check_multigpu.py : sharded ingest + one all-gather must give
the same node features as the single-GPU job (fp32 sums differ only by summation order)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth, ingest

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F, H, W, M, FB = 96, 480, 640, 32, 16
eng = HmsgEngine(local)
d, c, T, K = synth.make_frames(np.arange(F), H, W, device=f"cuda:{local}")
eng.scene_begin(H, W, K, 1000.0, 0.05, F)
eng.add_frames(d.view(torch.int16), c, torch.from_numpy(T.reshape(F, 16)).cuda()); eng.sync()
boxes = torch.from_numpy(np.stack([synth.make_mask_boxes(i, H, W, M) for i in range(F)])).cuda()
eng.encoder_load(synth.make_vit_weights())
job = ingest.IngestJob(eng, F, FB, M, 512, boxes, rank=rank, world=world)
job.step_device(); eng.sync()
multi = job.full_feats.clone()
single_job = ingest.IngestJob(eng, F, FB, M, 512, boxes, rank=0, world=1)
single_job.step_device(); eng.sync()
single = single_job.full_feats
err = (multi - single).abs().max().item()
g = [torch.zeros_like(multi) for _ in range(world)]
dist.all_gather(g, multi)
same = all(torch.equal(g[0], x) for x in g)
if rank == 0:
    print(f"world={world} nodes={multi.shape[0]} max|multi-single|={err:.3e} identical_across_ranks={same}")
    assert err < 1e-4 and same
dist.barrier(); dist.destroy_process_group()
