"""torchrun --nproc-per-node N scripts/check_multigpu.py [--collective c|torch] [--frames F]: the frame-sharded ingest
(every rank holds only its contiguous frame block; geometry collectives + node-embedding merge) must give the same node
table and node features as the single-GPU job over all frames (fp32 sums differ only by summation order) and the same
3-D mask store for the frames each rank owns."""
import argparse, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth, ingest

ap = argparse.ArgumentParser()
ap.add_argument("--collective", default="c")
ap.add_argument("--frames", type=int, default=96)
ap.add_argument("--H", type=int, default=480)
ap.add_argument("--W", type=int, default=640)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = f"cuda:{local}"
F, H, W, M, FB = args.frames, args.H, args.W, 32, 16
eng = HmsgEngine(local)
eng.encoder_load(synth.make_vit_weights())
K = synth.intrinsics(H, W)


def load(ids):
    d, c, T, _ = synth.make_frames(ids, H, W, device=dev)
    eng.scene_begin(H, W, K, 1000.0, 0.05, max(len(ids), 1))
    eng.add_frames(d.view(torch.int16), c, torch.from_numpy(T.reshape(len(ids), 16)).to(dev)); eng.sync()
    return torch.from_numpy(np.stack([synth.make_mask_boxes(int(i), H, W, M) for i in ids])).to(dev)


# ---- sharded: this rank's block only
g0, cnt = ingest.frame_block(F, world, rank)
boxes = load(np.arange(g0, g0 + cnt))
job = ingest.IngestJob(eng, cnt, FB, M, 512, boxes, rank=rank, world=world, collective=args.collective, total_frames=F)
job.step_device(); eng.sync()
multi = job.full_feats.clone()
nodes_multi = eng.nodes_read()
mine = eng.mask_store_read(cnt - 1)                       # local frame id of the last frame of the block
# ---- single GPU over all frames (every rank repeats it)
boxes_all = load(np.arange(F))
single_job = ingest.IngestJob(eng, F, FB, M, 512, boxes_all, rank=0, world=1)
single_job.step_device(); eng.sync()
single = single_job.full_feats
nodes_single = eng.nodes_read()
ref = eng.mask_store_read(g0 + cnt - 1)
err = (multi - single).abs().max().item()
same_nodes = np.array_equal(nodes_multi[2], nodes_single[2]) and np.allclose(nodes_multi[0], nodes_single[0], rtol=1e-12, atol=1e-12)
same_masks = np.array_equal(mine[0], ref[0]) and np.array_equal(mine[3], ref[3]) and np.allclose(mine[1], ref[1], rtol=1e-9, atol=1e-9)
g = [torch.zeros_like(multi) for _ in range(world)]
dist.all_gather(g, multi)
same = all(torch.equal(g[0], x) for x in g)
ok = torch.tensor([int(same_nodes and same_masks)], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} collective={args.collective} nodes={multi.shape[0]} max|multi-single|={err:.3e} identical_across_ranks={same} "
          f"node_table_and_mask_store_equal={bool(ok.item())}")
    assert err < 1e-4 and same and ok.item() == 1
dist.barrier(); dist.destroy_process_group()
