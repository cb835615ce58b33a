import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth
e = HmsgEngine(0)
E, Q = synth.make_knn_tables(1_000_000, 8, 512, device="cuda"); torch.cuda.synchronize()
e.index_set(E, borrow=True)
for _ in range(3): e.query_topk(Q, 5)
e.sync()
