import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holoagent_b200.engine import HmsgEngine
from holoagent_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1040
e = HmsgEngine(0); e.encoder_load(synth.make_vit_weights())
x = torch.randn(B, 3, 224, 224, device="cuda"); torch.cuda.synchronize()
for _ in range(2): e.encode_images(x)
e.sync()
