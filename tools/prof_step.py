import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda',0)
hp = bench.HotPath(B, dev, 1)
x, mus, lvs = bench.synth_inputs(B, 1000, dev)
hp.load(x, mus, lvs)
for _ in range(3): hp.step()
torch.cuda.synchronize()
