"""Device memory per frame of batch capacity of each handle (cudaMemGetInfo deltas)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pl_vi_orbslam3_b200 import ORBextractor, Lineextractor
from pl_vi_orbslam3_b200.frontend import FrontEnd
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.cuda.init(); torch.zeros(1, device="cuda")
def used():
    torch.cuda.synchronize(); f, t = torch.cuda.mem_get_info(); return t - f
u0 = used(); o = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B); u1 = used()
l = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=B); u2 = used()
print(f"B={B}: ORB handle {(u1-u0)/B/1e6:.2f} MB/frame, line handle {(u2-u1)/B/1e6:.2f} MB/frame")
o.close(); l.close()
u0 = used(); fe = FrontEnd(B, pairs=True, out_sets=2); u1 = used()
print(f"FrontEnd(pairs, 2 output sets) {(u1-u0)/B/1e6:.2f} MB/frame = {(u1-u0)/1e9:.2f} GB at B={B}")
fe.close()
