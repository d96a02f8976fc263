import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gsdf_slam_b200 as gs
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = torch.Generator(device="cuda").manual_seed(P)
pts = torch.randn((P,3), device="cuda", generator=g) * torch.tensor([4.0,1.0,3.0], device="cuda")
gs.distCUDA2(pts); torch.cuda.synchronize()
