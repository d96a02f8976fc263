"""2+ GPU check (torchrun): the fused peer-memory exchange (gsr_adam_step_peer) against the NCCL
all-reduce + fused Adam path, same keyframes, 3 steps.  Prints max differences; replicas must be
bit-identical across ranks."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import scenes
from gsdf_slam_b200.mapping import SEGMENTS, Keyframe, KeyframeBatchMapper

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
scene = scenes.room_scene(60_000, 640, 360, 320.0, 320.0, seed=2, name="small_room")
cams = scenes.ring_cameras(8, 640, 360, 320.0, 320.0, radius=0.3)
cam = cams[rank % 8]
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
base = dict(means3D=t(scene.means3D), dc=t(scene.dc), sh=t(scene.sh), opacity=t(scene.opacity),
            scales=t(scene.scales), rotations=t(scene.rotations))
kf = Keyframe(t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), cam.tanfovx, cam.tanfovy, cam.H, cam.W,
              t(scene.dL_dout))
cam2 = cams[(rank + 3) % 8]
kf2 = Keyframe(t(cam2.viewmatrix), t(cam2.projmatrix), t(cam2.campos), cam2.tanfovx, cam2.tanfovy, cam2.H, cam2.W,
               t(scene.dL_dout))
worst = 0.0
# one keyframe per rank (the backward skips the zero fill, the exchange reads only rendering ranks) and two
# keyframes per rank (dense accumulation)
for kfs in ([kf], [kf, kf2]):
  res = {}
  for peer in (False, True):
    m = KeyframeBatchMapper({k: v.clone() for k, v in base.items()}, scene.degree, t(scene.bg),
                            freeze_render_inputs=True, peer_exchange=peer)
    assert (m.peer is not None) == peer, "peer path not active"
    if peer:
        m.arena.fill_(float("nan"))   # rows nobody renders must never be read
    for _ in range(3):
        m.step(kfs)
    torch.cuda.synchronize()
    res[peer] = {k: m.opt_params[k].clone() for k in SEGMENTS}
    del m
  for k in SEGMENTS:
    a, b = res[False][k].double(), res[True][k].double()
    d0 = (res[False][k] - base[k]).double()
    rel = float((a - b).norm() / max(float(d0.norm()), 1e-30))
    worst = max(worst, rel)
    # replicas identical across ranks
    g = [torch.empty_like(res[True][k]) for _ in range(world)]
    dist.all_gather(g, res[True][k].contiguous())
    same = all(torch.equal(g[0].view(torch.int32), x.view(torch.int32)) for x in g)
    if rank == 0:
        print(f"K={len(kfs)} {k:10s} rel-L2(update) peer vs nccl {rel:.3e}  replicas bit-identical: {same}")
    assert same, k
    assert bool(torch.isfinite(res[True][k]).all()), k
assert worst <= 2e-3, worst
if rank == 0:
    print("peer exchange OK")
dist.destroy_process_group()
