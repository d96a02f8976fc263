"""2+ GPU check (torchrun): the NVLink byte counters of gsr_adam_step_peer against a direct count — the number of
parameter elements whose bits changed in one step, per group, and the rows each rank rendered."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import scenes
from gsdf_slam_b200.mapping import SEGMENTS, Keyframe, KeyframeBatchMapper

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P = int(os.environ.get("CT_P", "120000"))
scene = scenes.room_scene(P, 640, 360, 320.0, 320.0, seed=2, name="small_room")
cams = scenes.ring_cameras(8, 640, 360, 320.0, 320.0, radius=0.3)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
base = dict(means3D=t(scene.means3D), dc=t(scene.dc), sh=t(scene.sh), opacity=t(scene.opacity),
            scales=t(scene.scales), rotations=t(scene.rotations))
kfs = [Keyframe(t(c.viewmatrix), t(c.projmatrix), t(c.campos), c.tanfovx, c.tanfovy, c.H, c.W, t(scene.dL_dout))
       for c in cams[rank::world]]
stats = ({k: torch.zeros(P, device=dev) for k in ("max_radii2D", "xyz_gradient_accum", "denom")}
         if os.environ.get("CT_STATS", "0") == "1" else None)
m = KeyframeBatchMapper({k: v.clone() for k, v in base.items()}, scene.degree, t(scene.bg), freeze_render_inputs=True,
                        stats=stats)
assert m.peer is not None
for _ in range(int(os.environ.get("CT_STEPS", "6"))):
    m.step(kfs)
torch.cuda.synchronize()
before = {k: m.opt_params[k].clone() for k in SEGMENTS}
tr = m.measure_exchange_traffic(lambda: m.step(kfs))
after = {k: m.opt_params[k] for k in SEGMENTS}
tot = chg = 0
per = {}
for k in SEGMENTS:
    d = (before[k].view(torch.int32) != after[k].view(torch.int32))
    per[k] = round(float(d.float().mean()), 4)
    tot += d.numel(); chg += int(d.sum())
rendered = float((m.last_radii > 0).float().mean())
if rank == 0:
    print({"changed_fraction_by_group": per, "changed_elements_bytes_all_ranks": 4 * chg, "arena_bytes": 4 * tot,
           "counted_param_changed_bytes_this_rank": tr["param_changed_bytes"], "counted_grad_in_bytes_this_rank": tr["grad_in_bytes"],
           "rendered_fraction_last_keyframe": round(rendered, 4), "world": world})
dist.destroy_process_group()
