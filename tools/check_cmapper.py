"""2+ GPU check (torchrun): the C mapper (gsr_mapper_*, CUDA-IPC arenas + on-device barrier + fused peer exchange) against
the Python mapper's NCCL all-reduce path and against one rank accumulating all keyframes sequentially.  Prints one JSON
line on rank 0; exits non-zero when a bar is missed."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import scenes
from gsdf_slam_b200.cmapper import CMapper
from gsdf_slam_b200.mapping import SEGMENTS, Keyframe, KeyframeBatchMapper, partition_keyframes

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
scene = scenes.room_scene(80_000, 640, 368, 320.0, 320.0, seed=15, box=(16.0, 3.0, 12.0), name="cmapper2")
cams = scenes.ring_cameras(8, 640, 368, 320.0, 320.0, radius=0.5)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
base = dict(means3D=t(scene.means3D), dc=t(scene.dc), sh=t(scene.sh), opacity=t(scene.opacity),
            scales=t(scene.scales), rotations=t(scene.rotations))
dL = t(scene.dL_dout)
kfs = [Keyframe(t(c.viewmatrix), t(c.projmatrix), t(c.campos), c.tanfovx, c.tanfovy, c.H, c.W, dL) for c in cams]
local = [kfs[i] for i in partition_keyframes(len(kfs), world, rank)]
bg = t(scene.bg)
P = scene.P
steps = 3
cm = CMapper(base, scene.degree, bg, rank=rank, world=world)
st_n = {k: torch.zeros(P, device=dev) for k in ("max_radii2D", "xyz_gradient_accum", "denom")}
nc = KeyframeBatchMapper({k: v.clone() for k, v in base.items()}, scene.degree, bg, stats=st_n, peer_exchange=False)
for _ in range(steps):
    cm.step(local)
    nc.step(local)
torch.cuda.synchronize()
got, gst = cm.params(), cm.stats()
out = {"world": world, "steps": steps, "P": P}
worst = 0.0
for k in SEGMENTS:
    d0 = (nc.opt_params[k] - base[k]).double()
    worst = max(worst, float((got[k].double() - nc.opt_params[k].double()).norm() / max(float(d0.norm()), 1e-30)))
out["c_mapper_vs_nccl_rel_l2"] = float(f"{worst:.3e}")
dig = torch.cat([got[k].contiguous().view(torch.int32).long().sum().view(1) for k in SEGMENTS] +
                [gst[k].contiguous().view(torch.int32).long().sum().view(1) for k in gst])
allg = [torch.zeros_like(dig) for _ in range(world)]
dist.all_gather(allg, dig)
out["replicas_bit_identical"] = bool(all(torch.equal(allg[0], x) for x in allg))
out["denom_equal_nccl"] = bool(torch.equal(gst["denom"], st_n["denom"]))
out["max_radii_equal_nccl"] = bool(torch.equal(gst["max_radii2D"], st_n["max_radii2D"]))
if rank == 0:
    st_s = {k: torch.zeros(P, device=dev) for k in ("max_radii2D", "xyz_gradient_accum", "denom")}
    sq = KeyframeBatchMapper({k: v.clone() for k, v in base.items()}, scene.degree, bg, stats=st_s, single_rank=True)
    for _ in range(steps):
        sq.step(kfs)
    torch.cuda.synchronize()
    w2 = 0.0
    for k in SEGMENTS:
        d0 = (sq.opt_params[k] - base[k]).double()
        w2 = max(w2, float((got[k].double() - sq.opt_params[k].double()).norm() / max(float(d0.norm()), 1e-30)))
    out["c_mapper_vs_sequential_rel_l2"] = float(f"{w2:.3e}")
    out["denom_equal_sequential"] = bool(torch.equal(gst["denom"], st_s["denom"]))
    num = float((gst["xyz_gradient_accum"].double() - st_s["xyz_gradient_accum"].double()).norm())
    out["accum_rel_l2_sequential"] = float(f"{num / max(float(st_s['xyz_gradient_accum'].double().norm()), 1e-30):.3e}")
dist.barrier()
cm.close()
ok = out["replicas_bit_identical"] and out["c_mapper_vs_nccl_rel_l2"] <= 1e-4 and out["denom_equal_nccl"]
if rank == 0:
    ok = ok and out["c_mapper_vs_sequential_rel_l2"] <= 1e-4 and out["denom_equal_sequential"]
    print(json.dumps(out))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
