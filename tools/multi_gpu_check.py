#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-sharded render with the frame exchange
fused into the film resolve (PeerFrames), the photon pass sharded by emission ranges, and a reconstruction-filter
film summed over ranks - each against the reference's golden outputs / the single-GPU result.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import importlib, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m = importlib.import_module("monte-carlo-ray-tracer_b200")
mdist = importlib.import_module("monte-carlo-ray-tracer_b200.distributed")
res = {"world": world}

# 1. path tracer, rows interleaved, every rank's resolve stores into every rank's float3 frame over NVLink
cid = "c2_hexagon_room_96"
scene = m.Scene.from_pack(os.path.join(GOLDEN, cid + ".mcrtpack")); g = np.load(os.path.join(GOLDEN, cid + ".npz"))
cam = scene.cameras()[0]
pt = m.PathTracer(scene, device=local, precision=m.PRECISION_F64, global_seed=int(g["seed"]))
for f32 in (True, False):
    frames = mdist.PeerFrames(pt, rank, world, cam.height, cam.width, float32=f32, device=dev)
    frames.render(cam); frames.barrier()
    img = frames.tensor().cpu().numpy().astype(np.float64)
    err = float(np.abs(img - g["image"]).max() / max(1.0, np.abs(g["image"]).max()))
    res[f"peer_frames_{'f32' if f32 else 'f64'}_max_rel_err"] = err
    assert err < (1e-6 if f32 else 1e-9), err
    frames.close()

# 2. reconstruction filter summed over ranks
k = np.load(os.path.join(GOLDEN, "film_kat.npz")); films = json.loads(str(k["films"]))
fscene = m.Scene.from_pack(os.path.join(GOLDEN, "film_hexagon_room_64.mcrtpack"))
fpt = m.PathTracer(fscene, device=local, precision=m.PRECISION_F64, global_seed=int(k["seed"]))
for name in ("mitchell", "gaussian_cached", "lanczos_r3"):
    fcam = fscene.cameras()[0]; fcam.film = films[name]
    img = mdist.render_filtered(fpt, fcam, rank, world, dev).cpu().numpy()
    ref = k["image_" + name]
    err = float(np.abs(img - ref).max() / max(1.0, np.abs(ref).max()))
    res[f"filtered_{name}_max_rel_err"] = err
    assert err < 1e-9, (name, err)
fpt.close()

# 3. photon pass sharded by emission ranges + all-gather of the photons, then a photon-mapped sharded render
cid = "pm_hexagon_room_64"
pscene = m.Scene.from_pack(os.path.join(GOLDEN, cid + ".mcrtpack")); pg = np.load(os.path.join(GOLDEN, cid + ".npz"))
params = pscene.extra["photon_emit_params"]; ref_maps = pscene.photon_maps()
emit = dict(emissions=int(params[0]), caustic_factor=float(params[1]), max_photons_per_octree_leaf=int(params[2]),
            k_nearest_photons=ref_maps[2], direct_visualization=bool(ref_maps[3]), scene_bounds=params[3:9])
pm = m.PhotonMapper(pscene, device=local, global_seed=int(pg["seed"]))
nc, ng = pm.emit_sharded(rank, world, **emit)
res["photons"] = [int(nc), int(ng)]
for which in (0, 1):
    got = pm._maps[which]["photons"].reshape(-1, 8); ref = ref_maps[which]["photons"].reshape(-1, 8)
    assert len(got) == len(ref), (which, len(got), len(ref))
    gs = got[np.lexsort(got[:, [7, 6, 2, 1, 0, 5, 4, 3]].T[::-1])]; rs = ref[np.lexsort(ref[:, [7, 6, 2, 1, 0, 5, 4, 3]].T[::-1])]
    assert np.array_equal(gs[:, :6], rs[:, :6])
    for key in ("octant_start", "octant_count", "octant_next", "octant_leaf", "octant_bounds"):
        assert np.array_equal(pm._maps[which][key], ref_maps[which][key]), key
pcam = pscene.cameras()[0]
frames = mdist.PeerFrames(pm, rank, world, pcam.height, pcam.width, float32=False, device=dev)
frames.render(pcam); frames.barrier()
img = frames.tensor().cpu().numpy()
res["photon_mapped_image_rel_rmse"] = float(np.sqrt(np.mean((img - pg["image"]) ** 2)) / max(1.0, np.abs(pg["image"]).mean()))
assert res["photon_mapped_image_rel_rmse"] < 1e-6
frames.close(); pm.close(); pt.close()
if world > 1:
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK OK " + json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_{world}.json"), "w").write(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
