"""GPU tool: one clip sharded by frame range across the ranks of a torchrun job.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_demo.py cfg3_2160p_rgb32 120 30

Every rank encodes its GOP-aligned range on its own GPU; the mvs[] blob travels rank to rank around the resolve
(pipelined) -- CPU tensors over gloo, the data path itself has no collective.  Rank 0 concatenates the bitstreams and
checks them byte for byte against its own single-GPU encode of the whole clip, then decodes the ranges back."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from screenpressor_b200 import shard, synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_2160p_rgb32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 120
interval = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
cfg = synth.CONFIGS[name]
clip = synth.make_clip(cfg, n)
keys = synth.keyframe_flags(n, interval)
ranges = shard.assign_ranges(keys, world)
mine = next((r for r in ranges if r.rank == rank), None)
codec = ScreenCodec(local); codec.Init(CodecParameters(cfg.width, cfg.height, 32)); codec.reserve_clip_output(512 << 20)
d = torch.from_numpy(clip[mine.first:mine.first + mine.count].reshape(-1)).cuda() if mine else None
for rep in range(2):   # rep 0 warms the workspaces up
    codec.Reset()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rng, stream, sizes, fts = shard.encode_sharded(codec, None, keys, rank, world, dist if world > 1 else None,
                                                   device_ptr=d.data_ptr() if d is not None else None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    stream = stream.copy()
res = [None] * world
if world > 1:
    dist.gather_object((rng, stream, sizes, fts, dt), res if rank == 0 else None, dst=0)
else:
    res = [(rng, stream, sizes, fts, dt)]
if rank == 0:
    s, sz, ft = shard.gather_streams([r[:4] for r in res])
    whole = ScreenCodec(local); whole.Init(CodecParameters(cfg.width, cfg.height, 32)); whole.reserve_clip_output(1024 << 20)
    dall = torch.from_numpy(clip.reshape(-1)).cuda()
    for rep in range(2):
        whole.Reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
        ws, wsz, wft = whole.CompressClip(None, keys, device_ptr=dall.data_ptr(), n=n)
        torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    same = np.array_equal(s, ws) and np.array_equal(sz, wsz) and np.array_equal(ft, wft)
    dec = ScreenCodec(local); dec.Init(CodecParameters(cfg.width, cfg.height, 32))
    out = dec.DecompressClip(s, sz, ft)
    ok = np.array_equal(out.reshape(n, -1), clip.reshape(n, -1))
    print(f"[shard_demo] {name} {n} frames, keyframe interval {interval}, {world} rank(s): ranges {[(r.first, r.count) for r in ranges]}")
    print(f"[shard_demo] sharded encode {max(r[4] for r in res) * 1e3:.1f} ms (slowest rank), single-GPU encode {t1 * 1e3:.1f} ms, "
          f"bitstream identical: {same} (md5 {hashlib.md5(s.tobytes()).hexdigest()}), decode bit-exact: {ok}")
    assert same and ok
if world > 1:
    dist.barrier(); dist.destroy_process_group()
