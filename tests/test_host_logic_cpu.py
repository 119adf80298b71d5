"""CPU: host-side logic that does not need a GPU -- the clip generator, the keyframe policy, the
CPU arm of bench.py, and the N > 1 sharding / timing reduction under a 2-rank gloo group."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_keyframe_policy_matches_vfw_default():
    from screenpressor_b200 import synth

    # forced_kf = npframes + 1 >= interval (screenpressor.cpp:402-406): interval 500 -> I at 0 and 500
    k = synth.keyframe_flags(600, 500)
    assert list(np.nonzero(k)[0]) == [0, 500]
    assert list(np.nonzero(synth.keyframe_flags(10, 1))[0]) == list(range(10))
    assert list(np.nonzero(synth.keyframe_flags(10, 4))[0]) == [0, 4, 8]


def test_synthetic_clips_are_deterministic_and_well_formed():
    from screenpressor_b200 import synth

    for name, cfg in synth.CONFIGS.items():
        a = synth.make_clip(cfg, 3)
        b = synth.make_clip(cfg, 3)
        assert np.array_equal(a, b), name
        if cfg.bpp == 32:
            assert a.shape == (3, cfg.height, cfg.width, 4) and (a[..., 3] == 255).all()
        else:
            stride = (cfg.width * 3 + 3) & ~3
            assert a.shape == (3, cfg.height, stride)
        assert not np.array_equal(a[0], a[1]) or cfg.kind == "multimon"


def test_cfg5_contains_duplicate_frames_and_cfg2_scrolls():
    from screenpressor_b200 import synth

    c5 = synth.make_clip(synth.CONFIGS["cfg5_5120x1440"], 40)
    dups = sum(np.array_equal(c5[i], c5[i - 1]) for i in range(1, 40))
    assert dups >= 1
    c2 = synth.make_clip(synth.CONFIGS["cfg2_1080p_rgb32"], 32)
    changed = [(c2[i] != c2[i - 1]).any(axis=2).sum() for i in (29, 30)]
    assert changed[1] > 2 * changed[0]  # frame 30 scrolls the main text pane (on top of the window drag)


def test_bench_reference_arm_prints_contract_line(oracle_built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frames", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    assert line["metric"] == "1080p_rgb32_encode_decode_frames_per_s"
    sys.path.insert(0, ROOT)
    import bench

    assert line["config"] == bench.workload_config(4)   # the dict the CUDA arm prints: the driver compares the two
    assert "all 4 frames" in line["cpu_baseline"]["sample"] and line["scaling"] == "weak"


def _rank_main(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import bench
    from screenpressor_b200 import shard, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # N > 1: ONE clip cut by GOP-aligned frame ranges (strong scaling); every rank generates its own range only
    cfg = synth.CONFIGS["cfg1_720p_rgb24"]
    frames, interval = 16, 2
    keys = synth.keyframe_flags(frames, interval)
    ranges = shard.assign_ranges(keys, world)
    mine = next(r for r in ranges if r.rank == rank)
    part = synth.make_clip_range(cfg, mine.first, mine.count)
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # timing = max over ranks
    q.put((rank, mine.first, mine.count, int(part.astype(np.uint64).sum() % 1000003), float(t.item()), bench.split_config(1200, 150, world)["gops_per_clip"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduction_gloo():
    import torch.multiprocessing as mp

    from screenpressor_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, sum0, t0, g0), (r1, f1, c1, sum1, t1, g1) = res
    assert (f0, c0, f1, c1) == (0, 8, 8, 8)          # the two ranges partition the clip at a keyframe
    whole = synth.make_clip(synth.CONFIGS["cfg1_720p_rgb24"], 16)
    assert sum0 == int(whole[:8].astype(np.uint64).sum() % 1000003) and sum1 == int(whole[8:].astype(np.uint64).sum() % 1000003)
    assert t0 == t1 == 11.0                         # MAX over ranks
    assert g0 == g1 == 8


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


# ---- frame-range sharding (screenpressor_b200/shard.py): partition + relay order under a 2-rank gloo group -----------
class _StubCodec:
    """Stands in for the CUDA codec on CPU: the 'bitstream' of frame i is (i, state) so the test can see which state
    every range was encoded with; state = number of frames encoded by the ranges before (travels in the blob)."""

    def __init__(self):
        self.state = np.zeros(4, np.uint8)
        self.hooks = (None, None)

    def ExportRangeState(self, full=False):
        return self.state.copy()

    def ImportRangeState(self, blob):
        self.state = np.array(blob, np.uint8)

    def set_mvs_hooks(self, wait=None, ready=None):
        self.hooks = (wait, ready)

    def CompressClip(self, frames, keys, device_ptr=None, n=None):
        if self.hooks[0]:
            self.hooks[0]()
        seen = int(self.state[0])
        self.state = self.state.copy()
        self.state[0] = seen + n
        if self.hooks[1]:
            self.hooks[1]()
        stream = np.repeat(np.uint8(seen), n)
        return stream, np.ones(n, np.uint32), (1 - np.asarray(keys, np.uint8))


def test_assign_ranges_cuts_only_at_keyframes():
    from screenpressor_b200 import shard, synth

    k = synth.keyframe_flags(3600, 450)
    for world in (1, 2, 4, 8, 16):
        r = shard.assign_ranges(k, world)
        assert r[0].first == 0 and sum(x.count for x in r) == 3600 and len(r) == min(world, 8)
        assert all(k[x.first] for x in r) and all(a.first + a.count == b.first for a, b in zip(r, r[1:]))
    r = shard.assign_ranges(synth.keyframe_flags(600, 500), 8)   # cfg 2 has two GOPs: ranks beyond them stay idle
    assert [(x.first, x.count) for x in r] == [(0, 500), (500, 100)]
    assert shard.gop_ranges(np.array([0, 0, 1, 0, 1], np.uint8)) == [(0, 2), (2, 2), (4, 1)]
    # a requested keyframe that is a single-colour frame does not start a GOP in the reference (screencap.cpp:1488-1511):
    # the planner must not cut there
    clip = np.zeros((6, 4, 8, 4), np.uint8)
    clip[:, 1, 2, 0] = 7            # not flat ...
    clip[2] = 0
    clip[2, ..., :3] = (9, 9, 9)    # ... except frame 2
    keys = np.array([1, 0, 1, 0, 1, 0], np.uint8)
    assert shard.flat_keyframes(clip, keys) == [2]
    assert shard.gop_ranges(keys, shard.flat_keyframes(clip, keys)) == [(0, 4), (4, 2)]
    assert [(x.first, x.count) for x in shard.assign_ranges(keys, 3, [2])] == [(0, 4), (4, 2)]


def _shard_rank_main(rank, world, port, q, pipelined):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from screenpressor_b200 import shard, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys = synth.keyframe_flags(40, 10)
    rng, stream, sizes, fts = shard.encode_sharded(_StubCodec(), None, keys, rank, world, dist, pipelined=pipelined)
    q.put((rank, (rng.first, rng.count), stream.tolist(), fts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("pipelined", [False, True])
def test_sharded_encode_relays_state_in_rank_order_gloo(pipelined):
    import torch.multiprocessing as mp

    from screenpressor_b200 import shard

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + int(pipelined)
    procs = [ctx.Process(target=_shard_rank_main, args=(r, 2, port, q, pipelined)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 20) and res[1][1] == (20, 20)
    assert set(res[0][2]) == {0} and set(res[1][2]) == {20}      # rank 1 encoded with the state rank 0 left behind
    assert res[1][3][0] == 0                                        # a range starts on a keyframe
    parts = [(shard.FrameRange(f, c, r), np.array(s, np.uint8), np.ones(c, np.uint32), np.array(t, np.uint8)) for r, (f, c), s, t in res]
    stream, sizes, fts = shard.gather_streams(parts[::-1])
    assert stream.tolist() == [0] * 20 + [20] * 20 and sizes.size == 40


def test_cpp_range_planner_equals_the_python_planner():
    """csrc/multi.cu plans the frame ranges of the one-process multi-GPU entry, screenpressor_b200/shard.py those of the
    one-process-per-GPU driver (bench.py --gpus N): both must cut a clip the same way (scpr_plan_ranges needs no device)."""
    import ctypes as C
    import os
    import subprocess

    import numpy as np

    from screenpressor_b200 import shard

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(root, "screenpressor_b200", "csrc")], check=True)
    lib = C.CDLL(os.path.join(root, "screenpressor_b200", "libscpr_b200.so"))
    lib.scpr_plan_ranges.restype = C.c_int
    lib.scpr_plan_ranges.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(1, 400))
        keys = (rng.random(n) < rng.choice([0.01, 0.05, 0.3])).astype(np.uint8)
        keys[0] = 1
        world = int(rng.integers(1, 10))
        want = [(r.first, r.count) for r in shard.assign_ranges(keys, world)]
        first, count = np.zeros(world, np.int32), np.zeros(world, np.int32)
        k = lib.scpr_plan_ranges(keys.ctypes.data, n, world, first.ctypes.data, count.ctypes.data)
        got = [(int(first[i]), int(count[i])) for i in range(k)]
        assert got == want, (trial, n, world, np.flatnonzero(keys).tolist(), got, want)
        assert sum(c for _, c in got) == n and all(keys[f] for f, _ in got)
