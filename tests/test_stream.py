"""Frame-pair streaming (vfi_b200.stream) against a literal restatement of the reference's video loop.

``reference_loop`` below follows /root/reference/inference.py line by line (:45-58 frame conversion, :139-199 the loop) with
``cap.read()`` / ``out.write()`` replaced by a list of frames and a list of written frames, and the model by any callable.
It is test infrastructure: the product never runs it.  The streamer has to write the same frames in the same order, bit
for bit, whatever the batch size or the number of ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from vfi_b200 import shard, stream

MEAN = np.array([0.485, 0.456, 0.406])
STD = np.array([0.229, 0.224, 0.225])


def _to_tensor_normalize(frame):
    """transforms.ToTensor() + transforms.Normalize(mean, std) (inference.py:38-41) on one HWC uint8 frame."""
    t = torch.from_numpy(frame).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    mean = torch.as_tensor(MEAN, dtype=torch.float32).view(-1, 1, 1)
    std = torch.as_tensor(STD, dtype=torch.float32).view(-1, 1, 1)
    return t.sub_(mean).div_(std).unsqueeze(0)


def _denormalize_frame(t):
    """inference.py:52-58."""
    frame = t.squeeze(0).cpu().float().numpy()
    frame = np.transpose(frame, (1, 2, 0))
    frame = (frame * STD) + MEAN
    frame = np.clip(frame, 0, 1)
    return (frame * 255).astype(np.uint8)


def reference_loop(frames, model, frame_interval, interpolation_factor):
    """inference.py:139-199 with the video I/O replaced by lists."""
    written = []
    it = iter(frames)

    def read():
        f = next(it, None)
        return f is not None, f

    success, frame = read()
    if not success:
        return written
    frame1, frame1_tensor = frame, _to_tensor_normalize(frame)
    frame_num = 0
    with torch.no_grad():
        while success:
            frame_num += 1
            if frame_num % frame_interval == 0:
                success, frame2 = read()
                if not success:
                    written.append(frame1)
                    break
                frame2_tensor = _to_tensor_normalize(frame2)
                for _ in range(1, interpolation_factor + 1):
                    written.append(_denormalize_frame(model(frame1_tensor, frame2_tensor)))
                written.append(_denormalize_frame(frame1_tensor))
                frame1, frame1_tensor = frame2, frame2_tensor
            else:
                success, frame2 = read()
                if success:
                    frame1, frame1_tensor = frame2, _to_tensor_normalize(frame2)
                else:
                    written.append(_denormalize_frame(frame1_tensor))
                    break
    return written


def fake_model(a, b):
    """Per-sample, batch-size independent, non-linear enough that pair order and frame identity matter."""
    return 0.25 * a + 0.75 * b.flip(-1) + 0.1 * torch.tanh(a * b)


def _frames(n, h=6, w=10, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for _ in range(n)]


def test_torchvision_transform_is_what_the_restatement_does():
    tv = pytest.importorskip("torchvision.transforms")
    tf = tv.Compose([tv.ToTensor(), tv.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    f = _frames(1, 9, 7, seed=3)[0]
    assert torch.equal(tf(f).unsqueeze(0), _to_tensor_normalize(f))
    assert torch.equal(stream.normalize_u8(torch.from_numpy(f)[None]), _to_tensor_normalize(f))


def test_denormalize_matches_numpy_float64_path_bit_for_bit():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 17, 13, generator=g) * 1.5
    ours = stream.denormalize_u8(x).numpy()
    for i in range(3):
        assert np.array_equal(ours[i], _denormalize_frame(x[i:i + 1]))
    # every uint8 level survives the normalise -> denormalise round trip exactly as in the reference
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, 2)
    assert np.array_equal(stream.denormalize_u8(stream.normalize_u8(torch.from_numpy(ramp)[None]))[0].numpy(),
                          _denormalize_frame(_to_tensor_normalize(ramp)))


@pytest.mark.parametrize("interval", [1, 2, 3])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 10])
def test_plan_is_the_reference_control_flow(n, interval):
    calls = []

    def recording_model(a, b):
        calls.append((int(a[0, 0, 0, 0] * 1000), int(b[0, 0, 0, 0] * 1000)))
        return a

    frames = _frames(n, 2, 2, seed=n)
    written = reference_loop(frames, recording_model, interval, 2)
    pairs, emits = stream.plan_stream(n, interval, 2)
    assert len(emits) == len(written)
    assert len(pairs) * 2 == len(calls)                      # the reference calls the model once per written prediction
    for (i, j), c in zip(pairs, calls[::2]):
        ti, tj = _to_tensor_normalize(frames[i]), _to_tensor_normalize(frames[j])
        assert c == (int(ti[0, 0, 0, 0] * 1000), int(tj[0, 0, 0, 0] * 1000))
    if n:
        assert emits[-1].kind in (stream.RAW, stream.ROUND_TRIP)
    with pytest.raises(ValueError):
        stream.plan_stream(3, 0, 1)


@pytest.mark.parametrize("batch_pairs", [1, 3, 8])
@pytest.mark.parametrize("interval,factor", [(1, 1), (1, 3), (2, 1), (3, 2), (1, 0)])
@pytest.mark.parametrize("n", [0, 1, 2, 5, 12])
def test_stream_equals_reference_loop(n, interval, factor, batch_pairs):
    frames = _frames(n, seed=10 + n)
    want = reference_loop(frames, fake_model, interval, factor)
    s = stream.PairStreamer(fake_model, "cpu", batch_pairs=batch_pairs)
    got = list(s.run(frames, interval, factor))
    assert [p for p, _ in got] == list(range(len(want)))
    for (_, g), w in zip(got, want):
        assert g.dtype == np.uint8 and np.array_equal(g, w)
    pairs, _ = stream.plan_stream(n, interval, factor)
    assert s.stats["model_calls"] == -(-len(pairs) // batch_pairs)        # one call per batch, not per written frame


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("n,interval", [(1, 1), (2, 1), (9, 1), (11, 2), (30, 1)])
def test_ranks_partition_the_stream(n, interval, world):
    frames = _frames(n, seed=n)
    want = reference_loop(frames, fake_model, interval, 2)
    merged, per_rank = {}, []
    for r in range(world):
        s = stream.PairStreamer(fake_model, "cpu", batch_pairs=2, topology=shard.Topology(r, world, r))
        mine = dict(s.run(frames, interval, 2))
        assert not (set(mine) & set(merged))               # every written frame has exactly one owner
        merged.update(mine)
        per_rank.append(s.stats["model_calls"])
    assert sorted(merged) == list(range(len(want)))
    assert all(np.array_equal(merged[i], want[i]) for i in range(len(want)))
    pairs, _ = stream.plan_stream(n, interval, 2)
    assert sum(per_rank) == sum(-(-len(shard.shard_pairs(len(pairs), r, world, "contiguous")) // 2) for r in range(world))


def test_bad_frames_are_rejected():
    s = stream.PairStreamer(fake_model, "cpu", batch_pairs=2)
    frames = _frames(3)
    frames[1] = frames[1].astype(np.float32)
    with pytest.raises(ValueError):
        list(s.run(frames))
    with pytest.raises(ValueError):
        stream.PairStreamer(fake_model, "cpu", batch_pairs=0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    topo = shard.init_distributed("gloo")
    frames = _frames(9, seed=5)
    out = stream.PairStreamer(fake_model, "cpu", batch_pairs=3, topology=topo).run_all(frames, 1, 2)
    want = reference_loop(frames, fake_model, 1, 2)
    q.put((rank, len(out) == len(want) and all(np.array_equal(a, b) for a, b in zip(out, want))))
    torch.distributed.destroy_process_group()


def test_world_size_2_gather_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


@pytest.mark.gpu
def test_stream_on_device_equals_reference_loop():
    """Pinned double buffers + copy streams: same frames as the reference loop run on the CPU (the fake model's fp32
    elementwise arithmetic is IEEE on both; tanh may differ in the last bit, so it is left out here)."""
    def model(a, b):
        return 0.25 * a + 0.75 * b.flip(-1)

    frames = _frames(21, 64, 96, seed=2)
    want = reference_loop(frames, model, 1, 2)
    s = stream.PairStreamer(model, "cuda:0", batch_pairs=4)
    got = s.run_all(frames, 1, 2)
    assert len(got) == len(want)
    diff = max(int(np.abs(g.astype(np.int16) - w.astype(np.int16)).max()) for g, w in zip(got, want))
    assert diff == 0, diff
    assert s.stats["model_calls"] == 5 and s.stats["d2h_bytes"] > 0, s.stats
    it = dict(stream.PairStreamer(model, "cuda:0", batch_pairs=4).run_iter(iter(frames), 1, 2))     # same loop, planned as it arrives
    assert sorted(it) == list(range(len(want))) and all(np.array_equal(it[i], want[i]) for i in range(len(want)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_stream_with_the_reference_model_on_cpu():
    """The unmodified EMA_VFI (random init, eval, CPU) through the streamer against the reference loop: batch 1 performs the
    same model calls, so the written frames are identical; batches of 3 may pick different CPU conv kernels (at most one
    uint8 level)."""
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.ema_vfi import EMA_VFI
    finally:
        sys.path.remove("/root/reference")
    torch.manual_seed(0)
    model = EMA_VFI().eval()
    frames = _frames(5, 24, 32, seed=8)
    want = reference_loop(frames, model, 1, 1)
    got1 = stream.PairStreamer(model, "cpu", batch_pairs=1).run_all(frames, 1, 1)
    assert len(got1) == len(want) and all(np.array_equal(a, b) for a, b in zip(got1, want))
    got3 = stream.PairStreamer(model, "cpu", batch_pairs=3).run_all(frames, 1, 1)
    worst = max(int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) for a, b in zip(got3, want))
    assert worst <= 1, worst


def test_fuzz_plan_ownership_and_stream():
    """Random (frames, interval, factor, batch, world): the union of the ranks' shares is the reference's stream."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=40, deadline=None)
    @hyp.given(n=st.integers(0, 14), interval=st.integers(1, 4), factor=st.integers(0, 3), batch=st.integers(1, 5),
               world=st.integers(1, 4), seed=st.integers(0, 99))
    def check(n, interval, factor, batch, world, seed):
        frames = _frames(n, 4, 6, seed=seed)
        want = reference_loop(frames, fake_model, interval, factor)
        merged = {}
        for r in range(world):
            s = stream.PairStreamer(fake_model, "cpu", batch_pairs=batch, topology=shard.Topology(r, world, r))
            got = list(s.run(frames, interval, factor))
            assert [p for p, _ in got] == sorted(p for p, _ in got)         # each rank yields in stream order
            assert not (set(p for p, _ in got) & set(merged))
            merged.update(got)
        assert sorted(merged) == list(range(len(want)))
        assert all(np.array_equal(merged[i], want[i]) for i in range(len(want)))

    check()


@pytest.mark.parametrize("batch_pairs", [1, 3, 8])
@pytest.mark.parametrize("interval,factor", [(1, 1), (1, 2), (2, 1), (3, 2), (1, 0)])
@pytest.mark.parametrize("n", [0, 1, 2, 5, 12])
def test_unbounded_stream_equals_reference_loop(n, interval, factor, batch_pairs):
    frames = _frames(n, seed=20 + n)
    want = reference_loop(frames, fake_model, interval, factor)
    s = stream.PairStreamer(fake_model, "cpu", batch_pairs=batch_pairs)
    got = list(s.run_iter((f for f in frames), interval, factor))          # a generator: no len(), one pass
    assert [p for p, _ in got] == list(range(len(want)))
    assert all(np.array_equal(g, w) for (_, g), w in zip(got, want))


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("n,interval", [(1, 1), (2, 1), (9, 1), (11, 2), (30, 1)])
def test_unbounded_stream_ranks_partition_by_batches(n, interval, world):
    frames = _frames(n, seed=n + 1)
    want = reference_loop(frames, fake_model, interval, 2)
    merged = {}
    for r in range(world):
        s = stream.PairStreamer(fake_model, "cpu", batch_pairs=2, topology=shard.Topology(r, world, r))
        mine = dict(s.run_iter(iter(frames), interval, 2))
        assert not (set(mine) & set(merged))
        merged.update(mine)
    assert sorted(merged) == list(range(len(want)))
    assert all(np.array_equal(merged[i], want[i]) for i in range(len(want)))


def test_unbounded_stream_holds_a_bounded_number_of_frames():
    import gc
    import weakref
    refs, peak = [], [0]

    def source(n):
        rng = np.random.default_rng(0)
        for _ in range(n):
            gc.collect()
            peak[0] = max(peak[0], sum(r() is not None for r in refs))
            f = rng.integers(0, 256, size=(4, 6, 3), dtype=np.uint8)
            refs.append(weakref.ref(f))
            yield f
            del f

    s = stream.PairStreamer(fake_model, "cpu", batch_pairs=3)
    n_out = sum(1 for _ in s.run_iter(source(60), 1, 1))
    assert n_out == 59 * 2 + 1
    assert peak[0] <= 2 * 3 + 1, peak[0]                 # the batch being executed (batch_pairs + 1) + the open one, which share a frame


def test_planner_is_incremental():
    pl = stream.StreamPlanner(2, 1)
    assert pl.push(0) == (None, [])
    assert pl.push(1) == (None, [])                       # frame_num 1: skipped, replaces frame1
    pair, out = pl.push(2)                                # frame_num 2: a pair (1, 2)
    assert pair == (1, 2) and [(p, e.kind, e.index) for p, e in out] == [(0, stream.PRED, 0), (1, stream.ROUND_TRIP, 1)]
    assert [(p, e.kind, e.index) for p, e in pl.finish()] == [(2, stream.ROUND_TRIP, 2)]
    with pytest.raises(ValueError):
        stream.StreamPlanner(0, 1)


def test_choose_interpolation_factor_follows_the_script():
    assert stream.choose_interpolation_factor(30.0, None, 4) == (1, 60.0)
    assert stream.choose_interpolation_factor(24.0, None, 4)[0] == 1          # |48-60| = |72-60|: the first wins
    assert stream.choose_interpolation_factor(15.0, None, 4) == (3, 60.0)
    assert stream.choose_interpolation_factor(25.0, 100.0, 4) == (3, 100.0)
    assert stream.choose_interpolation_factor(25.0, 60.0, 4) == (1, 50.0)     # round(1.4) = 1, target capped at 50


def _read_video(path):
    import cv2
    cap, frames = cv2.VideoCapture(str(path)), []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    fps = cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    return fps, frames


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("interval,n_frames", [(1, 6), (2, 7)])
def test_stream_video_writes_what_the_unmodified_inference_script_writes(tmp_path, interval, n_frames):
    """The reference's inference.py, run byte-for-byte unmodified (runpy, --device cpu, random-init checkpoint), against
    ``stream_video`` with the same model: the two output files decode to identical frames at the same frame rate."""
    import runpy
    import sys
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    src = tmp_path / "in.avi"
    w = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (64, 48))
    base = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    for i in range(n_frames):
        w.write(np.roll(base, 3 * i, axis=1))
    w.release()

    sys.path.insert(0, "/root/reference")
    argv, dont = sys.argv, sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from src.models.ema_vfi import EMA_VFI
        torch.manual_seed(1)
        model = EMA_VFI().eval()
        ckpt = tmp_path / "m.pth"
        torch.save(model.state_dict(), ckpt)
        ref_out = tmp_path / "ref.avi"
        sys.argv = ["inference.py", "--input_video", str(src), "--output_video", str(ref_out), "--model_path", str(ckpt),
                    "--device", "cpu", "--codec", "MJPG", "--frame_interval", str(interval), "--scale", "0.5"]
        runpy.run_path("/root/reference/inference.py", run_name="__main__")
    finally:
        sys.argv, sys.dont_write_bytecode = argv, dont
        sys.path.remove("/root/reference")

    ours_out = tmp_path / "ours.avi"
    n = stream.stream_video(str(src), str(ours_out), model, "cpu", codec="MJPG", frame_interval=interval, scale=0.5, batch_pairs=1)
    fps_r, fr = _read_video(ref_out)
    fps_o, fo = _read_video(ours_out)
    assert len(fr) > n_frames // interval and n == len(fo) == len(fr) and fps_r == fps_o
    assert all(np.array_equal(a, b) for a, b in zip(fr, fo))
    with pytest.raises(ValueError):
        stream.stream_video(str(tmp_path / "missing.avi"), str(ours_out), model, "cpu")
