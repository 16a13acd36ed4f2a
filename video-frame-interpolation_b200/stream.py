"""Frame-pair streaming around the hot path (BASELINE config 5).

The reference's video loop (/root/reference/inference.py:160-199) runs the model on one pair at a time, converts every
frame on the CPU (`process_frame`, :45-49: ToTensor + Normalize), and blocks on a device->host copy per written frame
(`denormalize_frame`, :52-58).  With the hot path on the GPU that loop is host-bound.  This module keeps its *output
stream* -- the same frames, in the same order, bit for bit -- and changes how it is produced:

* ``plan_stream`` restates the loop's control flow as data: which frame pairs go through the model and what is written in
  which order (including the reference's quirks: the prediction of a pair is written ``interpolation_factor`` times, the
  first frame of a pair is written *after* its predictions and as the normalise -> denormalise round trip of the frame,
  the very last frame is written raw).
* ``PairStreamer`` shards the pairs over ranks (independent units, no collective: ``shard.shard_pairs``), batches them,
  moves uint8 frames through pinned double buffers on a copy stream, does normalise / denormalise on the device in the
  reference's arithmetic (fp32 ToTensor + Normalize; float64 de-normalisation as numpy does at :55-57), and reads results
  back as uint8 on a second copy stream, so H2D of batch i+1, the model on batch i and D2H of batch i-1 overlap and the
  only host synchronisation is one event per batch.

The model is whatever the caller hands over (the reference's ``EMA_VFI`` with ``vfi_b200.install()`` applied, in
practice).  Nothing here computes the hot path: the warp / DCN kernels are reached through the model's own call sites,
and they refuse CPU tensors (``_lib.require_cuda``).  This module is host plumbing only -- scheduling, staging, ordering --
which is why it also accepts ``device="cpu"``: that is how tests/test_stream.py covers the N > 1 ownership logic over gloo
without a GPU, with a stand-in model; it is not a fallback for the kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import shard

MEAN = (0.485, 0.456, 0.406)          # inference.py:38-41 (applied to the channels as cv2 delivers them)
STD = (0.229, 0.224, 0.225)

PRED, ROUND_TRIP, RAW = "pred", "round_trip", "raw"


@dataclass(frozen=True)
class Emit:
    """One written frame: ``pred`` = model output of pair ``index``; ``round_trip`` = denormalise(normalise(frame
    ``index``)) (inference.py:184-185, :195); ``raw`` = frame ``index`` as read (:166)."""
    kind: str
    index: int


class StreamPlanner:
    """The control flow of inference.py:139-199, one read frame at a time (so a stream of unknown length can be planned as it
    arrives).  ``push(i)`` stands for a successful ``cap.read()`` returning frame ``i``; ``finish()`` for the failed read that
    ends the video.  Both return what the loop writes at that point as ``(position, Emit)``; ``push`` also returns the frame
    pair handed to the model, if any.  Pairs are numbered in the order they appear (``Emit(PRED, k)`` = the k-th pair)."""

    def __init__(self, frame_interval: int = 1, interpolation_factor: int = 1):
        if frame_interval < 1 or interpolation_factor < 0:
            raise ValueError("frame_interval >= 1, interpolation_factor >= 0 required")
        self.interval, self.factor = frame_interval, interpolation_factor
        self.first: Optional[int] = None      # index of the loop's frame1
        self.frame_num = 0                    # the loop counter (:161)
        self.n_pairs = 0
        self.position = 0                     # frames written so far

    def _emit(self, out: List[Tuple[int, "Emit"]], kind: str, index: int) -> None:
        out.append((self.position, Emit(kind, index)))
        self.position += 1

    def push(self, index: int) -> Tuple[Optional[Tuple[int, int]], List[Tuple[int, "Emit"]]]:
        out: List[Tuple[int, Emit]] = []
        if self.first is None:                                    # :139-144, the read before the loop
            self.first = index
            return None, out
        self.frame_num += 1
        if self.frame_num % self.interval:                        # :188-193: the frame replaces frame1, nothing is written
            self.first = index
            return None, out
        pair = (self.first, index)
        for _ in range(self.factor):                              # :172-182
            self._emit(out, PRED, self.n_pairs)
        self._emit(out, ROUND_TRIP, self.first)                   # :184-185
        self.n_pairs += 1
        self.first = index
        return pair, out

    def finish(self) -> List[Tuple[int, "Emit"]]:
        out: List[Tuple[int, Emit]] = []
        if self.first is None:                                    # :146-150 "video is empty": nothing is written
            return out
        self.frame_num += 1
        self._emit(out, ROUND_TRIP if self.frame_num % self.interval else RAW, self.first)   # :195-196 / :166
        return out


def plan_stream(num_frames: int, frame_interval: int = 1, interpolation_factor: int = 1) -> Tuple[List[Tuple[int, int]], List[Emit]]:
    """Control flow of inference.py:139-199 for a video of ``num_frames`` readable frames, as data.

    Returns ``(pairs, emits)``: ``pairs[k] = (i, j)`` are the frame indices of the k-th model call's inputs; ``emits`` is the
    output stream in writing order.  ``interpolation_factor`` is what :103-117 derive from the frame rates.
    """
    if num_frames < 0:
        raise ValueError("num_frames >= 0 required")
    planner = StreamPlanner(frame_interval, interpolation_factor)
    pairs: List[Tuple[int, int]] = []
    emits: List[Emit] = []
    for i in range(num_frames):
        pair, out = planner.push(i)
        if pair is not None:
            pairs.append(pair)
        emits.extend(e for _, e in out)
    emits.extend(e for _, e in planner.finish())
    return pairs, emits


def normalize_u8(frames_u8: torch.Tensor) -> torch.Tensor:
    """[N,H,W,3] uint8 -> [N,3,H,W] fp32, the arithmetic of ToTensor (x / 255) + Normalize ((x - mean) / std)."""
    # division by a device TENSOR: torch's CUDA kernel for a Python-scalar divisor multiplies by 1/255 instead, which is
    # one ulp away from ToTensor's IEEE division for some levels (measured on the B200: 1 LSB in the written frames)
    x = frames_u8.permute(0, 3, 1, 2).to(torch.float32)
    x = x.div_(torch.full((), 255.0, dtype=torch.float32, device=x.device))
    mean = torch.tensor(MEAN, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(STD, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
    return x.sub_(mean).div_(std)


def denormalize_u8(x: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] any float dtype -> [N,H,W,3] uint8 as inference.py:52-58: fp32 values times float64 constants, clip, * 255,
    truncation."""
    y = x.float().permute(0, 2, 3, 1).to(torch.float64)
    std = torch.tensor(STD, dtype=torch.float64, device=x.device)
    mean = torch.tensor(MEAN, dtype=torch.float64, device=x.device)
    y = (y * std + mean).clamp_(0, 1).mul_(255)
    return y.to(torch.uint8)


def _rows(t: torch.Tensor, idx: List[int]) -> torch.Tensor:
    """Rows ``idx`` of ``t``: a view when they are consecutive (the usual case: pairs (i, i+1), (i+1, i+2), ...), so no index
    tensor has to cross to the device; a gather otherwise."""
    if idx and idx == list(range(idx[0], idx[0] + len(idx))):
        return t[idx[0]:idx[0] + len(idx)]
    return t[torch.tensor(idx, dtype=torch.long, device=t.device)]


class _Slot:
    """One pinned staging area: uint8 frames in, uint8 results out, and the events that guard their reuse."""

    def __init__(self, n_in: int, n_out: int, H: int, W: int, cuda: bool):
        mk = (lambda *s: torch.empty(s, dtype=torch.uint8).pin_memory()) if cuda else (lambda *s: torch.empty(s, dtype=torch.uint8))
        self.h_in = mk(n_in, H, W, 3)
        self.h_out = mk(n_out, H, W, 3)
        self.done: Optional[torch.cuda.Event] = None


class PairStreamer:
    """Runs ``model(frame1, frame2)`` over the pairs of a frame sequence and yields the reference's output stream.

    ``frames``: a sequence of [H,W,3] uint8 arrays (already resized, inference.py:47).  ``batch_pairs`` pairs per model
    call (the reference uses 1).  ``topology``: this rank's place in a ``world``-rank job; each rank computes a
    contiguous chunk of the pairs; ``run`` yields only the written frames this rank owns, ``run_all`` returns the full
    stream on every rank after one host-side object gather.
    """

    def __init__(self, model: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], device, *, batch_pairs: int = 8,
                 topology: Optional[shard.Topology] = None, autocast_dtype: Optional[torch.dtype] = None):
        self.model = model
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        if batch_pairs < 1:
            raise ValueError("batch_pairs must be >= 1")
        self.batch_pairs = batch_pairs
        self.topo = topology or shard.Topology(0, 1, 0)
        self.autocast_dtype = autocast_dtype
        self._streams = None
        self.stats: Dict[str, int] = {"model_calls": 0, "h2d_bytes": 0, "d2h_bytes": 0, "frames_uploaded": 0}

    # ------------------------------------------------------------------------------------------ ownership
    def owned_pairs(self, num_pairs: int) -> List[int]:
        return shard.shard_pairs(num_pairs, self.topo.rank, self.topo.world, mode="contiguous")

    @staticmethod
    def emit_owner(e: Emit, start_of: Dict[int, int], owner_of_pair: Sequence[int]) -> int:
        """Rank that produces an emit: predictions and the round trip of a pair's first frame belong to the pair's owner;
        the closing frame (no pair starts at it) belongs to the owner of the last pair, or rank 0 when there is none.
        ``start_of[f]`` = the pair whose first frame is ``f``."""
        if e.kind == PRED:
            return owner_of_pair[e.index]
        if e.index in start_of:
            return owner_of_pair[start_of[e.index]]
        return owner_of_pair[-1] if owner_of_pair else 0

    # ------------------------------------------------------------------------------------------ the loop
    def _model(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        self.stats["model_calls"] += 1
        with torch.no_grad():
            if self.autocast_dtype is not None:
                with torch.autocast(self.device.type, dtype=self.autocast_dtype):
                    return self.model(a, b)
            return self.model(a, b)

    def run(self, frames: Sequence[np.ndarray], frame_interval: int = 1, interpolation_factor: int = 1) -> Iterator[Tuple[int, np.ndarray]]:
        """Yields ``(position, frame)`` for every emit this rank owns, in increasing position (= index into the reference's
        output stream)."""
        pairs, emits = plan_stream(len(frames), frame_interval, interpolation_factor)
        world = self.topo.world
        owner_of_pair = [0] * len(pairs)
        for r in range(world):
            for k in shard.shard_pairs(len(pairs), r, world, mode="contiguous"):
                owner_of_pair[k] = r
        start_of = {i: k for k, (i, _) in enumerate(pairs)}
        mine = [(pos, e) for pos, e in enumerate(emits) if self.emit_owner(e, start_of, owner_of_pair) == self.topo.rank]
        if not mine:
            return
        my_pairs = [k for k in range(len(pairs)) if owner_of_pair[k] == self.topo.rank]
        # emits grouped by the batch that produces them: a batch = up to batch_pairs consecutive owned pairs
        batches = list(shard.batches(my_pairs, self.batch_pairs)) or [[]]
        batch_of_pair = {k: bi for bi, ks in enumerate(batches) for k in ks}
        per_batch: List[List[Tuple[int, Emit]]] = [[] for _ in batches]
        for pos, e in mine:
            if e.kind == PRED:
                per_batch[batch_of_pair[e.index]].append((pos, e))
            else:
                k = start_of.get(e.index)
                per_batch[batch_of_pair[k] if k in batch_of_pair else len(batches) - 1].append((pos, e))

        items = []
        for bi, ks in enumerate(batches):
            local = {k: i for i, k in enumerate(ks)}
            items.append(([pairs[k] for k in ks],
                          [(pos, Emit(PRED, local[e.index]) if e.kind == PRED else e) for pos, e in per_batch[bi]], frames))
        yield from self._execute(items)

    def run_iter(self, frames, frame_interval: int = 1, interpolation_factor: int = 1) -> Iterator[Tuple[int, np.ndarray]]:
        """``run`` for a stream of unknown length: ``frames`` is any iterable of [H,W,3] uint8 arrays, planned as it arrives
        (``StreamPlanner``), so at most ``2 * batch_pairs + 1`` frames are held (the batch in flight and the open one).  Without the total, ranks take *batches* of
        ``batch_pairs`` consecutive pairs round-robin (batch b -> rank b % world); every rank walks the whole iterable and stages
        only its own batches.  The closing frame goes with the last batch."""
        planner = StreamPlanner(frame_interval, interpolation_factor)
        rank, world = self.topo.rank, self.topo.world

        def items():
            held: Dict[int, np.ndarray] = {}
            plist: List[Tuple[int, int]] = []
            elist: List[Tuple[int, Emit]] = []
            batch_no, last_owner = 0, 0
            for idx, fr in enumerate(frames):
                held[idx] = fr
                pair, out = planner.push(idx)
                if pair is not None:
                    plist.append(pair)
                    elist.extend((pos, Emit(PRED, len(plist) - 1) if e.kind == PRED else e) for pos, e in out)
                    if len(plist) == self.batch_pairs:
                        last_owner = batch_no % world
                        if last_owner == rank:
                            yield plist, elist, dict(held)
                        batch_no += 1
                        plist, elist = [], []
                keep = {planner.first} | {f for pr in plist for f in pr}
                for f in [f for f in held if f not in keep]:
                    del held[f]
            closing = planner.finish()
            owner = batch_no % world if plist else last_owner
            if owner == rank and (plist or closing):
                yield plist, elist + closing, held

        yield from self._execute(items())

    def _execute(self, items) -> Iterator[Tuple[int, np.ndarray]]:
        """The double-buffered loop.  ``items`` yields ``(pairs, emits, frames)`` per batch: frame-index pairs for one model call,
        the ``(position, Emit)`` this batch writes (``Emit(PRED, i)`` = the i-th pair of the batch), and a frame lookup."""
        slots: Optional[List[_Slot]] = None
        H = W = 0
        if self.cuda and self._streams is None:
            self._streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        pending: Optional[Tuple[_Slot, List[Tuple[int, int]]]] = None       # (slot, [(position, row of h_out)])

        def drain(p):
            slot, rows = p
            if slot.done is not None:
                slot.done.synchronize()              # the one host wait per batch
            for pos, row in rows:
                yield pos, slot.h_out[row].numpy().copy()

        for bi, (plist, elist, lookup) in enumerate(items):
            # frames this batch touches: the pairs' inputs plus any raw / round-trip frame it writes, each uploaded once
            need: List[int] = []
            for pr in plist:
                need.extend(pr)
            need.extend(e.index for _, e in elist if e.kind != PRED)
            uniq = sorted(set(need))
            if slots is None:
                H, W = lookup[uniq[0]].shape[:2]
                n_in_max = 2 * self.batch_pairs + 1
                slots = [_Slot(n_in_max, n_in_max + self.batch_pairs, H, W, self.cuda) for _ in range(2)]
            slot = slots[bi % 2]
            row_of = {f: i for i, f in enumerate(uniq)}
            for f, i in row_of.items():
                fr = lookup[f]
                if fr.shape != (H, W, 3) or fr.dtype != np.uint8:
                    raise ValueError(f"frame {f}: expected uint8 [{H},{W},3], got {fr.dtype} {fr.shape}")
                slot.h_in[i].copy_(torch.from_numpy(np.ascontiguousarray(fr)))
            n_in = len(uniq)
            self.stats["frames_uploaded"] += n_in
            self.stats["h2d_bytes"] += n_in * H * W * 3
            if self.cuda:
                s_in, s_out = self._streams
                cur = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(s_in):
                    d_u8 = slot.h_in[:n_in].to(self.device, non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(s_in)
                cur.wait_event(up)
                d_u8.record_stream(cur)
            else:
                d_u8 = slot.h_in[:n_in].clone()
            x = normalize_u8(d_u8)
            outs: List[torch.Tensor] = []
            rows: List[Tuple[int, int]] = []
            if plist:
                a = _rows(x, [row_of[i] for i, _ in plist])
                b = _rows(x, [row_of[j] for _, j in plist])
                pred = denormalize_u8(self._model(a, b))
                outs.append(pred)
            rt_frames = sorted({e.index for _, e in elist if e.kind == ROUND_TRIP})
            if rt_frames:
                outs.append(denormalize_u8(_rows(x, [row_of[f] for f in rt_frames])))
            rt_row = {f: len(plist) + i for i, f in enumerate(rt_frames)}
            raw_frames = sorted({e.index for _, e in elist if e.kind == RAW})
            if raw_frames:
                outs.append(_rows(d_u8, [row_of[f] for f in raw_frames]))
            raw_row = {f: len(plist) + len(rt_frames) + i for i, f in enumerate(raw_frames)}
            for pos, e in elist:
                rows.append((pos, e.index if e.kind == PRED else rt_row[e.index] if e.kind == ROUND_TRIP else raw_row[e.index]))
            res = torch.cat(outs, 0) if outs else None
            # drain the previous batch only now: its D2H ran while this batch was staged and enqueued
            if pending is not None:
                yield from drain(pending)
                pending = None
            if res is not None:
                n_out = res.shape[0]
                self.stats["d2h_bytes"] += n_out * H * W * 3
                if self.cuda:
                    ready = torch.cuda.Event()
                    ready.record(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ready)
                        slot.h_out[:n_out].copy_(res, non_blocking=True)
                        res.record_stream(s_out)
                        slot.done = torch.cuda.Event()
                        slot.done.record(s_out)
                else:
                    slot.h_out[:n_out].copy_(res)
                pending = (slot, sorted(rows))
        if pending is not None:
            yield from drain(pending)

    def run_all(self, frames: Sequence[np.ndarray], frame_interval: int = 1, interpolation_factor: int = 1) -> List[np.ndarray]:
        """The whole output stream on every rank (host-side gather by position; single process: no communication)."""
        local = dict(self.run(frames, frame_interval, interpolation_factor))
        merged = shard.gather_by_index(local, self.topo.world)
        return [merged[i] for i in range(len(merged))]


# ------------------------------------------------------------------------------------------------ video in / video out
def choose_interpolation_factor(fps: float, target_fps: Optional[float], max_interpolation_factor: int = 4) -> Tuple[int, float]:
    """inference.py:103-123: the factor whose output rate is closest to 60 fps when no target is given (first wins on ties),
    ``round(target / fps - 1)`` otherwise; the target is capped at what the factor can deliver."""
    if target_fps is None:
        best, best_diff = 0, float("inf")
        for k in range(1, max_interpolation_factor + 1):
            diff = abs(fps * (k + 1) - 60)
            if diff < best_diff:
                best, best_diff = k, diff
        factor, target_fps = best, fps * (best + 1)
    else:
        factor = round(target_fps / fps - 1)
    return factor, min(target_fps, fps * (factor + 1))


def stream_video(input_video_path: str, output_video_path: str, model, device, *, target_fps: Optional[float] = None,
                 max_interpolation_factor: int = 4, frame_interval: int = 1, codec: str = "mp4v", scale: float = 0.5,
                 batch_pairs: int = 8, autocast_dtype="reference", queue_depth: int = 32) -> int:
    """``interpolate_video`` of inference.py:60-205 with the frame loop replaced by ``PairStreamer.run_iter``: same arguments
    (the model object instead of a checkpoint path), same file written.  ``autocast_dtype="reference"`` (default) mirrors
    inference.py:158-159 -- ``torch.cuda.amp.autocast()``, i.e. fp16 on a CUDA device and nothing on CPU; pass a dtype or ``None``
    to choose otherwise.  Decoding + resizing (:47) and encoding run in
    their own threads behind bounded queues, so they overlap the copies and the model.  Returns the number of frames written.
    A video that cannot be opened raises ``ValueError`` (the reference logs the same message and returns)."""
    import queue
    import threading

    import cv2

    cap = cv2.VideoCapture(input_video_path)
    if not cap.isOpened():
        raise ValueError(f"cannot open video file: {input_video_path}")
    fps = cap.get(cv2.CAP_PROP_FPS)
    new_w = int(int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) * scale)        # :86-93
    new_h = int(int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)) * scale)
    factor, target_fps = choose_interpolation_factor(fps, target_fps, max_interpolation_factor)
    out = cv2.VideoWriter(output_video_path, cv2.VideoWriter_fourcc(*codec), target_fps, (new_w, new_h))   # :127-128

    END = object()
    q_in: "queue.Queue" = queue.Queue(maxsize=queue_depth)
    q_out: "queue.Queue" = queue.Queue(maxsize=queue_depth)
    errors: List[BaseException] = []
    stop = threading.Event()

    def put(q, item) -> bool:
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def decode():
        try:
            while not stop.is_set():
                ok, frame = cap.read()
                if not ok:
                    break
                if not put(q_in, cv2.resize(frame, (new_w, new_h))):
                    return
        except BaseException as e:  # noqa: BLE001 - handed to the caller's thread
            errors.append(e)
        finally:
            put(q_in, END)

    def encode():
        try:
            while True:
                item = q_out.get()
                if item is END:
                    return
                out.write(item)
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            stop.set()

    def frames():
        while True:
            item = q_in.get()
            if item is END:
                return
            yield item

    t_dec, t_enc = threading.Thread(target=decode, daemon=True), threading.Thread(target=encode, daemon=True)
    t_dec.start()
    t_enc.start()
    written = 0
    try:
        if isinstance(autocast_dtype, str):
            if autocast_dtype != "reference":
                raise ValueError("autocast_dtype: a torch dtype, None, or 'reference'")
            autocast_dtype = torch.float16 if torch.device(device).type == "cuda" else None
        streamer = PairStreamer(model, device, batch_pairs=batch_pairs, autocast_dtype=autocast_dtype)
        for _, frame in streamer.run_iter(frames(), frame_interval, factor):
            if not put(q_out, frame):
                break
            written += 1
    finally:
        if not put(q_out, END):
            stop.set()
            try:
                q_out.put_nowait(END)
            except queue.Full:
                pass
        t_enc.join(timeout=60)
        stop.set()
        t_dec.join(timeout=10)
        cap.release()
        out.release()
    if errors:
        raise errors[0]
    return written
