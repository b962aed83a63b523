"""Mapping loop of the hot path: `GridTrainer` mirrors grid_opt/trainer.py:370-491 (per-level Adam
optimisers, 'coordinate' / 'coordinate+joint' / 'joint' schedule, level switching every
`max_epochs_in_level` epochs or on relative-change convergence) with `train_epoch` =
prepare_batch -> loss -> backward -> step (trainer.py:196-228), and `Mapper.mapping` mirrors
grid_opt/slam/mapper.py:65-98.

One epoch is one batch (reference datasets have `__len__ == 1`, sdf_rgbd.py:86-87).  On the fused
path an epoch is 3 launches: eikonal-sample count, the fused mapping step (scatter straight into
`feature.grad`), and one fused Adam(+zero) pass per active level.  Host buffers are staged through
pinned memory on a copy stream (`HostBatchStager`) so the H2D copy of batch i+1 overlaps step i.
"""
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from .loss import MisoLossMapping
from .models import GridNet
from .optim import FusedAdam

Batch = Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]


def prepare_batch(model_input, gt, device="cuda:0"):
    """utils.py:500-505 (+ sanitize :487-493): move to the device, nan_to_num float tensors."""
    def mv(v):
        v = v.to(device, non_blocking=True)
        return torch.nan_to_num(v) if v.is_floating_point() else v
    return {k: mv(v) for k, v in model_input.items()}, {k: mv(v) for k, v in gt.items()}


class CompactBatch:
    """Host batch in the compact wire format (18 B/point instead of the reference dict's 33 B/point):
    coords_frame (N,3) f32, sample_frame_ids as int16, sdf (N,) f32 and -- only when they are not all ones --
    weights (N,) f32.  `sdf_valid` / `sdf_signs` are NOT shipped: the reference's datasets define them as
    |sdf| < trunc and +-1 beyond +-trunc (sdf_rgbd.py:452-455) and `miso_expand_batch` rebuilds them (and the
    int64 ids) on the device, bit-identically.  Build with `CompactBatch.from_reference` -- it verifies that the
    masks of the given batch really are those functions of the sdf and refuses otherwise."""

    def __init__(self, coords, ids16, sdf, weights, trunc_dist):
        self.coords, self.ids16, self.sdf, self.weights, self.trunc_dist = coords, ids16, sdf, weights, float(trunc_dist)

    @staticmethod
    def from_reference(model_input, gt, trunc_dist, pin=True):
        coords = model_input["coords_frame"][0].contiguous().float()
        ids = model_input["sample_frame_ids"][0, :, 0]
        sdf = gt["sdf"][0, :, 0].contiguous().float()
        if int(ids.max()) > 32767 or int(ids.min()) < 0:
            raise ValueError("compact batches carry keyframe ids as int16")
        valid = torch.abs(sdf) < trunc_dist
        signs = torch.zeros_like(sdf)
        signs[sdf < -trunc_dist] = -1
        signs[sdf > trunc_dist] = 1
        if not torch.equal(valid, gt["sdf_valid"][0, :, 0].bool()) or not torch.equal(signs, gt["sdf_signs"][0, :, 0].float()):
            raise ValueError("sdf_valid / sdf_signs of this batch are not the dataset's functions of sdf and "
                             "trunc_dist; ship the reference format instead")
        w = model_input["weights"][0, :, 0].contiguous().float()
        w = None if bool((w == 1).all()) else w
        # one contiguous host buffer [coords | sdf | weights? | ids16] so a batch is ONE host->device copy; the typed
        # tensors below are views into it
        N = coords.shape[0]
        nbytes = 12 * N + 4 * N + (4 * N if w is not None else 0) + 2 * N
        buf = torch.empty(nbytes, dtype=torch.uint8)
        if pin:
            buf = buf.pin_memory()
        views = CompactBatch.views_of(buf, N, w is not None)
        views["coords"].copy_(coords)
        views["sdf"].copy_(sdf)
        views["ids16"].copy_(ids.to(torch.int16))
        if w is not None:
            views["weights"].copy_(w)
        cb = CompactBatch(views["coords"], views["ids16"], views["sdf"], views.get("weights"), trunc_dist)
        cb.buffer = buf
        return cb

    @staticmethod
    def views_of(buf: torch.Tensor, N: int, has_weights: bool) -> Dict[str, torch.Tensor]:
        """Typed views into a packed [coords | sdf | weights? | ids16] byte buffer (host or device)."""
        o = 0
        out = {"coords": buf[o:o + 12 * N].view(torch.float32).view(N, 3)}
        o += 12 * N
        out["sdf"] = buf[o:o + 4 * N].view(torch.float32)
        o += 4 * N
        if has_weights:
            out["weights"] = buf[o:o + 4 * N].view(torch.float32)
            o += 4 * N
        out["ids16"] = buf[o:o + 2 * N].view(torch.int16)
        return out

    def tensors(self):
        d = {"coords": self.coords, "ids16": self.ids16, "sdf": self.sdf}
        if self.weights is not None:
            d["weights"] = self.weights
        return d


def expand_compact_on_device(dev_tensors: Dict[str, torch.Tensor], trunc_dist: float, cache: dict) -> Batch:
    """Device-side inverse of CompactBatch.from_reference: one `miso_expand_batch` launch; output buffers are
    allocated once per shape (`cache`) and laid out as the reference's batch dict."""
    from . import _lib
    lib = _lib.load()
    coords, ids16, sdf = dev_tensors["coords"], dev_tensors["ids16"], dev_tensors["sdf"]
    N, dev = coords.shape[0], coords.device
    key = (N, dev)
    if cache.get("key") != key:
        cache.clear()
        cache.update(key=key, ids=torch.empty((1, N, 1), dtype=torch.int64, device=dev),
                     valid=torch.empty((1, N, 1), dtype=torch.uint8, device=dev),
                     sign=torch.empty((1, N, 1), dtype=torch.float32, device=dev),
                     ones=torch.ones((1, N, 1), dtype=torch.float32, device=dev))
    with torch.cuda.device(dev):
        _lib.check(lib.miso_expand_batch(ids16.data_ptr(), sdf.data_ptr(), float(trunc_dist), N, cache["ids"].data_ptr(),
                                         cache["valid"].data_ptr(), cache["sign"].data_ptr(), _lib.stream_ptr(dev)),
                   "expand_batch")
    w = dev_tensors["weights"].view(1, N, 1) if "weights" in dev_tensors else cache["ones"]
    model_input = {"coords_frame": coords.view(1, N, 3), "sample_frame_ids": cache["ids"], "weights": w}
    gt = {"sdf": sdf.view(1, N, 1), "sdf_valid": cache["valid"].view(torch.bool), "sdf_signs": cache["sign"]}
    return model_input, gt


class HostBatchStager:
    """Double-buffered host -> device staging of (model_input, gt) batches on a copy stream.

    Device buffers are allocated once per slot (no allocator traffic, no record_stream bookkeeping in the
    loop); `stage()` enqueues the H2D copies of a batch into the free slot, `acquire()` makes the compute
    stream wait for them, `release()` marks the slot reusable once the step that consumed it has been
    enqueued.  With pinned host tensors (DataLoader(pin_memory=True)) the copy of batch i+1 overlaps step i."""

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None] * slots
        self.ready = [None] * slots      # copy-stream event: data landed
        self.free = [None] * slots       # compute-stream event: consumer enqueued
        self.h2d_bytes = 0
        self._next = 0
        self.compact = [None] * slots    # trunc_dist of a CompactBatch staged in the slot (None: reference format)
        self._expand_cache = [dict() for _ in range(slots)]

    def _buffers(self, slot, batch):
        bufs = self.slots[slot]
        if isinstance(batch, CompactBatch):
            batch = ({"packed": batch.buffer} if getattr(batch, "buffer", None) is not None else batch.tensors(), {})
        model_input, gt = batch
        ok = bufs is not None and all(k in bufs[0] and bufs[0][k].shape == v.shape and bufs[0][k].dtype == v.dtype
                                      for k, v in model_input.items()) and \
            all(k in bufs[1] and bufs[1][k].shape == v.shape and bufs[1][k].dtype == v.dtype for k, v in gt.items())
        if not ok:
            bufs = ({k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in model_input.items()},
                    {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in gt.items()})
            self.slots[slot] = bufs
        return bufs

    def stage(self, batch: Batch) -> int:
        """Enqueue the async copy of `batch` (CPU tensors, ideally pinned) into the next slot; returns it."""
        slot = self._next
        self._next = (self._next + 1) % len(self.slots)
        bufs = self._buffers(slot, batch)
        self.compact[slot] = (batch.trunc_dist, batch.coords.shape[0], batch.weights is not None) \
            if isinstance(batch, CompactBatch) else None
        if isinstance(batch, CompactBatch):
            batch = ({"packed": batch.buffer} if getattr(batch, "buffer", None) is not None else batch.tensors(), {})
        nbytes = 0
        with torch.cuda.stream(self.stream):
            if self.free[slot] is not None:
                self.stream.wait_event(self.free[slot])
            for src, dst in zip(batch, bufs):
                for k, v in src.items():
                    dst[k].copy_(v, non_blocking=True)
                    nbytes += v.numel() * v.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[slot] = ev
        self.h2d_bytes = nbytes
        return slot

    def acquire(self, slot: int) -> Batch:
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        if self.compact[slot] is not None:
            trunc, n, has_w = self.compact[slot]
            dev = self.slots[slot][0]
            if "packed" in dev:
                dev = CompactBatch.views_of(dev["packed"], n, has_w)
            return expand_compact_on_device(dev, trunc, self._expand_cache[slot])
        return self.slots[slot]

    def release(self, slot: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev


class GridTrainer:
    """trainer.py:370-491 for a GridNet on the fused path."""

    def __init__(self, cfg: dict, model: GridNet, loss_func: MisoLossMapping, batch_fn: Callable[[int], Batch],
                 device="cuda:0"):
        self.cfg = cfg
        self.model = model
        self.loss_func = loss_func
        self.batch_fn = batch_fn
        self.device = device
        self.epochs = cfg.get("epochs", 50)
        self.lr = cfg.get("learning_rate", 1e-3)
        self.relchange_tol = cfg.get("relchange_tol", 0)
        self.max_epochs_in_level = cfg.get("max_epochs_in_level", 100)
        self.grid_training_mode = cfg.get("grid_training_mode", "coordinate+joint")
        if cfg.get("optimizer", "adam") != "adam":
            raise NotImplementedError("fused trainer implements optimizer: adam (the shipped configs)")
        # cuda_graph: Adam keeps its step counters on the device and `graphed_train_step` replays one captured graph
        # per (batch buffers, level schedule state) instead of issuing the step's launches from Python
        self.cuda_graph = bool(cfg.get("cuda_graph", False))
        self._graphs = {}
        self.train_dict = {"loss": []}
        self.total_steps = 0
        self.set_optimizer()

    def set_optimizer(self):
        m = self.model
        self.level_optimizers: List[FusedAdam] = []
        if self.grid_training_mode != "joint":
            for level in range(m.num_levels):
                self.level_optimizers.append(FusedAdam(m.params_at_level(level), lr=self.lr, device_step=self.cuda_graph))
        self.joint_optimizer = FusedAdam(list(m.parameters()), lr=self.lr, device_step=self.cuda_graph)
        self.reset_convergence_check()
        if self.grid_training_mode in ("coordinate", "coordinate+joint"):
            self.active_level = 0
            self.optimizer = self.level_optimizers[0]
        elif self.grid_training_mode == "joint":
            self.active_level = m.num_levels
            self.optimizer = self.joint_optimizer
        else:
            raise ValueError(f"Invalid grid training mode: {self.grid_training_mode}")

    def reset_convergence_check(self):
        self.params_prev = None
        self.relchange = np.inf
        self.epochs_in_level = 0

    def active_levels(self):
        return None if self.active_level >= self.model.num_levels else {self.active_level}

    def pre_epoch(self, epoch):
        """trainer.py:455-480."""
        if self.relchange < self.relchange_tol or self.epochs_in_level >= self.max_epochs_in_level:
            if self.active_level < self.model.num_levels:
                self.train_dict[f"level{self.active_level}_last_epoch"] = epoch
                self.active_level += 1
                if self.active_level >= self.model.num_levels:
                    if self.grid_training_mode == "coordinate+joint":
                        self.optimizer = self.joint_optimizer
                else:
                    self.optimizer = self.level_optimizers[self.active_level]
                self.reset_convergence_check()
        self.epochs_in_level += 1

    def train_step(self, model_input, gt, n_total=0, allreduce=None) -> torch.Tensor:
        """loss.compute -> backward -> optimizer.step (trainer.py:209-217) as fused launches.  Returns the
        (4,) device tensor [sdf, fs, eik, total]; nothing is synchronised.  Point-sharded multi-GPU fit:
        pass the global batch size and `miso_b200.dist.allreduce_sum_`; the dense grid gradients (and the
        loss terms) are summed over ranks before the identical Adam update on every rank."""
        terms = self.loss_func.step_into_grads(self.model, model_input, gt, self.active_levels(), n_total=n_total,
                                               count_allreduce=allreduce)
        if allreduce is not None:
            allreduce([p.grad for p in self.optimizer.params if p.grad is not None] + [terms])
        if self.cuda_graph:
            self.optimizer.step(gate=terms[3:4])    # NaN total (e.g. a keyframe without a pose): update skipped on the device
        else:
            self.optimizer.step()
        self.total_steps += 1
        return terms

    def graphed_train_step(self, model_input, gt) -> torch.Tensor:
        """`train_step` on device batch tensors that live at FIXED addresses (a resident batch, a staging slot), as ONE
        CUDA-graph launch: the first call for a given (buffers, active level, optimizer) runs eagerly (module loading,
        first-use allocations), the second captures the step and replays it, later calls only replay.  Every call is
        exactly one training step.  The returned (4,) loss tensor is the graph's static output: read or copy it before
        the next replay of the same graph."""
        if not self.cuda_graph:
            raise RuntimeError("GridTrainer was built without cfg['cuda_graph'] = True (Adam needs device-side step counters)")
        key = (tuple(v.data_ptr() for v in model_input.values()), tuple(v.data_ptr() for v in gt.values()),
               self.active_level, id(self.optimizer))
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "warm"
            return self.train_step(model_input, gt)
        if entry == "warm":
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                terms = self.train_step(model_input, gt)
            self.total_steps -= 1           # the capture itself ran nothing
            entry = self._graphs[key] = (graph, terms)
        graph, terms = entry
        graph.replay()
        self.total_steps += 1
        return terms

    def train_epoch(self, epoch):
        model_input, gt = self.batch_fn(epoch)
        model_input, gt = prepare_batch(model_input, gt, self.device)
        terms = self.train_step(model_input, gt)
        self.train_dict["loss"].append(terms)
        return terms

    def train_host_batches(self, batches, loss_sink: Optional[torch.Tensor] = None):
        """Run one fused step per host batch with the copy of batch i+1 overlapping step i (the pipelined
        form of train_epoch for batches that live in pinned host memory).  When `loss_sink` (a pinned
        (len, 4) tensor) is given, every step's loss terms are read back device->host asynchronously."""
        if not hasattr(self, "_stager"):
            self._stager = HostBatchStager(self.device)
        st = self._stager
        it = iter(batches)
        first = next(it, None)
        if first is None:
            return
        slot = st.stage(first)
        i = 0
        while slot is not None:
            nxt = next(it, None)
            nslot = st.stage(nxt) if nxt is not None else None
            model_input, gt = st.acquire(slot)
            terms = self.graphed_train_step(model_input, gt) if self.cuda_graph else self.train_step(model_input, gt)
            st.release(slot)
            if loss_sink is not None:
                loss_sink[i].copy_(terms, non_blocking=True)
            slot = nslot
            i += 1

    def train(self):
        for epoch in range(self.epochs):
            self.pre_epoch(epoch)
            self.train_epoch(epoch)
            if self.relchange_tol > 0:
                self.relchange = self.relative_param_change(self.model.params_at_level(self.active_level))
        return self.train_dict

    def relative_param_change(self, params):
        cur = [p.detach().clone() for p in params if p.requires_grad]
        if self.params_prev is None:
            self.params_prev = cur
            return np.inf
        num = sum(torch.sum((c - p) ** 2) for c, p in zip(cur, self.params_prev))
        den = sum(torch.sum(p ** 2) for p in self.params_prev)
        self.params_prev = cur
        return torch.sqrt(num / den).item()


class Mapper:
    """slam/mapper.py:27-98: builds the mapping loss + trainer on one GridNet; `mapping(kfs, iterations,
    level_iterations)` unlocks features, locks poses, trains."""

    def __init__(self, model: GridNet, batch_fn: Callable[[int], Batch], cfg: dict, device="cuda:0"):
        self.model = model
        self.cfg = cfg
        self.device = device
        cm = cfg["mapping"]
        self.loss_func = MisoLossMapping(loss_type=cm["loss_type"], weight_sdf=cm["weight_sdf"],
                                         weight_eik=cm["weight_eik"], weight_fs=cm["weight_fs"],
                                         trunc_dist=cm["trunc_dist"], finite_diff_eps=cm["finite_diff_eps"],
                                         grad_method=cm["grad_method"], eik_trunc_dist=cm.get("eik_trunc_dist"))
        self.batch_fn = batch_fn

    def mapping(self, kfs=None, iterations=300, level_iterations=50):
        grid = self.model
        grid.unlock_feature()
        grid.lock_pose()
        cfg_train = dict(self.cfg.get("train", {}))
        cfg_train["epochs"] = iterations
        cfg_train["max_epochs_in_level"] = level_iterations
        cfg_train["learning_rate"] = self.cfg["mapping"]["learning_rate"]
        trainer = GridTrainer(cfg_train, grid, self.loss_func, self.batch_fn, device=self.device)
        return trainer.train()
