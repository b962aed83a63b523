"""Large single-grid fit over several GPUs (BASELINE.json configs[3]) -- domain decomposition.

The reference's loss is a mean over the batch (grid_opt/loss.py:634), so the batch can be split across ranks any way
that lets the per-rank terms and gradients sum.  Two splits are implemented:

  * `GridTrainer.train_step(n_total=, allreduce=)`  (miso_b200.trainer) -- contiguous point chunks, the dense grid
    gradients all-reduced before a replicated Adam step.  This is the split the north star names; its cost per step
    is the all_reduce of the whole fine level (324 MB for the NCD quad grid) plus a full Adam sweep on every rank,
    neither of which shrinks with the number of GPUs: measured 0.94x at 2 GPUs (profiles/r01_ncd_point_sharded.json).

  * `SlabShardedFit` (this file) -- the batch is split by WHERE the samples fall.  The largest level is cut into
    contiguous ranges of planes along ONE axis -- z or y, whichever balances this scene better (an outdoor LiDAR batch
    has a third of its samples in the two z-planes of the ground) -- and is stored with that axis slowest, so a range
    of planes is one contiguous piece of the level, of its gradient and of its Adam moments (the kernels take arbitrary
    strides; only channels must stay innermost); every rank reads the whole batch (it is
    replicated: in the reference every process would load the same dataset) and keeps the samples whose cell of that
    level starts in its planes (`miso_slab_select`, device-side compaction, no host sync), runs the fused step on
    them with the GLOBAL batch size as denominator, and owns the Adam update of its planes.  A sample touches planes
    [z, z+1], so what crosses ranks per step is ONE plane of gradients up and one plane of parameters down per
    neighbour (0.8 MB for the NCD quad grid instead of 324 MB), plus ONE all_reduce of the small replicated levels'
    gradients and the four loss terms.  On one node (`halo="p2p"`) the two planes do not travel as messages at all:
    the boundary plane's Adam kernel (`miso_adam_step_halo`) reads the neighbour's gradient plane and writes the
    neighbour's parameter plane directly over NVLink peer memory (buffers mapped with CUDA IPC), the interior planes'
    Adam overlaps it on a second stream, and the neighbour is released by a counter in peer memory
    (`miso_peer_signal` / `miso_peer_wait`) instead of a second collective; `halo="nccl"` keeps batched isend/irecv.
    Slab boundaries come from a two-phase cost model (`calibrate`): max samples per rank + c * max parameters per rank.

Both reproduce the single-GPU step up to the order of the float32 atomics.  `gather_model` re-assembles the full level
on every rank (checkpointing / meshing).
"""
import ctypes as C
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from . import dist as mdist
from . import field as _field
from .loss import MisoLossMapping, _flat_f32, _flat_u8, mapping_step_raw
from .optim import FusedAdam


# ------------------------------------------------------------------------------------------------
# host logic (device-agnostic torch: exercised by the gloo tests)
# ------------------------------------------------------------------------------------------------
def slab_bounds_from_histogram(hist: torch.Tensor, world: int) -> List[int]:
    """world+1 plane indices b[0]=0 <= ... <= b[world]=Z such that the slabs [b[r], b[r+1]) hold about the same share
    of `hist` (samples per z-plane) and every slab has at least one plane."""
    Z = int(hist.numel())
    if world > Z:
        raise ValueError(f"{world} ranks for a level of {Z} planes")
    cum = torch.cumsum(hist.double().cpu(), 0)
    total = float(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        z = int(torch.searchsorted(cum, torch.tensor(target, dtype=torch.float64)).item()) + 1
        z = max(z, bounds[-1] + 1)              # at least one plane per slab
        z = min(z, Z - (world - r))             # leave a plane for every later slab
        bounds.append(z)
    bounds.append(Z)
    return bounds


def plane_of_points(z_world: torch.Tensor, zmin: float, zmax: float, Z: int) -> torch.Tensor:
    """z-plane that owns a sample: floor of the level's z index, clamped into [0, Z-1] (NaN -> 0).  Torch restatement
    of the arithmetic in `slab_select_kernel` / `make_cell` (normalize, unnormalize with align_corners=False, floor);
    used for the calibration histogram and by the CPU tests."""
    zn = 2 * (z_world - zmin) / (zmax - zmin) - 1
    iz = ((zn + 1) * Z - 1) / 2
    return torch.nan_to_num(torch.floor(iz), nan=0.0).clamp(0, Z - 1).long()


def exchange_halo_planes(send_up: Optional[torch.Tensor], recv_from_below: Optional[torch.Tensor],
                         rank: int, world: int, group=None):
    """One step of the nearest-neighbour exchange along the slab axis: rank r sends `send_up` to r+1 and receives into
    `recv_from_below` from r-1 (each None at the ends).  Batched P2P (NCCL on the box, gloo in the tests)."""
    ops = []
    if send_up is not None and rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, send_up, rank + 1, group))
    if recv_from_below is not None and rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv_from_below, rank - 1, group))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()


def exchange_halo_planes_down(send_down: Optional[torch.Tensor], recv_from_above: Optional[torch.Tensor],
                              rank: int, world: int, group=None):
    """The opposite direction: rank r sends to r-1, receives from r+1."""
    ops = []
    if send_down is not None and rank > 0:
        ops.append(dist.P2POp(dist.isend, send_down, rank - 1, group))
    if recv_from_above is not None and rank + 1 < world:
        ops.append(dist.P2POp(dist.irecv, recv_from_above, rank + 1, group))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()


# ------------------------------------------------------------------------------------------------
# the trainer
# ------------------------------------------------------------------------------------------------
class SlabShardedFit:
    """Adam fit of one GridNet on a replicated batch, the largest level cut into z-slabs over the ranks."""

    def __init__(self, model, loss: MisoLossMapping, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 rank: Optional[int] = None, world: Optional[int] = None, bounds: Optional[Sequence[int]] = None,
                 halo: str = "auto"):
        """`halo`: "p2p" -- the boundary plane's Adam reads the neighbour's gradient plane and writes the neighbour's
        parameter plane directly over NVLink peer memory (one node, CUDA IPC; miso_adam_step_halo), the interior
        planes' Adam overlaps it on a second stream; "nccl" -- batched isend/irecv of the two planes (any topology);
        "auto" -- p2p when a NCCL process group spans more than one rank."""
        r, w = mdist.world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.model, self.loss, self.lr, self.betas, self.eps = model, loss, float(lr), betas, float(eps)
        feats = model.level_tensors()
        self.slab_level = max(range(len(feats)), key=lambda l: feats[l].numel())
        f = feats[self.slab_level]
        if f.stride(1) != 1:
            raise RuntimeError("SlabShardedFit needs channels innermost (channels_last_3d or a permutation of it)")
        if halo not in ("auto", "p2p", "nccl"):
            raise ValueError(f"halo must be 'auto', 'p2p' or 'nccl', not {halo!r}")
        live = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.p2p = halo == "p2p" or (halo == "auto" and live and f.is_cuda and dist.get_backend() == "nccl"
                                     and self._one_node())
        if self.p2p and not (live and self.world == dist.get_world_size() and self.rank == dist.get_rank()):
            raise RuntimeError("halo='p2p' needs an initialised process group whose ranks are the slab owners")
        self._peer, self._side, self._side2, self._flat_buf = None, None, None, None
        self._sets, self._presel = None, None
        self.sync = torch.zeros(4, dtype=torch.int32, device=f.device) if self.p2p else None
        self.mark = None     # optional callable(name): phase boundaries of a step (benchmarks/slab_breakdown.py)
        self.axis = 2                        # slab axis in (x, y, z) numbering; z is slowest in channels_last_3d
        self._configure_axis(2 if f.stride(2) > f.stride(3) else 1)
        self.bounds = list(bounds) if bounds is not None else [round(self.Z * k / self.world) for k in range(self.world + 1)]
        self._set_slab()
        self.step_count = 0
        self.other = FusedAdam([p for l, p in enumerate(feats) if l != self.slab_level and p.requires_grad],
                               lr=lr, betas=betas, eps=eps, device_step=True)
        self._bufs = None

    @staticmethod
    def _one_node() -> bool:
        """True when every rank of the process group runs on this host (CUDA IPC peer mappings need that).  Collective."""
        import socket
        names = [None] * dist.get_world_size()
        dist.all_gather_object(names, socket.gethostname())
        return len(set(names)) == 1

    # ---- slabs ---------------------------------------------------------------------------------------
    def _configure_axis(self, axis: int):
        """Slab axis 2 (z) or 1 (y): `Z` = planes along it, `plane_elems` = floats per plane, [zmin, zmax] its bound."""
        f = self.model.level_tensors()[self.slab_level]
        self.axis = axis
        self.Z = f.shape[2] if axis == 2 else f.shape[3]
        self.plane_elems = f.numel() // self.Z
        self.zmin, self.zmax = self.model._bound_host[2 * axis], self.model._bound_host[2 * axis + 1]

    def _relayout(self, axis: int):
        """Store the slab level (and its gradient) with `axis` slowest: logical shape (1,C,Z,Y,X) and values unchanged."""
        p = self.model.features[self.slab_level].feature
        order = (0, 2, 3, 4, 1) if axis == 2 else (0, 3, 2, 4, 1)        # physical order: slab axis, other, x, channels
        back = (0, 4, 1, 2, 3) if axis == 2 else (0, 4, 2, 1, 3)
        with torch.no_grad():
            p.data = p.data.permute(order).contiguous().permute(back)
            if p.grad is not None:
                p.grad = p.grad.permute(order).contiguous().permute(back)
        self._configure_axis(axis)

    def restore_layout(self):
        """Back to channels_last_3d (z slowest), e.g. before saving a checkpoint."""
        if self.axis != 2:
            self._relayout(2)

    def _set_slab(self):
        self.zb, self.ze = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        dev = self.model.level_tensors()[self.slab_level].device
        self._peer = None
        # p2p halos: the first owned plane (it also receives the lower neighbour's gradient) has its own Adam launch
        self.zi = self.zb + 1 if (self.p2p and self.rank > 0) else self.zb
        if self.zi != self.zb:
            self.b_exp_avg = torch.zeros(self.plane_elems, dtype=torch.float32, device=dev)
            self.b_exp_avg_sq = torch.zeros(self.plane_elems, dtype=torch.float32, device=dev)
            self.b_step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self.b_scalars = torch.zeros(3, dtype=torch.float32, device=dev)
        n = (self.ze - self.zi) * self.plane_elems
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        # "ever touched" bitmap of the slab (one bit per 4-float voxel): never-touched voxels cost one gradient read
        self.touched = torch.zeros((n // 4 + 31) // 32, dtype=torch.int32, device=dev)
        # step counter + bias-correction scalars on the device: every step enqueues the same launches (CUDA graph)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.scalars = torch.zeros(3, dtype=torch.float32, device=dev)

    def _flat(self, t: torch.Tensor) -> torch.Tensor:
        """(1,C,Z,Y,X) tensor stored slab-axis-slowest -> flat (planes, plane_elems) view of its memory."""
        order = (0, 2, 3, 4, 1) if self.axis == 2 else (0, 3, 2, 4, 1)
        return t.detach().permute(order).reshape(self.Z, self.plane_elems)

    def calibrate(self, model_input: dict, voxel_cost: float = 1.0 / 56.0, axes=None):
        """Choose slab boundaries that balance the per-step work of this batch over the ranks (identical on every rank:
        the batch is replicated).  A step is two phases separated by a collective -- the fused step kernel (cost ~ the
        rank's sample count, ~0.2 ns per sample on B200) and the Adam sweep of the slab (`voxel_cost` sample-equivalents
        per parameter: it streams the slab's gradient and, for touched voxels, p / m / v, ~0.0034 ns per float on the NCD
        quad grid) -- so the step costs  max_r samples_r + voxel_cost * max_r params_r.  Candidates: both slab axes
        (z, y) x cuts that equalise `samples + w * voxel_cost * params` per slab for a few weights w; the candidate with
        the smallest modelled step wins, the level is re-laid with that axis slowest.  Resets the slab's Adam moments;
        call before the first step."""
        coords = model_input["coords_frame"][0]
        ids = model_input["sample_frame_ids"][0, :, 0]
        R, t, _ = self.loss.frame_table(self.model)
        f = self.model.level_tensors()[self.slab_level]
        best = None
        for axis in ((2, 1) if axes is None else axes):
            n_planes = f.shape[2] if axis == 2 else f.shape[3]
            if n_planes < self.world:
                continue
            lo, hi = self.model._bound_host[2 * axis], self.model._bound_host[2 * axis + 1]
            w = torch.einsum("nj,nj->n", R[ids][:, axis, :], coords) + t[ids][:, axis, 0]
            samples = torch.bincount(plane_of_points(w, lo, hi, n_planes), minlength=n_planes).double().cpu()
            per_plane = f.numel() // n_planes
            cum = torch.cat([torch.zeros(1, dtype=torch.float64), torch.cumsum(samples, 0)])
            for weight in (1.0, 0.5, 0.25, 0.0):
                bounds = slab_bounds_from_histogram(samples + weight * voxel_cost * per_plane, self.world)
                step = max(float(cum[b1] - cum[b0]) for b0, b1 in zip(bounds[:-1], bounds[1:])) \
                    + voxel_cost * per_plane * max(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:]))
                if best is None or step < 0.98 * best[0]:     # z (the native layout) unless y is clearly better
                    best = (step, axis, bounds)
        _, axis, self.bounds = best
        if axis != self.axis:
            self._relayout(axis)
        self._set_slab()
        return self.bounds

    # ---- one step --------------------------------------------------------------------------------------
    def _buffers(self, N, dev, need_w=True, which=0):
        """Compaction targets of the slab selection; two sets, so the next batch can be selected while this one trains."""
        if self._sets is None or self._sets[0]["N"] != N:
            def make():
                return {"N": N, "x": torch.empty((N, 3), dtype=torch.float32, device=dev),
                        "ids": torch.empty(N, dtype=torch.int64, device=dev),
                        "sdf": torch.empty(N, dtype=torch.float32, device=dev),
                        "valid": torch.empty(N, dtype=torch.uint8, device=dev),
                        "sign": torch.empty(N, dtype=torch.float32, device=dev),
                        "w": torch.empty(N, dtype=torch.float32, device=dev),
                        "count": torch.zeros(1, dtype=torch.int32, device=dev)}
            self._sets = [make(), make()]
            self._halo = torch.empty(self.plane_elems, dtype=torch.float32, device=dev)
            self._presel = None
        out = self._sets[which]
        out["halo"] = self._halo
        return out

    @staticmethod
    def _batch_key(model_input, gt):
        return tuple(v.data_ptr() for v in model_input.values()) + tuple(v.data_ptr() for v in gt.values())

    def _select(self, model_input: dict, gt: dict, which: int):
        """Enqueue the slab selection of a batch into buffer set `which` on the current stream."""
        lib, m, L = _lib.load(), self.model, self.loss
        coords = _field._prep_x(model_input["coords_frame"][0])
        ids = model_input["sample_frame_ids"][0, :, 0]
        sdf, valid, sign = _flat_f32(gt["sdf"][0]), _flat_u8(gt["sdf_valid"][0]), _flat_f32(gt["sdf_signs"][0])
        w = _flat_f32(model_input["weights"][0])
        N, dev = coords.shape[0], coords.device
        b = self._buffers(N, dev, True, which)
        fr = L._frames(m, ids).struct()
        with torch.cuda.device(dev):
            _lib.check(lib.miso_slab_select(
                C.byref(fr), coords.data_ptr(), N, float(self.zmin), float(self.zmax), self.Z, self.axis, self.zb, self.ze,
                sdf.data_ptr(), valid.data_ptr(), sign.data_ptr(), w.data_ptr(), b["x"].data_ptr(), b["ids"].data_ptr(),
                b["sdf"].data_ptr(), b["valid"].data_ptr(), b["sign"].data_ptr(), b["w"].data_ptr(),
                b["count"].data_ptr(), _lib.stream_ptr(dev)), "slab_select")
        return b

    def step(self, model_input: dict, gt: dict, prefetch=None) -> torch.Tensor:
        """One fit step on the (replicated, device-resident) batch.  Returns the GLOBAL (4,) loss terms
        [sdf, fs, eik, total]; nothing is synchronised with the host.  `prefetch = (model_input, gt)` of the NEXT step:
        its slab selection is enqueued on a second stream behind this step's kernel, where it overlaps the collectives
        and the Adam sweeps; the next `step` call on those tensors then starts with the fused kernel."""
        lib, m, L = _lib.load(), self.model, self.loss
        if not L._fused_ok(m) or (L.weight_eik > 0 and L.grad_method != "autograd"):
            raise RuntimeError("SlabShardedFit runs the fused analytic step (fixed decoder, locked poses)")
        ids = model_input["sample_frame_ids"][0, :, 0]
        sdf = _flat_f32(gt["sdf"][0])
        N, dev = model_input["coords_frame"].shape[1], model_input["coords_frame"].device
        self._buffers(N, dev)
        R, t, _ = L.frame_table(m)
        stream = _lib.stream_ptr(dev)
        key = self._batch_key(model_input, gt)
        if self._presel is not None and self._presel[0] == key:
            cur = self._presel[1]
            b = self._buffers(N, dev, True, cur)
        else:
            cur = 0
            b = self._select(model_input, gt, 0)
        self._bufs = b
        self._m("select")
        feats = m.level_tensors()
        loss_out = None
        if self.p2p and self.world > 1:
            loss_out = self._pack_replicated_grads(feats)
            if self.rank + 1 < self.world:
                # the upper neighbour's boundary-plane Adam of the previous step wrote my halo plane of parameters and
                # consumed my halo plane of gradients: hold the step kernel until it has signalled
                with torch.cuda.device(dev):
                    _lib.check(lib.miso_peer_wait(self.sync.data_ptr(), stream), "peer_wait")
        grads = []
        for f in feats:
            if f.requires_grad and f.grad is None:
                f.grad = torch.zeros_like(f)
            grads.append(f.grad if f.requires_grad else None)
        cfg = L._step_cfg()
        cfg.pop("fd_eps", None)
        own_frames = _field.FramesSpec(b["ids"], R, t)      # the selection already mapped keyframe ids to table rows
        # the |gt| < eik_trunc count runs over the FULL batch inside mapping_step_raw when gt_sdf_count is given
        terms = mapping_step_raw(feats, grads, m.fused_spec(), own_frames, b["x"], b["sdf"], b["valid"], b["sign"], b["w"],
                                 n_total=N, n_device=b["count"], count_on=sdf, loss_out=loss_out, **cfg)
        self._m("step_kernel")
        self._presel = None
        if prefetch is not None:
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream(dev)
            self._side2.wait_stream(main)       # behind the step kernel (it fills the GPU; buffer set 1 - cur is free by then)
            with torch.cuda.stream(self._side2):
                self._select(prefetch[0], prefetch[1], 1 - cur)
            self._presel = (self._batch_key(*prefetch), 1 - cur)
        self._exchange_and_update(feats, grads, terms, b)
        if prefetch is not None:
            torch.cuda.current_stream(dev).wait_stream(self._side2)
        return terms

    def _m(self, name):
        if self.mark is not None:
            self.mark(name)

    def _adam_interior(self, feats, grads):
        sl = self.slab_level
        n = (self.ze - self.zi) * self.plane_elems
        if n <= 0:
            return
        off = self.zi * self.plane_elems * 4
        dev = feats[sl].device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().miso_adam_step_dev(
                feats[sl].data_ptr() + off, grads[sl].data_ptr() + off, self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr(), self.touched.data_ptr(), n, self.lr, float(self.betas[0]),
                float(self.betas[1]), self.eps, self.step_dev.data_ptr(), self.scalars.data_ptr(), None, 1,
                _lib.stream_ptr(dev)), "adam_step")

    def _pack_replicated_grads(self, feats):
        """p2p mode: the gradients of the replicated (non-slab) levels and the step's 4 loss terms live in ONE flat
        buffer, so a single all_reduce sums them all (and orders the ranks).  Returns the loss-term view."""
        sl = self.slab_level
        rep = [f for l, f in enumerate(feats) if l != sl and f.requires_grad]
        key = tuple(f.data_ptr() for f in rep)
        if self._flat_buf is None or self._flat_buf[0] != key:
            total = sum(f.numel() for f in rep)
            flat = torch.zeros(total + 4, dtype=torch.float32, device=feats[sl].device)
            off = 0
            for f in rep:
                view = flat[off:off + f.numel()].as_strided(f.shape, f.stride())
                if f.grad is not None:
                    view.copy_(f.grad)
                f.grad = view
                off += f.numel()
            self._flat_buf = (key, flat, total)
        _, flat, total = self._flat_buf
        return flat[total:]

    def check_sync(self):
        """Raises if a neighbour wait timed out (a rank died or fell out of step).  Synchronises with the host."""
        if self.p2p and int(self.sync[2].item()) != 0:
            raise RuntimeError("SlabShardedFit: timed out waiting for the upper neighbour's boundary-plane update")

    def _open_peers(self, feats, grads):
        """Map the lower neighbour's gradient and parameter buffers of the slab level into this process (CUDA IPC).
        Collective (all_gather_object): every rank calls it at the same point, outside any graph capture."""
        lib, sl = _lib.load(), self.slab_level
        mine = []
        for t in (grads[sl], feats[sl], self.sync):
            h, off = (C.c_ubyte * 64)(), C.c_int64(0)
            _lib.check(lib.miso_ipc_export(t.data_ptr(), h, C.byref(off)), "ipc_export")
            mine.append((bytes(h), int(off.value)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self._peer = {"key": (grads[sl].data_ptr(), feats[sl].data_ptr()), "g": None, "p": None, "flag": None}
        if self.rank > 0:
            ptrs = []
            with torch.cuda.device(feats[sl].device):
                for hb, off in everyone[self.rank - 1]:
                    out = C.c_void_p()
                    _lib.check(lib.miso_ipc_import((C.c_ubyte * 64).from_buffer_copy(hb), off, C.byref(out)), "ipc_import")
                    ptrs.append(int(out.value))
            plane = self.zb * self.plane_elems * 4        # my first plane == the neighbour's halo plane `ze`
            self._peer["g"], self._peer["p"], self._peer["flag"] = ptrs[0] + plane, ptrs[1] + plane, ptrs[2]
        dist.barrier()

    def _exchange_and_update(self, feats, grads, terms, b):
        lib = _lib.load()
        sl, r, W = self.slab_level, self.rank, self.world
        coarse = [gr for l, gr in enumerate(grads) if l != sl and gr is not None]
        if W > 1 and self.p2p:
            dev = feats[sl].device
            if self._peer is None or self._peer["key"] != (grads[sl].data_ptr(), feats[sl].data_ptr()):
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("SlabShardedFit: run one eager step before capturing (peer buffers are mapped then)")
                self._open_peers(feats, grads)
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream(dev)
            # (1) every rank's step kernel is complete once this all_reduce returns: the neighbour's halo plane is final.
            # One buffer: the replicated levels' gradients and the loss terms (`terms` is its tail view)
            mdist.allreduce_sum_([self._flat_buf[1]])
            self._m("allreduce_coarse")
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self._adam_interior(feats, grads)          # touches planes [zi, ze) only: overlaps the halo work
            self.other.step()
            if r > 0:
                off = self.zb * self.plane_elems * 4
                with torch.cuda.device(dev):
                    _lib.check(lib.miso_adam_step_halo(
                        feats[sl].data_ptr() + off, grads[sl].data_ptr() + off, self.b_exp_avg.data_ptr(),
                        self.b_exp_avg_sq.data_ptr(), self.plane_elems, self._peer["g"], self._peer["p"], self.lr,
                        float(self.betas[0]), float(self.betas[1]), self.eps, self.b_step_dev.data_ptr(),
                        self.b_scalars.data_ptr(), _lib.stream_ptr(dev)), "adam_step_halo")
                    # (2) tell the lower neighbour: its halo plane of parameters is written, its halo plane of gradients
                    # consumed and cleared -- it waits for this before its next step kernel (miso_peer_wait in step())
                    _lib.check(lib.miso_peer_signal(self._peer["flag"], _lib.stream_ptr(dev)), "peer_signal")
            main.wait_stream(self._side)
            self._m("adam_interior|coarse+halo+signal")
            self.step_count += 1
            return
        g, p = self._flat(grads[sl]), self._flat(feats[sl])
        if W > 1:
            mdist.allreduce_sum_(coarse + [terms])
            # gradient halo: my samples also wrote plane `ze`, which rank r+1 owns
            exchange_halo_planes(g[self.ze] if self.ze < self.Z else None, b["halo"] if r > 0 else None, r, W)
            if r > 0:
                g[self.zb].add_(b["halo"])
            if self.ze < self.Z:
                g[self.ze].zero_()
        self.other.step()
        self.step_count += 1
        self._adam_interior(feats, grads)
        if W > 1:
            # parameter halo: the next step reads plane `ze` (owned and just updated by rank r+1)
            exchange_halo_planes_down(p[self.zb] if r > 0 else None, p[self.ze] if self.ze < self.Z else None, r, W)

    def graphed_step(self, model_input: dict, gt: dict, prefetch_same: bool = False):
        """Capture `step` on these (device-resident, fixed-address) batch tensors into a CUDA graph -- slab selection,
        fused step, the all_reduce, the halo exchange (peer-memory Adam or NCCL P2P) and the Adam sweeps -- after one eager
        step on a side stream (module loading, NCCL channel setup, peer mappings; it counts as a training step).  Returns
        `replay() -> loss terms`: one graph launch per step, which matters here because a step is ~15 short launches.
        `prefetch_same`: the tensors are a staging slot that holds the NEXT batch by the time a step's kernel has run
        (or a resident batch): every step also selects the slot's samples for the following step (two graphs, alternating
        between the two compaction buffer sets)."""
        pf = (model_input, gt) if prefetch_same else None
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.step(model_input, gt, prefetch=pf)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for _ in range(2 if prefetch_same else 1):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                terms = self.step(model_input, gt, prefetch=pf)
            graphs.append((graph, terms))
        self._graphs = graphs   # keep alive
        turn = [0]

        def replay():
            graph, terms = graphs[turn[0] % len(graphs)]
            turn[0] += 1
            graph.replay()
            return terms
        return replay

    @torch.no_grad()
    def gather_model(self):
        """Every rank ends up with the full slab level (one broadcast per slab from its owner)."""
        if self.world == 1:
            return
        self.check_sync()
        p = self._flat(self.model.level_tensors()[self.slab_level])
        for owner in range(self.world):
            dist.broadcast(p[self.bounds[owner]:self.bounds[owner + 1]], src=owner)
