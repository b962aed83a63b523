"""Per-launch time of one alignment iteration (bench workload, 16 submaps, 120 pairs) at both levels."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from miso_b200 import _lib  # noqa: E402
from miso_b200.align import AlignBatch, FusedPoseAligner  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    atlas = bench.build_align_atlas(dev)
    atlas.precompute_coordinates_for_alignment()
    pairs = [(s, d) for s in range(bench.ALIGN_SUBMAPS) for d in range(s + 1, bench.ALIGN_SUBMAPS)]
    lib = _lib.load()
    out = {}
    for level in (0, 1):
        b = AlignBatch(atlas, pairs, level, check_intersection=True)
        al = FusedPoseAligner(b, max_iters=100000)
        al.compose()
        b.update_intersections(al.poses24)
        stream = _lib.stream_ptr(dev)

        def align():
            _lib.check(lib.miso_align_batch(b.fields_dev.data_ptr(), b.num_fields, b.pairs_dev.data_ptr(), al.P, b.max_M,
                                            al.poses24.data_ptr(), al.out.data_ptr(), 0, stream), "align_batch")
        out[f"level{level}"] = {"compose_ms": timed(al.compose), "intersections_ms": timed(lambda: b.update_intersections(al.poses24)),
                                "align_batch_ms": timed(align), "whole_iteration_eager_ms": timed(al.iteration)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
