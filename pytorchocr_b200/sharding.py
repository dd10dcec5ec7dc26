"""Multi-GPU sharding of the post-processing path: one process per GPU, the batch is
split by image (detection) or by text line (CTC, dimension 1 of [T,B,C]); every rank post-processes
its own contiguous shard on its own device and only the small results (boxes / strings) are gathered
on the host. There is NO collective on the data path - images and text lines are independent
(R/pytocr/postprocess/db_postprocess.py:49, pse_postprocess.py:48, pan_postprocess.py:54,
rec_postprocess.py:40) - so NCCL is not involved; the gather below moves a few KB per image through
whatever process group the job already has (gloo on the host, or NCCL's object collectives).

The reference evaluates on rank 0 only (R/tools/program.py:331); `gather_results(..., dst=0)`
reproduces that hand-off: rank 0 receives the per-image results of the whole batch in item order.
"""


def shard_bounds(n_items, rank, world):
    """Contiguous split: rank g owns items [g*n/G, (g+1)*n/G) (sizes differ by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def shard_detection_inputs(maps, shape_list, rank, world):
    """maps [N,C,H,W] (tensor or array), shape_list [N,4] -> this rank's (maps, shape_list) views."""
    lo, hi = shard_bounds(len(maps), rank, world)
    return maps[lo:hi], shape_list[lo:hi]


def shard_ctc_inputs(preds, rank, world):
    """preds [T,B,C] as the CTC head emits it (rec_ctc_head.py:17-36) -> [T, B/G, C] view of this rank's
    lines (numpy input is [B,T,C] in the reference, rec_postprocess.py:80-82: sharded on dim 0)."""
    import numpy as np
    if isinstance(preds, np.ndarray):
        lo, hi = shard_bounds(preds.shape[0], rank, world)
        return preds[lo:hi]
    lo, hi = shard_bounds(preds.shape[1], rank, world)
    return preds[:, lo:hi]


def gather_results(local_results, group=None, dst=None):
    """Concatenates the per-item result lists of all ranks in rank (= item) order.
    dst=None: every rank gets the full list (all_gather_object); dst=k: only rank k does, the others
    get None (gather_object). Works without an initialised process group (single process)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(local_results)
    world = dist.get_world_size(group)
    if dst is None:
        parts = [None] * world
        dist.all_gather_object(parts, list(local_results), group=group)
    else:
        me = dist.get_rank(group)
        parts = [None] * world if me == dst else None
        dist.gather_object(list(local_results), parts, dst=dst, group=group)
        if me != dst:
            return None
    out = []
    for p in parts:
        out.extend(p)
    return out


class ShardedPostProcess(object):
    """Wraps one of the operators for data-parallel use: `op(outs_dict, shape_list)` is called with the
    FULL batch on every rank (or with `presharded=True` and the rank's own shard, as under
    DistributedSampler); each rank processes its shard on its own GPU; the results of the whole batch
    come back in item order on `dst` (default: every rank)."""

    def __init__(self, operator, group=None, dst=None, presharded=False):
        self.operator, self.group, self.dst, self.presharded = operator, group, dst, presharded

    def _rank_world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def __call__(self, outs, shape_list=None, **kwargs):
        rank, world = self._rank_world()
        if shape_list is None:      # CTC: outs = preds
            local = outs if self.presharded else shard_ctc_inputs(outs, rank, world)
            res = self.operator(local, **kwargs)
        else:
            maps = outs["maps"]
            if not self.presharded:
                maps, shape_list = shard_detection_inputs(maps, shape_list, rank, world)
            res = self.operator({"maps": maps}, shape_list, **kwargs) if len(maps) else []
        return gather_results(res, self.group, self.dst)
