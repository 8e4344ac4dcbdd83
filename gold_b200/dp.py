"""Batch-sharded data parallelism over the C ABI (BASELINE configs[3]).

One process per GPU (torchrun).  Every rank holds the full parameters and Adam state,
trains `batch_size` rows of a `global_batch = batch_size * world` step, and the flat
fp32 gradient (+ the loss) is summed with ONE all-reduce per step inside goldq_learn,
so all ranks apply the identical update.  torch.distributed is only the side channel
that distributes the 128-byte communicator id.
"""
import numpy as np

from . import _capi as G


def shard_indices(global_idx, rank, world):
    """Rank r trains rows [r*B/R, (r+1)*B/R) of the global index list (contiguous slice)."""
    global_idx = np.asarray(global_idx, np.int32)
    if global_idx.size % world:
        raise ValueError("global batch must divide evenly over the ranks")
    per = global_idx.size // world
    return global_idx[rank * per:(rank + 1) * per]


def scale_local_grads(local_grads, local_rows, global_rows):
    """A gradient computed as a mean over the LOCAL rows, rescaled to this rank's share of
    the mean over the GLOBAL batch (loss = Sum / (B_global * A), goro loss.go:28-42)."""
    return np.asarray(local_grads) * (float(local_rows) / float(global_rows))


def init_group(handle, dist):
    """Create the handle's communicator; `dist` is an initialised torch.distributed."""
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [G.Handle.dp_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    handle.dp_init(uid[0], rank, world)
    return rank, world
