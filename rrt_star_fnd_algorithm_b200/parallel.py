"""Multi-GPU plumbing: planners are independent, so ranks only meet once, after growth, to gather the
per-planner result records and to broadcast the globally best path. Works on any
torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

from ._structs import PLANNER_BYTES, PLANNER_DTYPE

_BEST_COST_COL = PLANNER_DTYPE.fields["best_cost"][1] // 8     # column of best_cost in a float64 view
_BEST_LEN_OFF = PLANNER_DTYPE.fields["best_len"][1] // 4       # column of best_len in an int32 view


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block of planner indices owned by `rank` (blocks differ by at most one planner)."""
    base, extra = divmod(n_total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_records(records: torch.Tensor, best_paths: torch.Tensor, group=None):
    """all_gather the [P, PLANNER_BYTES] uint8 records of every rank, pick the planner with the lowest
    best cost (lowest global index on ties) and broadcast its best path from the rank that owns it.

    Every rank must own the same number of planners. Returns (all_records [world*P, PLANNER_BYTES],
    global index of the best planner or -1 if none found a path, its path ids as an int32 tensor).
    """
    world = dist.get_world_size(group)
    P = records.shape[0]
    allrec = torch.empty((world * P, PLANNER_BYTES), dtype=torch.uint8, device=records.device)
    dist.all_gather_into_tensor(allrec, records.contiguous(), group=group)
    costs = allrec.view(torch.float64)[:, _BEST_COST_COL]
    best = torch.min(costs)
    if not bool(torch.isfinite(best)):
        return allrec, -1, torch.empty(0, dtype=torch.int32, device=records.device)
    gidx = int(torch.nonzero(costs == best)[0, 0])
    owner, lidx = divmod(gidx, P)
    n = int(allrec.view(torch.int32)[gidx, _BEST_LEN_OFF])
    path = torch.empty(best_paths.shape[1], dtype=torch.int32, device=records.device)
    if dist.get_rank(group) == owner:
        path.copy_(best_paths[lidx])
    dist.broadcast(path, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    return allrec, gidx, path[:n]


def gather_results(batch, group=None):
    """gather_records for a PlannerBatch."""
    return gather_records(batch.planners_t, batch.best_path_t, group)
