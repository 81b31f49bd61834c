"""Multi-GPU sharding of independent phase blocks (one process per GPU, torch.distributed).

Phase blocks share nothing (the reference runs them as independent thread-pool jobs, src/main.rs:385-408), so the
path shards with NO data-path collective: every rank solves its own blocks; the only communication is the result
hand-off (fixed-stride records gathered to every rank / rank 0, re-ordered by block index exactly like the
reference's OrderedVcfWriter does for its worker results, writers/ordered_vcf_writer.rs:158-170).
Works with the NCCL backend on GPUs and with gloo on CPU (tests).
"""
import numpy as np


def block_costs(n_var, n_cells):
    """Serial-chain cost model used for balancing: cells * min(N, 40) + N (heuristic look-ahead is 40 variants)."""
    n_var = np.asarray(n_var, np.int64)
    return np.asarray(n_cells, np.int64) * np.minimum(n_var, 40) + n_var


def lpt_partition(costs, world):
    """Longest-processing-time-first dealing: sort by cost descending, give each block to the least loaded rank.
    Returns a list of index arrays (one per rank); every block appears exactly once."""
    costs = np.asarray(costs, np.int64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world, np.int64)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += int(costs[i])
    return [np.array(sorted(p), np.int64) for p in parts]


def contiguous_shard(n_items, rank, world):
    """[lo, hi) of a balanced contiguous split (used when every rank regenerates its shard from the seed)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_block_records(local_ids, local_records, n_total, group=None):
    """All-gathers fixed-stride per-block records and re-orders them by global block index.

    local_ids: int64[k] global block indices solved by this rank; local_records: [k, stride] integer array.
    Returns an [n_total, stride] array on every rank.  Uses one all_gather of the padded id+record tensor.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local_ids = np.asarray(local_ids, np.int64)
    local_records = np.asarray(local_records)
    # a rank without blocks does not know the record width: the ranks agree on (rows, stride) first
    my_stride = int(local_records.size // len(local_ids)) if len(local_ids) else 0
    local_records = local_records.reshape(len(local_ids), my_stride).astype(np.int64)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    k = torch.tensor([len(local_ids), my_stride], dtype=torch.int64, device=dev)
    ks = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ks, k, group=group)
    sizes = torch.stack(ks).cpu().numpy()
    kmax, stride = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    if (sizes[:, 1][sizes[:, 0] > 0] != stride).any():
        raise ValueError("ranks disagree on the record width")
    pad = torch.full((kmax, stride + 1), -1, dtype=torch.int64, device=dev)
    if len(local_ids):
        pad[:len(local_ids), 0] = torch.from_numpy(local_ids).to(dev)
        pad[:len(local_ids), 1:] = torch.from_numpy(local_records).to(dev)
    allp = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(allp, pad, group=group)
    out = np.full((int(n_total), stride), -1, np.int64)
    seen = np.zeros(int(n_total), bool)
    for t in allp:
        a = t.cpu().numpy()
        a = a[a[:, 0] >= 0]
        if (seen[a[:, 0]]).any():
            raise RuntimeError("block solved by more than one rank")
        seen[a[:, 0]] = True
        out[a[:, 0]] = a[:, 1:]
    if not seen.all():
        raise RuntimeError("%d blocks were not solved by any rank" % int((~seen).sum()))
    return out
