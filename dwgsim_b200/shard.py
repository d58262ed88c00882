"""Pair-index sharding across ranks: one process per GPU, torch.distributed for the plumbing.

Pairs are independent given (genome, options, seed), so rank r owns the batches b with b % world == r.  The only
steady-state exchange is the number of random pairs each rank's batch held, because `rand_ii` in the names of random
pairs is a running count over all earlier pairs (reference src/dwgsim.c:1096)."""


def plan(total_pairs, batch, rank, world):
    """[(round, first_pair, n_pairs)] of the batches `rank` owns"""
    n_batches = (total_pairs + batch - 1) // batch
    out = []
    for b in range(rank, n_batches, world):
        first = b * batch
        out.append((b // world, first, min(batch, total_pairs - first)))
    return out


def n_rounds(total_pairs, batch, world):
    n_batches = (total_pairs + batch - 1) // batch
    return (n_batches + world - 1) // world


def prefix_of(counts, rank):
    """(random pairs in the lower ranks' batches of this round, total of the round)"""
    return sum(counts[:rank]), sum(counts)


def make_exchange(device=None, group=None):
    """exchange(round, my_random) -> (before_me, round_total) as an all-gather over torch.distributed
    (NCCL over NVLink on the GPU box, gloo in the CPU tests)"""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")

    # one collective and ONE device->host read per round (a list all_gather plus an .item() per rank costs a
    # synchronising copy per rank: 0.4 ms per round at 8 GPUs)
    mine = torch.zeros(1, dtype=torch.int64, device=device)
    allc = torch.zeros(world, dtype=torch.int64, device=device)
    state = {"flat": hasattr(dist, "all_gather_into_tensor")}

    def exchange(rnd, my_random):
        mine.fill_(int(my_random))
        if state["flat"]:
            try:
                dist.all_gather_into_tensor(allc, mine, group=group)
                return prefix_of(allc.tolist(), rank)
            except (RuntimeError, NotImplementedError):
                state["flat"] = False
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        return prefix_of([int(x.item()) for x in parts], rank)

    return exchange


def interleave(per_rank_batches):
    """per_rank_batches[r] = list of byte strings (one per batch rank r owned, in order) -> the unsharded stream"""
    world = len(per_rank_batches)
    out, i = [], 0
    while True:
        r, k = i % world, i // world
        if k >= len(per_rank_batches[r]):
            break
        out.append(per_rank_batches[r][k])
        i += 1
    return b"".join(out)
