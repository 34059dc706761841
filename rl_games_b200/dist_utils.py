"""Host-side multi-GPU logic (pure tensor math + one collective), kept free of CUDA so it is testable on CPU with
`gloo`: the pooled running-statistics merge of a2c_common.py:43-93 (`merge_rank_stats`) for ALL normalisers in ONE
packed fp64 all-reduce, and its snapshot re-seeding (:46-58, :767-780)."""
import torch


def stats_totals(count, mean, var):
    """(count, sum_x, sum_x2) totals equivalent to a RunningMeanStd state (a2c_common.py:43-47)."""
    cnt = count.to(torch.float64).reshape(-1)
    return (cnt.clone(), mean * cnt, (var + mean ** 2) * cnt)


def merge_stats_packed(mods, snapshots, all_reduce):
    """mods: list of (name, count[int64 1], mean[f64 D], var[f64 D]) tensors, updated IN PLACE.
    snapshots: dict name -> (n, sum_x, sum_x2) of the last merge (missing => whole history is rank-local).
    all_reduce(t): SUM `t` in place across ranks.  Returns the new snapshots dict."""
    packed, bases = [], []
    for name, count, mean, var in mods:
        cur = stats_totals(count, mean, var)
        prev = snapshots.get(name)
        if prev is None:
            prev = tuple(torch.zeros_like(c) for c in cur)
        bases.append(prev)
        packed += [c - p for c, p in zip(cur, prev)]
    flat = torch.cat([p.reshape(-1) for p in packed])
    all_reduce(flat)
    out, off = dict(snapshots), 0
    for (name, count, mean, var), base in zip(mods, bases):
        d = []
        for b in base:
            d.append(flat[off:off + b.numel()].view_as(b))
            off += b.numel()
        n, wm, wsq = base[0] + d[0], base[1] + d[1], base[2] + d[2]
        count.copy_(torch.round(n).to(torch.int64).view_as(count))
        mean.copy_(wm / n)
        var.copy_((wsq / n - mean ** 2).clamp_(min=1e-8))
        out[name] = (n.clone(), wm.clone(), wsq.clone())
    return out


def seed_snapshots(mods):
    return {name: stats_totals(count, mean, var) for name, count, mean, var in mods}
