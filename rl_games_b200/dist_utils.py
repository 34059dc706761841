"""Host-side multi-GPU logic (pure tensor math + one collective), kept free of CUDA-only code so it is testable on CPU
with `gloo`: the pooled running-statistics merge of a2c_common.py:43-93 (`merge_rank_stats`) for ALL normalisers in ONE
packed fp64 all-reduce, with the "snapshot of the last merged totals" bookkeeping (:46-58, :767-780).

Everything is in-place on persistent tensors (no Python-side state changes, no host syncs), so the whole merge can be
captured inside the per-epoch CUDA graph together with the NCCL all-reduce."""
import torch


def stats_totals(count, mean, var):
    """(count, sum_x, sum_x2) totals equivalent to a RunningMeanStd state (a2c_common.py:43-47)."""
    cnt = count.to(torch.float64).reshape(-1)
    return (cnt.clone(), mean * cnt, (var + mean ** 2) * cnt)


class PackedStatsSync:
    """mods: list of (name, count[int64 1], mean[f64 D], var[f64 D]) tensors, updated IN PLACE by sync().
    A zero snapshot means "the module's whole history is rank-local" (fresh start, merged whole)."""

    def __init__(self, mods):
        self.mods = mods
        self.snap = []
        n = 0
        for _, count, mean, var in mods:
            self.snap.append([torch.zeros(1, dtype=torch.float64, device=mean.device), torch.zeros_like(mean), torch.zeros_like(mean)])
            n += 1 + 2 * mean.numel()
        dev = mods[0][2].device if mods else 'cpu'
        self.flat = torch.zeros(n, dtype=torch.float64, device=dev)

    def seed(self):
        """Mark the current state as already-shared history (after loading a checkpoint on every rank)."""
        for (_, count, mean, var), s in zip(self.mods, self.snap):
            cur = stats_totals(count, mean, var)
            for d, c in zip(s, cur):
                d.copy_(c)

    def sync(self, all_reduce):
        """all_reduce(t): SUM `t` in place across ranks."""
        off = 0
        for (_, count, mean, var), s in zip(self.mods, self.snap):
            D = mean.numel()
            cnt = count.to(torch.float64).reshape(-1)
            self.flat[off:off + 1] = cnt - s[0]
            self.flat[off + 1:off + 1 + D] = mean * cnt - s[1]
            self.flat[off + 1 + D:off + 1 + 2 * D] = (var + mean ** 2) * cnt - s[2]
            off += 1 + 2 * D
        all_reduce(self.flat)
        off = 0
        for (_, count, mean, var), s in zip(self.mods, self.snap):
            D = mean.numel()
            s[0] += self.flat[off:off + 1]
            s[1] += self.flat[off + 1:off + 1 + D]
            s[2] += self.flat[off + 1 + D:off + 1 + 2 * D]
            off += 1 + 2 * D
            count.copy_(torch.round(s[0]).to(torch.int64).view_as(count))
            mean.copy_(s[1] / s[0])
            var.copy_((s[2] / s[0] - mean ** 2).clamp_(min=1e-8))


def merge_stats_packed(mods, snapshots, all_reduce):
    """Functional wrapper (kept for tests): snapshots is a dict name -> (n, sum_x, sum_x2) or missing."""
    sync = PackedStatsSync(mods)
    for (name, _, _, _), s in zip(mods, sync.snap):
        prev = snapshots.get(name)
        if prev is not None:
            for d, c in zip(s, prev):
                d.copy_(c.reshape(d.shape))
    sync.sync(all_reduce)
    out = dict(snapshots)
    for (name, _, _, _), s in zip(mods, sync.snap):
        out[name] = tuple(t.clone() for t in s)
    return out
