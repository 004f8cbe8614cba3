"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the GPU box, gloo in the CPU tests).

Sharding: the C*R*P_src delay rows are independent for sampling + FFT -> block partitioned over ranks;
U is all-gathered once; peeling is bin-sharded (each rank classifies and updates bins j in its range of every
group) with ONE all-gather of the round's finds per round -- the only collective type on the path."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as td


class DistContext:
    def __init__(self, group=None, symmetric=True, peel_mode="auto", replicate_peel_max_bytes=8 << 30):
        """peel_mode: "sharded" = bin-sharded on-device loop: every rank classifies its bins, the round's finds are exchanged
        INSIDE the persistent kernel over NVLink (symmetric workspaces; qsft_peel_blocks_sharded); "sharded_host" = the same
        sharding driven from the host with one NCCL all-gather of the finds per round (peel_sharded below; also the fallback
        where symmetric memory is not available); "replicated" = every rank peels its full copy of U with the single-GPU
        loop (no exchange at all); "auto" = sharded when symmetric memory is available, else replicated while U is at most
        `replicate_peel_max_bytes`, else sharded_host."""
        if not td.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if peel_mode not in ("auto", "sharded", "sharded_host", "replicated"):
            raise ValueError("peel_mode must be 'auto', 'sharded', 'sharded_host' or 'replicated'")
        self.peel_mode = peel_mode
        self.replicate_peel_max_bytes = int(replicate_peel_max_bytes)
        self.group = group
        self.rank = td.get_rank(group)
        self.world_size = td.get_world_size(group)
        # symmetric (peer-mapped) U buffers let K3 store its output straight into every peer (fused all-gather);
        # falls back to NCCL all_gather_into_tensor when symmetric memory is not available
        self.symmetric = bool(symmetric) and td.get_backend(group) == "nccl" and self.world_size <= 8
        self.multicast = os.environ.get("QSFT_NO_MULTICAST") is None       # (env: A/B of multimem.st against unicast stores)
        self._symm_free = {}
        self._pending_checks = []

    def shard_peel(self, u_bytes):
        """Placement of the peeling loop for a transform whose bins take `u_bytes`: "device" (bin-sharded on-device loop),
        "host" (bin-sharded, NCCL exchange per round) or "" (replicated)."""
        if self.world_size == 1 or self.peel_mode == "replicated":
            return ""
        if self.peel_mode == "sharded_host":
            return "host"
        if self.symmetric:
            return "device"
        if self.peel_mode == "sharded" or u_bytes > self.replicate_peel_max_bytes:
            return "host"
        return ""

    # -- symmetric buffers ---------------------------------------------------------------------------------
    def symm_acquire(self, nfloats, device):
        """A float32 symmetric buffer of `nfloats` elements + its rendezvous handle, reused across signals (the
        rendezvous costs milliseconds).  Returns None when symmetric memory cannot be set up."""
        if not self.symmetric:
            return None
        pool = self._symm_free.setdefault(int(nfloats), [])
        # a rank that reuses a pooled buffer while a peer sets up a fresh one would miss the peer's rendezvous: the choice
        # goes into the deferred agreement checks (verify() names it instead of leaving wrong bins behind)
        self.post_check("symmetric buffer reuse (pooled / fresh)", np.array([int(nfloats), 1 if pool else 0], dtype=np.int64))
        if pool:
            return pool.pop()
        buf, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else td.group.WORLD
            buf = symm_mem.empty(int(nfloats), dtype=torch.float32, device=device)
        except Exception as exc:  # pragma: no cover - depends on the platform
            err = repr(exc)
        # the fallback is decided by ALL ranks together (the rendezvous below is collective): one rank on the NCCL
        # all-gather path while its peers wait in the fused K3 store path would hang
        if not self._all_ok(buf is not None, device):
            self.symmetric = False
            self.symmetric_error = err or "symmetric allocation failed on a peer rank"
            return None
        try:
            hdl = symm_mem.rendezvous(buf, grp)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            # NVLS multicast alias of the buffer (0 where the platform has none): one multimem.st reaches every rank
            mc = 0
            if self.multicast:
                try:
                    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                except Exception:  # pragma: no cover - depends on the platform
                    mc = 0
            ok = True
        except Exception as exc:  # pragma: no cover - depends on the platform
            err, ok = repr(exc), False
        if not self._all_ok(ok, device):
            self.symmetric = False
            self.symmetric_error = err or "symmetric rendezvous failed on a peer rank"
            return None
        return (buf, hdl, ptrs, mc)

    def _all_ok(self, ok, device):
        """True iff `ok` holds on every rank (one small all-reduce; only on the set-up path of a NEW symmetric buffer)."""
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        td.all_reduce(t, op=td.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def symm_release(self, item):
        if item is not None:
            self._symm_free.setdefault(int(item[0].numel()), []).append(item)

    def all_gather_rows_(self, buf, rows_per_rank):
        """buf (world*rows_per_rank, W): every rank has filled its own row block; gather all blocks in place."""
        mine = buf[self.rank * rows_per_rank:(self.rank + 1) * rows_per_rank].clone()
        flat = torch.view_as_real(buf) if buf.is_complex() else buf
        minef = torch.view_as_real(mine) if mine.is_complex() else mine
        td.all_gather_into_tensor(flat.reshape(-1), minef.reshape(-1), group=self.group)

    def all_gather_var(self, tensors, count, cap, extra=0):
        """All-gather `count` leading rows of each tensor in `tensors` (fixed capacity `cap` per rank).
        `count` / `extra` may be Python ints or a 2-element int64 CUDA tensor [count, extra] (then no host sync is needed
        before the exchange).  Returns (list of concatenated tensors, counts per rank); the sum over ranks of `extra`
        is left in self.last_extra_sum (saves a separate all-reduce)."""
        dev = tensors[0].device
        if isinstance(count, torch.Tensor):
            cnt = count.to(torch.int64).contiguous()
        else:
            cnt = torch.tensor([int(count), int(extra)], dtype=torch.int64, device=dev)
        allc = torch.empty(2 * self.world_size, dtype=torch.int64, device=dev)
        td.all_gather_into_tensor(allc, cnt, group=self.group)
        allc = allc.cpu().view(self.world_size, 2)          # the round's only host synchronisation
        counts = allc[:, 0].tolist()
        self.last_extra_sum = int(allc[:, 1].sum().item())
        if max(counts) > cap:
            raise RuntimeError("find buffer overflow in sharded peel")
        m = max(counts)
        outs = []
        for t in tensors:
            real = torch.view_as_real(t) if t.is_complex() else t
            send = real[:m].contiguous()
            recv = torch.empty((self.world_size,) + tuple(send.shape), dtype=send.dtype, device=dev)
            td.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
            parts = [recv[r, :counts[r]] for r in range(self.world_size)]
            cat = torch.cat(parts, dim=0)
            outs.append(torch.view_as_complex(cat) if t.is_complex() else cat)
        return outs, counts

    def all_reduce_sum(self, value, device):
        """Sum of one integer over ranks (stop rule of the peel loop)."""
        t = torch.tensor([int(value)], dtype=torch.int64, device=device)
        td.all_reduce(t, group=self.group)
        return int(t.item())

    def barrier(self):
        td.barrier(group=self.group)

    # -- agreement of host-side randomness --------------------------------------------------------------------------
    # Ms / Ds, the group / repeat selection of get_MDU and the noise seed are drawn from the host NumPy RNG on every rank
    # (the reference's RNG order).  Ranks that were seeded differently would transform with different matrices and the
    # gathered bins would silently mix them, so: the matrices are compared (hash all-gather, error on mismatch) and small
    # index arrays are taken from rank 0.
    def _device(self):
        return torch.device("cuda", torch.cuda.current_device()) if td.get_backend(self.group) == "nccl" else torch.device("cpu")

    def assert_same(self, what, *arrays):
        """Raises on every rank when the byte contents of `arrays` differ between ranks."""
        import hashlib
        h = hashlib.blake2b(digest_size=8)
        for a in arrays:
            a = np.ascontiguousarray(a)
            h.update(str(a.shape).encode())
            h.update(a.tobytes())
        mine = torch.tensor([int.from_bytes(h.digest(), "little", signed=True)], dtype=torch.int64, device=self._device())
        everyone = torch.empty(self.world_size, dtype=torch.int64, device=mine.device)
        td.all_gather_into_tensor(everyone, mine, group=self.group)
        if not bool((everyone == everyone[0]).all()):
            raise RuntimeError(f"{what} differ between ranks: seed the NumPy RNG identically on every rank (or pass the same "
                               f"Ms= / Ds=), the delay rows of ONE transform are sharded over the ranks")

    def post_check(self, what, *arrays):
        """assert_same without a collective: only the hash is taken now; verify() compares the hashes of ALL pending checks
        across the ranks with one small all-gather (no NCCL kernels between the sampling, transform and peel kernels)."""
        import hashlib
        h = hashlib.blake2b(digest_size=8)
        for a in arrays:
            a = np.ascontiguousarray(a)
            h.update(str(a.shape).encode())
            h.update(a.tobytes())
        self._pending_checks.append((what, int.from_bytes(h.digest(), "little", signed=True)))

    def verify(self):
        """Compares the pending checks across the ranks (every rank must have queued the same number: they come from the
        same code path); raises on a mismatch."""
        pending, self._pending_checks = self._pending_checks, []
        if not pending:
            return
        mine = torch.tensor([len(pending)] + [v for _, v in pending], dtype=torch.int64).to(self._device())
        counts = torch.empty(self.world_size, dtype=torch.int64, device=mine.device)
        td.all_gather_into_tensor(counts, mine[:1].clone(), group=self.group)
        if not bool((counts == counts[0]).all()):
            raise RuntimeError("the ranks queued different numbers of agreement checks: they are not running the same transforms")
        everyone = torch.empty(self.world_size * mine.numel(), dtype=torch.int64, device=mine.device)
        td.all_gather_into_tensor(everyone, mine, group=self.group)
        v = everyone.cpu().view(self.world_size, -1)[:, 1:]
        for i, (what, _) in enumerate(pending):
            if not bool((v[:, i] == v[0, i]).all()):
                raise RuntimeError(f"{what} differ between ranks: seed the NumPy RNG identically on every rank (or pass the same "
                                   f"Ms= / Ds=), the delay rows of ONE transform are sharded over the ranks")

    def from_rank0(self, arr):
        """The int64 array `arr` of rank 0, on every rank (same shape everywhere)."""
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self._device())
        td.broadcast(t, src=td.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return t.cpu().numpy()


def bin_range(B, rank, world):
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def peel_sharded(prob, U, dist, max_rounds=15, to_host=False):
    """Bin-sharded peeling loop (qsft.py:151-241 semantics).  Every rank holds the full U but only classifies and
    updates bins in its own j-range; finds are exchanged with one all-gather per round.
    Returns (cj, k, rho, round, n_rounds), identical on every rank (NumPy arrays when to_host else CUDA tensors);
    the distinct-k list is left in prob.uniq_* [0:prob.n_uniq] (see PeelProblem.distinct)."""
    q, n, C, B = prob.q, prob.n, prob.C, prob.B
    dev = prob.device
    jb, je = bin_range(B, dist.rank, dist.world_size)
    cap = max(1024, 2 * C * (je - jb))
    prob.alloc(max_finds=cap, max_uniq=max(cap, min(C * B, 4 * cap * dist.world_size)))
    find_id_full = torch.empty((C, B), dtype=torch.int32, device=dev)
    prob.counters.zero_()
    prob.seen0.zero_()
    all_cj, all_k, all_rho, all_round = [], [], [], []
    peeling_max = float(q) ** n
    guard_can_bind = peeling_max <= 15.0 * C * B
    num_peeling, rnd, cont = 0, 0, True
    while cont and num_peeling < peeling_max and rnd < max_rounds:
        rnd += 1
        prob.counters[:4].zero_()
        prob.classify(U, jb, je, rnd)
        # counters[0:2] = (finds, multitons) of this rank travel with the exchange: one host sync per round
        (cj, k, rho), counts = dist.all_gather_var([prob.find_cj, prob.find_k, prob.find_rho], prob.counters[:2], cap)
        n_multi = dist.last_extra_sum
        nf = int(sum(counts))
        if n_multi == 0 or nf == 0:
            cont = False
        if nf > 0:
            all_cj.append(cj)
            all_k.append(k[:, :n])
            all_rho.append(rho)
            all_round.append(torch.full((nf,), rnd, dtype=torch.int32, device=dev))
            # rebuild the (C, B) find table of the whole round: duplicates of a k ("last (i, j) wins", averaging of
            # repeated finds) are looked up through it, locally on every rank
            find_id_full.fill_(-1)
            find_id_full.view(-1)[cj] = torch.arange(nf, dtype=torch.int32, device=dev)
            cjc, kc, rhoc = cj.contiguous(), k.contiguous(), rho.contiguous()
            prob.reduce(cjc, kc, rhoc, find_id_full, 0, nf, rnd)
            if cont or guard_can_bind:
                owners = torch.zeros(1, dtype=torch.int64, device=dev) if guard_can_bind else None
                prob.apply(U, jb, je, cjc, kc, rhoc, find_id_full, 0, nf, dedupe=True, owner_count=owners)
                if guard_can_bind:
                    num_peeling += int(owners.item())
    if all_cj:
        out = (torch.cat(all_cj), torch.cat(all_k), torch.cat(all_rho), torch.cat(all_round))
    else:
        out = (torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros((0, n), dtype=torch.int8, device=dev),
               torch.zeros(0, dtype=torch.complex64, device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
    prob.n_uniq = int(prob.counters[4].item())
    if prob.n_uniq > prob.max_uniq:
        raise RuntimeError("distinct-k buffer overflow in sharded peel")
    if to_host:
        out = tuple(t.cpu().numpy() for t in out)
    return out + (rnd,)
