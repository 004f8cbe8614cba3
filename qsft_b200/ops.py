"""Thin Python wrappers over the C ABI (include/qsft_b200.h).  PyTorch tensors are used ONLY as device buffers
(data_ptr) and for the current CUDA stream; every computation happens in libqsft_b200.so.  No fallbacks."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .utils import index_limbs, padded_ld


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelTimers:
    """Optional CUDA-event timing of the library calls (bench.py roofline): events are recorded on the launching
    stream around each call; totals() synchronises and sums per kernel name."""

    def __init__(self):
        self.enabled = False
        self.records = []

    def reset(self, enabled=True):
        self.enabled = enabled
        self.records = []

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, work in self.records:
            ms, n, w = out.get(name, (0.0, 0, 0))
            out[name] = (ms + e0.elapsed_time(e1), n + 1, w + work)
        return out


TIMERS = KernelTimers()


class _timed:
    def __init__(self, name, work=0):
        self.name, self.work = name, work

    def __enter__(self):
        if TIMERS.enabled:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if TIMERS.enabled:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMERS.records.append((self.name, self.e0, e1, self.work))
        return False


def _pinned_block(shape, dtype):
    """Page-locked block for ONE copy to or from the device, from torch's caching host allocator: it is not handed out again
    before the asynchronous copy that uses it has run (the allocator records the stream), so a host that runs ahead of the GPU
    -- QSFT.transform(output="device_async") -- cannot overwrite data still waiting for its DMA, and results can be handed to
    the caller as views of the block.  After the first use of a size this costs microseconds (cudaHostAlloc itself
    synchronises the device and costs milliseconds)."""
    return torch.empty(tuple(int(d) for d in shape), dtype=dtype, pin_memory=True)


def _upload(arr, device):
    """Small host array -> device without stalling the host: a pageable `.to(device)` is a synchronous copy ordered behind
    everything already queued on the stream (the host then sleeps through the previous block's GEMM and launches the next
    kernels late); a pinned staging block + non_blocking copy returns at once (torch's pinned allocator keeps the block
    until the copy has run)."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.pin_memory().to(device, non_blocking=True)


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors")


def _pack_threads():
    # host threads of the packer: the ranks of one box share its cores
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(4, (os.cpu_count() or 2) // ranks))


def pad_digits(rows, ld, device, transposed=False):
    """Integer digit rows -> zero padded int8 device tensor (N, ld).  `rows` is (N, n), or (n, N) with transposed=True
    (the reference keeps the support as columns: locq (n, S)).  The narrowing cast, the transposition and the padding are
    done by the library's host packer (qsft_host_pack_digits: strided reads, its own threads -- under torchrun
    OMP_NUM_THREADS=1 makes the same cast in torch / NumPy take 3-4 ms at S = 1e5, n = 40, serial host time in front of every
    synchronous transform) straight into a pinned block, which ONE asynchronous copy uploads in the device layout."""
    a = np.asarray(rows)
    if a.ndim != 2:
        raise ValueError("digit rows must be a 2-D array")
    if a.dtype.kind not in "iu" or a.dtype.itemsize not in (1, 2, 4, 8) or a.dtype.byteorder not in "=|<":
        a = np.ascontiguousarray(a, dtype=np.int64)
    if transposed:
        a = a.T
    N, n = a.shape
    if ld < n:
        raise ValueError("ld must be at least the number of digits")
    stage = _pinned_block((N, ld), torch.int8)
    if N:
        isz = a.dtype.itemsize
        _lib.check(_lib.lib().qsft_host_pack_digits(C.c_void_p(a.ctypes.data), isz, N, n, a.strides[0] // isz, a.strides[1] // isz,
                                                    C.c_void_p(stage.data_ptr()), ld, _pack_threads()))
    return stage.to(device, non_blocking=True)


# supports at least this long are staged by all ranks together (pad_digits_sharded)
SHARD_PACK_MIN_ROWS = 32768


def shard_rows(N, world, rank):
    """Equal row blocks for an all-gather: (rows per rank, first row, one past the last row) of `rank`; the last blocks may be
    short or empty (their padding rows are zero)."""
    per = -(-int(N) // int(world)) if N > 0 else 0
    return per, min(N, rank * per), min(N, (rank + 1) * per)


def pad_digits_sharded(rows, ld, device, dist, transposed=False):
    """pad_digits for a table every rank of `dist` holds (the support): rank r stages and uploads only rows
    [r N / world, (r + 1) N / world) and one NCCL all-gather over NVLink completes the table on every GPU -- the host-side
    narrowing, serial time in front of every synchronous transform, shrinks by the number of ranks (config 5 at 8 ranks with two
    host threads each: 1.9 -> 0.25 ms).  Like the delay-row sharding itself this relies on every rank holding the same table."""
    import torch.distributed as td
    a = np.asarray(rows)
    if transposed:
        a = a.T
    N = a.shape[0]
    world, rank = dist.world_size, dist.rank
    per, lo, hi = shard_rows(N, world, rank)
    full = torch.empty((per * world, ld), dtype=torch.int8, device=device)
    mine = torch.zeros((per, ld), dtype=torch.int8, device=device) if hi - lo < per else torch.empty((per, ld), dtype=torch.int8, device=device)
    if hi > lo:
        mine[:hi - lo].copy_(pad_digits(a[lo:hi], ld, device), non_blocking=True)
    td.all_gather_into_tensor(full.view(-1), mine.view(-1), group=dist.group)
    return full[:N]


def query_lattice(M, D, q, *, device, want_idx=True, want_digits=False, limbs=None, ld=None):
    """K1.  M (n, b), D (P, n) integer arrays.  Returns (idx, dig): idx int64 tensor (P, B) or (P, B, 2) holding the
    uint64 limbs (hi, lo) bit patterns, dig int8 (P, B, ld)."""
    M = np.ascontiguousarray(M, dtype=np.int8)
    D = np.ascontiguousarray(D, dtype=np.int8)
    n, b = M.shape
    P = D.shape[0]
    if D.shape[1] != n:
        raise ValueError("D must have n columns")
    B = q ** b
    limbs = limbs or index_limbs(q, n)
    ld = ld or padded_ld(n)
    Md, Dd = torch.from_numpy(M).to(device), torch.from_numpy(D).to(device)
    idx = dig = None
    if want_idx:
        idx = torch.empty((P, B) if limbs == 1 else (P, B, 2), dtype=torch.int64, device=device)
    if want_digits:
        dig = torch.empty((P, B, ld), dtype=torch.int8, device=device)
    with torch.cuda.device(device), _timed("k1_lattice", P * B):
        _lib.check(_lib.lib().qsft_query_lattice(_ptr(Md), _ptr(Dd), q, n, b, P, _ptr(idx), limbs, _ptr(dig), ld, _stream()))
    return idx, dig


def dec_to_qary(idx, q, n, ld=None):
    """idx: int64 tensor (N,) or (N, 2) of uint64 limb bit patterns -> int8 digit rows (N, ld)."""
    _need_cuda(idx)
    limbs = 2 if idx.dim() == 2 else 1
    N = idx.shape[0]
    ld = ld or padded_ld(n)
    dig = torch.empty((N, ld), dtype=torch.int8, device=idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(_lib.lib().qsft_dec_to_qary(_ptr(idx), limbs, N, q, n, _ptr(dig), ld, _stream()))
    return dig


def qary_to_dec(dig, q, n, limbs=None):
    _need_cuda(dig)
    N, ld = dig.shape
    limbs = limbs or index_limbs(q, n)
    idx = torch.empty((N,) if limbs == 1 else (N, 2), dtype=torch.int64, device=dig.device)
    with torch.cuda.device(dig.device):
        _lib.check(_lib.lib().qsft_qary_to_dec(_ptr(dig), ld, N, q, n, _ptr(idx), limbs, _stream()))
    return idx


def eval_synth(qdig, loc, strengths, q, n, out=None, impl=0):
    """K2.  qdig (N, ld) int8, loc (S, ld) int8, strengths (S,) complex64 -> (N,) complex64."""
    _need_cuda(qdig, loc, strengths)
    N, ld = qdig.shape
    S = loc.shape[0]
    if S and loc.shape[1] != ld:
        raise ValueError("query and support digit rows must share the row stride")
    if strengths.dtype != torch.complex64:
        raise ValueError("strengths must be complex64")
    if out is None:
        out = torch.empty((N,), dtype=torch.complex64, device=qdig.device)
    with torch.cuda.device(qdig.device), _timed("k2_eval", N * S):
        _lib.check(_lib.lib().qsft_eval_synth(_ptr(qdig), N, _ptr(loc), _ptr(strengths), S, q, n, ld, _ptr(out), impl, _stream()))
    return out


def lattice_supported(q, n, b, P, S):
    return bool(_lib.lib().qsft_eval_lattice_supported(q, n, b, P, S))


def eval_synth_lattice(M, D, loc, strengths, q, out=None, residual_passes=-1):
    """Fused K1+K2 for the lattice {M l + d_p} (q = 4): M (n, b), D (P, n) integer arrays, loc (S, ld) int8 device,
    strengths (S,) complex64 device -> samples (P, q^b) complex64.  residual_passes: 0 = one GEMM pass (absolute error
    ~7e-7 max|a| per coefficient), 1 = second pass over the quantisation residual (for strengths of very different sizes),
    -1 = decided from the device data (one stream synchronisation)."""
    _need_cuda(loc, strengths)
    M = np.ascontiguousarray(M, dtype=np.int8)
    D = np.ascontiguousarray(D, dtype=np.int8)
    n, b = M.shape
    P = D.shape[0]
    S, ld = loc.shape
    dev = loc.device
    Md, Dd = _upload(M, dev), _upload(D, dev)
    if out is None:
        out = torch.empty((P, q ** b), dtype=torch.complex64, device=dev)
    with torch.cuda.device(dev), _timed("k2_eval_lattice", P * (q ** b) * S):
        _lib.check(_lib.lib().qsft_eval_synth_lattice_ex(_ptr(Md), _ptr(Dd), _ptr(loc), _ptr(strengths), S, q, n, b, P, ld,
                                                         _ptr(out), int(residual_passes), _stream()))
    return out


def gwht_batch_(x, q, b):
    """K3, in place.  x (..., q^b) complex64 contiguous."""
    _need_cuda(x)
    if x.dtype != torch.complex64 or x.shape[-1] != q ** b:
        raise ValueError("x must be complex64 with last dimension q^b")
    batch = x.numel() // (q ** b)
    with torch.cuda.device(x.device), _timed("k3_gwht", x.numel()):
        _lib.check(_lib.lib().qsft_gwht_batch(_ptr(x), batch, q, b, _stream()))
    return x


def add_noise_(x, sd, seed, offset=0):
    """x (complex64 CUDA tensor, contiguous) += sd * (N(0,1) + i N(0,1)) per element, Philox keyed by (seed, offset +
    element / 2): the same (seed, offset) adds the same noise on every rank."""
    _need_cuda(x)
    if x.dtype != torch.complex64:
        raise ValueError("x must be complex64")
    with torch.cuda.device(x.device), _timed("k5_noise", x.numel()):
        _lib.check(_lib.lib().qsft_add_noise(_ptr(x), x.numel(), float(sd), int(seed) & (2 ** 64 - 1), int(offset), _stream()))
    return x


def gwht_batch_bcast_(x, q, b, peer_ptrs):
    """K3 fused with the all-gather of its output: in place on x (rows, q^b) and, in the same kernel, stored to the same
    rows of the peers' symmetric buffers (`peer_ptrs`: device addresses of those rows, one per peer)."""
    _need_cuda(x)
    if x.dtype != torch.complex64 or x.shape[-1] != q ** b:
        raise ValueError("x must be complex64 with last dimension q^b")
    batch = x.numel() // (q ** b)
    arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[C.c_void_p(int(p)) for p in peer_ptrs])
    with torch.cuda.device(x.device), _timed("k3_gwht", x.numel()):
        _lib.check(_lib.lib().qsft_gwht_batch_bcast(_ptr(x), batch, q, b, arr, len(peer_ptrs), _stream()))
    return x


def gwht_batch_mcast_(x, q, b, mc_ptr):
    """K3 fused with the all-gather of its output through the NVLS multicast mapping of the symmetric U buffers: in place
    on x (rows, q^b); the last pass stores every element once to `mc_ptr` (the multicast address of x), NVSwitch delivers
    it to every rank including this one."""
    _need_cuda(x)
    if x.dtype != torch.complex64 or x.shape[-1] != q ** b:
        raise ValueError("x must be complex64 with last dimension q^b")
    batch = x.numel() // (q ** b)
    with torch.cuda.device(x.device), _timed("k3_gwht", x.numel()):
        _lib.check(_lib.lib().qsft_gwht_batch_mcast(_ptr(x), batch, q, b, C.c_void_p(int(mc_ptr)), _stream()))
    return x


def bins_per_rank(B, world):
    """Bins of a group every rank classifies in the bin-sharded peel: whole 128-bin tiles (the library's rule)."""
    return ((B + world - 1) // world + 127) // 128 * 128


def gwht_batch_scatter_(x, q, b, rank_ptrs, rank):
    """K3 whose last pass stores every element to the ONE rank that owns its bin in the bin-sharded peel: in place on x
    (rows, q^b); rank_ptrs[r] = device address of these rows in rank r's symmetric buffer (rank_ptrs[rank] = x)."""
    _need_cuda(x)
    if x.dtype != torch.complex64 or x.shape[-1] != q ** b:
        raise ValueError("x must be complex64 with last dimension q^b")
    batch = x.numel() // (q ** b)
    world = len(rank_ptrs)
    arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in rank_ptrs])
    with torch.cuda.device(x.device), _timed("k3_gwht", x.numel()):
        _lib.check(_lib.lib().qsft_gwht_batch_scatter(_ptr(x), batch, q, b, arr, world, int(rank), bins_per_rank(q ** b, world),
                                                      _stream()))
    return x


def channel_code(channel, nso_subtype="nso1"):
    """reconstruct_method_channel (+ nso_subtype) -> the C ABI's channel code (qsft_peel_desc.channel)."""
    if channel == "identity":
        return 0
    if channel == "nso":
        if nso_subtype not in ("nso1", "nso2"):
            raise ValueError(f"nso_subtype={nso_subtype!r} (nso1 | nso2)")
        return 1 if nso_subtype == "nso1" else 2
    if channel == "mle":
        raise NotImplementedError("reconstruct_method_channel='mle' cannot run inside QSFT.transform (the reference never "
                                  "passes selection / S_slice, qsft.py:165-173); call reconstruct.singleton_detection_mle")
    raise NotImplementedError(f"reconstruct_method_channel={channel!r} is not supported (identity | nso)")


def singleton_detect(cols, q, P_src, channel, source="identity", n=None, rs=None, nso_subtype="nso1"):
    """Batch form of reconstruct.singleton_detection: cols (N, P) complex64 CUDA tensor, one column of U per row ->
    int8 (N, n_out) digits (n_out = P_src - 1 symbols for source "identity", n decoded digits for "coded")."""
    _need_cuda(cols)
    if cols.dtype != torch.complex64 or cols.dim() != 2:
        raise ValueError("cols must be a (N, P) complex64 tensor")
    N, P = cols.shape
    chan = channel_code(channel, nso_subtype)
    dev = cols.device
    rs_exp = rs_log = None
    rs_t = rs_s = 0
    if source == "coded":
        if rs is None or n is None:
            raise ValueError("coded source decoding needs the ReedSolomon object and n")
        e, l = rs.device_tables()
        rs_exp, rs_log = torch.from_numpy(e).to(dev), torch.from_numpy(l).to(dev)
        rs_t, rs_s = rs.t, rs.s
        n_out = int(n)
    elif source == "identity":
        n_out = P_src - 1
    else:
        raise NotImplementedError(f"reconstruct_method_source={source!r} is not supported (identity | coded)")
    out = torch.empty((N, n_out), dtype=torch.int8, device=dev)
    with torch.cuda.device(dev), _timed("k4_detect", N * P):
        _lib.check(_lib.lib().qsft_singleton_detect(_ptr(cols), N, q, n_out, P, P_src, chan, 1 if source == "coded" else 0,
                                                    rs_t, rs_s, _ptr(rs_exp), _ptr(rs_log), _ptr(out), n_out, _stream()))
    return out


def detect_mle(cols, S_slice):
    """reconstruct.singleton_detection_mle for N columns sharing one candidate set: cols (N, P), S_slice (P, K) complex64
    CUDA tensors -> (k_sel int32 (N,), residual float32 (N,))."""
    _need_cuda(cols, S_slice)
    if cols.dtype != torch.complex64 or S_slice.dtype != torch.complex64 or cols.dim() != 2 or S_slice.dim() != 2:
        raise ValueError("cols (N, P) and S_slice (P, K) must be complex64 tensors")
    N, P = cols.shape
    if S_slice.shape[0] != P or S_slice.shape[1] < 1:
        raise ValueError("S_slice must have one row per delay and at least one candidate")
    k_sel = torch.empty(N, dtype=torch.int32, device=cols.device)
    res = torch.empty(N, dtype=torch.float32, device=cols.device)
    with torch.cuda.device(cols.device), _timed("k4_mle", N * P * S_slice.shape[1]):
        _lib.check(_lib.lib().qsft_detect_mle(_ptr(cols), N, P, _ptr(S_slice), S_slice.shape[1], _ptr(k_sel), _ptr(res),
                                              _stream()))
    return k_sel, res


class PeelProblem:
    """Device-side description of one peeling problem (qsft_peel_desc) + its workspaces."""

    def __init__(self, q, n, b, Ms, Ds, P_src, channel, source, cutoff, device, rs=None, nso_subtype="nso1"):
        """Ms: list of C (n, b) arrays; Ds: array (C, P, n).  channel "nso" runs nso1 (what the reference's
        QSFT.transform hard-codes, qsft.py:171) unless nso_subtype="nso2" asks for the hard-decision detector."""
        self.q, self.n, self.b = q, n, b
        self.C = len(Ms)
        Ds = np.asarray(Ds)
        self.P = Ds.shape[1]
        self.P_src = P_src
        self.B = q ** b
        self.ld = padded_ld(n)
        self.device = device
        MT = np.zeros((self.C, b, self.ld), dtype=np.int8)
        for c, M in enumerate(Ms):
            MT[c, :, :n] = np.asarray(M).T
        Dp = np.zeros((self.C, self.P, self.ld), dtype=np.int8)
        Dp[:, :, :n] = Ds
        self.MT = _upload(MT, device)
        self.D = _upload(Dp, device)
        self.rs_exp = self.rs_log = None
        rs_t = rs_s = 0
        if source == "coded":
            if rs is None:
                raise ValueError("coded source decoding needs the ReedSolomon object (source_decoder)")
            e, l = rs.device_tables()
            self.rs_exp, self.rs_log = _upload(e, device), _upload(l, device)
            rs_t, rs_s = rs.t, rs.s
        chan = channel_code(channel, nso_subtype)
        src = {"identity": 0, "coded": 1}.get(source)
        if src is None:
            raise NotImplementedError(f"reconstruct_method_source={source!r} is not supported (identity | coded)")
        self.desc = _lib.PeelDesc(q=q, n=n, b=b, C=self.C, P=self.P, P_src=P_src, channel=chan, source=src,
                                  rs_t=rs_t, rs_s=rs_s, ld=self.ld, cutoff=float(cutoff),
                                  MT=self.MT.data_ptr(), D=self.D.data_ptr(),
                                  rs_exp=0 if self.rs_exp is None else self.rs_exp.data_ptr(),
                                  rs_log=0 if self.rs_log is None else self.rs_log.data_ptr())

    _WORKSPACES = {}

    def alloc(self, max_finds, max_uniq=None, reuse=False):
        """Find / distinct-k buffers.  reuse=True takes them from a per-shape cache (one set per device and shape, handed
        to one problem at a time): QSFT.transform peels one signal after the other and ~10 allocations of up to hundreds
        of MB per transform are pure overhead."""
        key = (str(self.device), self.C, self.B, self.ld, int(max_finds), int(max_uniq or max_finds))
        if reuse and key in PeelProblem._WORKSPACES:
            self.__dict__.update(PeelProblem._WORKSPACES[key])
            self.counters.zero_()
            return
        self._alloc(max_finds, max_uniq)
        if reuse:
            names = ("max_finds", "find_cj", "find_k", "find_rho", "find_round", "find_id", "counters", "max_uniq", "seen0",
                     "uniq_k", "uniq_sum", "uniq_cnt", "uniq_key", "uniq_next", "uniq")
            PeelProblem._WORKSPACES.clear()            # keep one shape: the buffers are large
            PeelProblem._WORKSPACES[key] = {nm: getattr(self, nm) for nm in names}

    def _alloc(self, max_finds, max_uniq=None):
        dev = self.device
        self.max_finds = int(max_finds)
        self.find_cj = torch.empty(self.max_finds, dtype=torch.int64, device=dev)
        self.find_k = torch.empty((self.max_finds, self.ld), dtype=torch.int8, device=dev)
        self.find_rho = torch.empty(self.max_finds, dtype=torch.complex64, device=dev)
        self.find_round = torch.empty(self.max_finds, dtype=torch.int32, device=dev)
        self.find_id = torch.empty((self.C, self.B), dtype=torch.int32, device=dev)
        self.counters = torch.zeros(8, dtype=torch.int64, device=dev)
        # distinct-k list (device-side averaging of duplicate finds, qsft.py:247-255)
        self.max_uniq = int(max_uniq) if max_uniq else self.max_finds
        self.seen0 = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.uniq_k = torch.empty((self.max_uniq, self.ld), dtype=torch.int8, device=dev)
        self.uniq_sum = torch.empty(self.max_uniq, dtype=torch.complex64, device=dev)
        self.uniq_cnt = torch.empty(self.max_uniq, dtype=torch.int32, device=dev)
        self.uniq_key = torch.empty(self.max_uniq, dtype=torch.int64, device=dev)
        self.uniq_next = torch.empty(self.max_uniq, dtype=torch.int32, device=dev)
        self.uniq = _lib.Uniq(seen0=self.seen0.data_ptr(), uniq_k=self.uniq_k.data_ptr(), uniq_sum=self.uniq_sum.data_ptr(),
                              uniq_cnt=self.uniq_cnt.data_ptr(), uniq_key=self.uniq_key.data_ptr(),
                              uniq_next=self.uniq_next.data_ptr(), max_uniq=self.max_uniq)

    # -- whole loop on one GPU -------------------------------------------------------------------------
    def peel(self, U):
        """Runs the full round loop in the library.  U (C, P, B) complex64 is modified in place.
        Returns (n_finds, n_rounds); finds are in self.find_* [0:n_finds], the distinct k in self.uniq_* [0:self.n_uniq]."""
        _need_cuda(U)
        assert U.shape == (self.C, self.P, self.B) and U.dtype == torch.complex64
        nf, nu, nr = C.c_int64(0), C.c_int64(0), C.c_int(0)
        with torch.cuda.device(self.device), _timed("k4_peel", U.numel()):
            _lib.check(_lib.lib().qsft_peel(C.byref(self.desc), _ptr(U), _ptr(self.find_cj), _ptr(self.find_k),
                                            _ptr(self.find_rho), _ptr(self.find_round), _ptr(self.find_id),
                                            self.max_finds, _ptr(self.counters), C.byref(self.uniq), C.byref(nf),
                                            C.byref(nu), C.byref(nr), _stream()))
        self.n_uniq = nu.value
        return nf.value, nr.value

    def finds(self, n_finds):
        """Host copy of the find list after peel(): (cj, k (F, n), rho, round) of the F valid finds among the first
        `n_finds` slots (the on-device loop hands slots out in chunks per warp; unused ones carry find_cj = -1)."""
        cj = self.find_cj[:n_finds].cpu().numpy()
        ok = cj >= 0
        return (cj[ok], self.find_k[:n_finds, :self.n].cpu().numpy()[ok], self.find_rho[:n_finds].cpu().numpy()[ok],
                self.find_round[:n_finds].cpu().numpy()[ok])

    def peel_blocks(self, blocks, wait=True):
        """The same loop on bins given as C * R separate (P_src, B) complex64 row blocks (block c * R + r; what get_MDU
        returns), read in place -- no (C, P, B) copy.  Returns (n_finds, n_rounds), or None when the shape does not fit the
        on-device loop (the caller then assembles U and uses peel()).  wait=False: the loop is only queued (returns
        (-1, -1), n_uniq stays unknown); the outcome is in self.counters -- see PeelOutcome."""
        R = self.P // self.P_src
        if len(blocks) != self.C * R:
            raise ValueError("expected C * R blocks")
        ldU = None
        for t in blocks:
            _need_cuda(t)
            if t.dtype != torch.complex64 or t.dim() != 2 or t.shape != (self.P_src, self.B):
                raise ValueError("every block must be a (P_src, q^b) complex64 tensor")
            ldU = t.stride(0) if ldU is None else ldU
            if t.stride(0) != ldU:
                return None
        arr = (C.c_void_p * len(blocks))(*[C.c_void_p(t.data_ptr()) for t in blocks])
        nf, nu, nr = C.c_int64(0), C.c_int64(0), C.c_int(0)
        outs = (C.byref(nf), C.byref(nu), C.byref(nr)) if wait else (None, None, None)
        with torch.cuda.device(self.device), _timed("k4_peel", self.C * self.P * self.B):
            rc = _lib.lib().qsft_peel_blocks(C.byref(self.desc), arr, ldU, _ptr(self.find_cj), _ptr(self.find_k),
                                             _ptr(self.find_rho), _ptr(self.find_round), _ptr(self.find_id), self.max_finds,
                                             _ptr(self.counters), C.byref(self.uniq), outs[0], outs[1], outs[2], _stream())
        if rc == -3:                                   # QSFT_EUNSUPPORTED
            return None
        _lib.check(rc)
        if not wait:
            self.n_uniq = -1
            return -1, -1
        self.n_uniq = nu.value
        return nf.value, nr.value

    # -- the on-device loop sharded over the ranks of a DistContext -----------------------------------------------------
    _SHARD_WS = {}

    def peel_blocks_sharded(self, blocks, dist, wait=True):
        """Bin-sharded on-device loop (qsft_peel_blocks_sharded): every rank classifies its bins, the round's finds travel
        between the ranks inside the kernel (NVLink stores into the peers' symmetric workspaces).  Needs alloc() for the
        distinct-k buffers.  Returns n_rounds, or None when the shape / platform does not fit (caller falls back)."""
        R = self.P // self.P_src
        if len(blocks) != self.C * R or dist is None or dist.world_size < 2 or dist.world_size > 8:
            return None
        ldU = blocks[0].stride(0)
        for t in blocks:
            _need_cuda(t)
            if t.dtype != torch.complex64 or t.shape != (self.P_src, self.B) or t.stride(0) != ldU:
                return None
        max_finds = self.max_finds
        nbytes = int(_lib.lib().qsft_peel_sharded_workspace_bytes(C.byref(self.desc), max_finds))
        if nbytes <= 0:
            return None
        key = (str(self.device), nbytes, id(dist))
        ws = PeelProblem._SHARD_WS.get(key)
        if ws is None:
            item = dist.symm_acquire((nbytes + 3) // 4, self.device)
            if item is None:
                return None
            item[0][:2048].zero_()                          # control block (8 KB)
            torch.cuda.current_stream(self.device).synchronize()
            dist.barrier()
            ws = {"item": item, "epoch": 0}
            PeelProblem._SHARD_WS.clear()                   # one shape at a time: the workspace is large
            PeelProblem._SHARD_WS[key] = ws
        buf, hdl, ptrs, _mc = ws["item"]
        ws["epoch"] += 1
        hdl.barrier()                                       # every rank is done with the previous peel on this workspace
        peers = (C.c_void_p * dist.world_size)(*[C.c_void_p(int(p)) for p in ptrs])
        shard = _lib.Shard(rank=dist.rank, world=dist.world_size, peers=peers, epoch=ws["epoch"])
        arr = (C.c_void_p * len(blocks))(*[C.c_void_p(t.data_ptr()) for t in blocks])
        nu, nr = C.c_int64(0), C.c_int(0)
        outs = (C.byref(nu), C.byref(nr)) if wait else (None, None)
        with torch.cuda.device(self.device), _timed("k4_peel", self.C * self.P * self.B):
            rc = _lib.lib().qsft_peel_blocks_sharded(C.byref(self.desc), arr, ldU, C.byref(shard), max_finds, _ptr(self.counters),
                                                     C.byref(self.uniq), outs[0], outs[1], _stream())
        if rc == -3:
            return None
        _lib.check(rc)
        if not wait:                                        # queued only: the outcome is in self.counters (PeelOutcome)
            self.n_uniq = -1
            return -1
        self.n_uniq = nu.value
        return nr.value

    def distinct(self, n_uniq=None):
        """Host copy of the distinct-k list in the reference's first-seen order:
        (k (K, n) int8, mean rho (K,) complex128, count (K,) int32).  The entries are ordered, gathered and averaged on the
        device (qsft_peel_distinct) and come back through pinned blocks with one synchronisation."""
        nu = self.n_uniq if n_uniq is None else n_uniq
        if nu == 0:
            return np.zeros((0, self.n), dtype=np.int8), np.zeros(0, dtype=np.complex128), np.zeros(0, dtype=np.int32)
        order = torch.argsort(self.uniq_key[:nu])                       # keys are unique: (round << 48) | (c B + j)
        dev = self.device
        k_d = torch.empty((nu, self.n), dtype=torch.int8, device=dev)
        m_d = torch.empty((nu, 2), dtype=torch.float64, device=dev)
        c_d = torch.empty(nu, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().qsft_peel_distinct(C.byref(self.uniq), _ptr(order), nu, self.n, self.ld, _ptr(k_d), _ptr(m_d),
                                                     _ptr(c_d), _stream()))
        # page-locked blocks of their own (caching host allocator): the arrays handed out are views of them, no host copy
        k_h = _pinned_block(k_d.shape, k_d.dtype)
        m_h = _pinned_block(m_d.shape, m_d.dtype)
        c_h = _pinned_block(c_d.shape, c_d.dtype)
        k_h.copy_(k_d, non_blocking=True)
        m_h.copy_(m_d, non_blocking=True)
        c_h.copy_(c_d, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return k_h.numpy(), m_h.numpy().view(np.complex128).reshape(nu), c_h.numpy()

    def reduce(self, find_cj, find_k, find_rho, find_id, f_begin, n_finds, round_no):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().qsft_peel_reduce(C.byref(self.desc), _ptr(find_cj), _ptr(find_k), _ptr(find_rho),
                                                   _ptr(find_id), f_begin, n_finds, round_no, C.byref(self.uniq),
                                                   _ptr(self.counters), _stream()))

    # -- single steps (bin-sharded multi-GPU loop drives these) ----------------------------------------
    def classify(self, U, j_begin, j_end, round_no):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().qsft_peel_classify(C.byref(self.desc), _ptr(U), j_begin, j_end, _ptr(self.find_cj),
                                                     _ptr(self.find_k), _ptr(self.find_rho), _ptr(self.find_round),
                                                     _ptr(self.find_id), self.max_finds, round_no, _ptr(self.counters),
                                                     _stream()))

    def apply(self, U, j_begin, j_end, find_cj, find_k, find_rho, find_id, f_begin, n_finds, dedupe=True,
              owner_count=None):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().qsft_peel_apply(C.byref(self.desc), _ptr(U), j_begin, j_end, _ptr(find_cj), _ptr(find_k),
                                                  _ptr(find_rho), _ptr(find_id), f_begin, n_finds, 1 if dedupe else 0,
                                                  _ptr(owner_count), _stream()))


def closed_form_bins(M, D, q, loc, strengths, out=None):
    """Verification helper: bins of one (M, D block) from the support directly.  M (n, b), D (P, n) arrays;
    loc (S, ld) int8 device, strengths (S,) complex64 device.  Returns (P, B) complex64."""
    M = np.asarray(M)
    D = np.asarray(D)
    n, b = M.shape
    P = D.shape[0]
    ld = loc.shape[1]
    dev = loc.device
    MT = np.zeros((b, ld), dtype=np.int8)
    MT[:, :n] = M.T
    Dp = np.zeros((P, ld), dtype=np.int8)
    Dp[:, :n] = D
    MTd, Dd = torch.from_numpy(MT).to(dev), torch.from_numpy(Dp).to(dev)
    if out is None:
        out = torch.zeros((P, q ** b), dtype=torch.complex64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qsft_closed_form_bins(_ptr(MTd), _ptr(Dd), q, n, b, P, _ptr(loc), ld, _ptr(strengths),
                                                    loc.shape[0], _ptr(out), _stream()))
    return out
