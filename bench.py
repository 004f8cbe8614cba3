#!/usr/bin/env python
"""Benchmark of the q-SFT transform path (sample + FFT + peel) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--repeat R] [--sparsity S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the config the north-star target is quoted on): synthetic sparse signal
q=4, n=40, b=10, S=1e5, C=3 random subsampling matrices, identity source delays + NSO channel delays,
num_repeat=R (default in DEFAULT_REPEAT), noiseless; one *step* = one full transform of one fresh signal:
query lattice (K1) -> synthetic evaluation (K2) -> batched q-ary DFT (K3) -> peeling (K4), exact support recovery
checked on the last step.  Prints ONE JSON line (see README / DESIGN.md for the keys).

At N = 1 the line also carries "extras": untimed side checks run AFTER the measurement in subprocesses (pending GPU tests,
A/B timings of opt-in kernel variants); `--no-extras` skips them.

`--impl reference` times the CPU oracle port of the reference NumPy path (oracle/qsft_oracle.py; the reference is
pure Python and cannot travel to the GPU box) on bounded samples of the same workload and extrapolates.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q, N_DIM, B_DIM, SPARSITY, C_SUB = 4, 40, 10, 100_000, 3
DEFAULT_REPEAT = 1


def log(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--repeat", type=int, default=DEFAULT_REPEAT, help="num_repeat R of the NSO delays")
    ap.add_argument("--sparsity", type=int, default=SPARSITY)
    ap.add_argument("--b", type=int, default=B_DIM)
    ap.add_argument("--n", type=int, default=N_DIM)
    ap.add_argument("--eval-impl", type=int, default=0, help="K2 kernel: 0 auto, 1 SIMT, 2 tcgen05")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the untimed side checks run in subprocesses after the measurement (N = 1 only)")
    ap.add_argument("--peel-mode", default="auto", choices=["auto", "sharded", "sharded_host", "replicated"],
                    help="multi-GPU peeling: bin-sharded on-device loop (finds exchanged inside the kernel), bin-sharded with one "
                         "NCCL all-gather per round, or replicated on every rank")
    return ap.parse_args()


def query_args(R, b):
    return {"query_method": "complex", "num_subsample": C_SUB, "delays_method_source": "identity",
            "subsampling_method": "qsft", "delays_method_channel": "nso", "num_repeat": R, "b": b}


def workload_name(a):
    return (f"synthetic q={Q} n={a.n} b={a.b} S={a.sparsity} C={C_SUB} identity/nso num_repeat={a.repeat} noiseless "
            f"(BASELINE configs[4])")


# ------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(s) > 2 + i and s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# CPU baseline = oracle port of the reference NumPy path, bounded samples, extrapolated
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_transform_seconds(a, budget_s=20.0, seed=0):
    """Returns (estimated seconds per full transform on this host, detail dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qsft_oracle as orc
    cores = os.cpu_count() or 1
    R, b, n, S = a.repeat, a.b, a.n, a.sparsity
    P = R * (n + 1)
    B = Q ** b
    G = C_SUB * P
    rng_state = np.random.get_state()
    np.random.seed(seed)
    sw, locq, strengths = orc.generate_signal_w(n, Q, S, 1, 1)
    Ms, Ds = orc.get_Ms_and_Ds(n, Q, **query_args(R, b))
    # (1) sampling: time batches of 10 000 queries (the reference's batch size) of the first lattice, all cores
    dig = orc.query_digits(Ms[0], Ds[0][0][:1], Q)[0].T                     # (B, n) digit rows of delay row 0
    idx = orc.qary_vec_to_dec(dig[: min(B, 200_000)].T, Q)
    # the reference evaluates 10 000 queries per batch (a 10 000 x S complex128 matrix = 16 GB at S = 1e5); the port
    # uses the same arithmetic on smaller batches so that `cores` threads fit in a few GB
    batch = int(max(1, min(10_000, 2e7 // S)))
    t0 = time.time()
    done = 0
    chunk = batch * cores
    while time.time() - t0 < budget_s * 0.6 and done < len(idx):
        orc.synth_subsample(idx[done:done + chunk], locq, strengths, Q, n, batch=batch, threads=cores)
        done += min(chunk, len(idx) - done)
    t_sample = time.time() - t0
    pairs_per_s = done * S / t_sample
    est_sampling = G * B * S / pairs_per_s
    # (2) index generation for one delay row (python big ints) and (3) FFT of one q^b block
    t0 = time.time()
    orc.qary_vec_to_dec(dig.T, Q)
    t_idx = time.time() - t0
    x = np.random.normal(size=B) + 1j * np.random.normal(size=B)
    t0 = time.time()
    orc.gwht(x, Q, b)
    t_fft = time.time() - t0
    # (4) peel on a reduced instance with the same bin load S/B (b-2 => 16x fewer bins and balls; the reference loop
    #     costs time proportional to bins + balls), scaled back.  Keeps the host memory of the baseline small.
    b_s = max(2, b - 2)
    scale = (Q ** b) / (Q ** b_s)
    S_peel = max(50, int(S / scale))
    sw2, locq2, st2 = orc.generate_signal_w(n, Q, S_peel, 1, 1)
    Ms_s = [M[:, :b_s] for M in Ms]
    osig = orc.OracleSignal(n, Q, query_args(R, b_s), locq2, st2, 0.0, sw2, Ms=Ms_s, Ds=Ds, use_closed_form=True)
    t0 = time.time()
    res = orc.transform(osig, C_SUB, R, b_s, "identity", "nso")
    t_peel = (time.time() - t0) * scale
    np.random.set_state(rng_state)
    total = est_sampling + G * (t_idx + t_fft) + t_peel
    detail = {"sampling_pairs_per_s": pairs_per_s, "sampled_queries": done, "est_sampling_s": est_sampling,
              "index_s_per_row": t_idx, "fft_s_per_block": t_fft, "est_peel_s": t_peel,
              "peel_sample_recovered": len(res) == len(sw2)}
    sample = (f"{done} queries x S={S} timed in {t_sample:.1f}s on {cores} threads (extrapolated to G*B={G * B} queries)"
              f" + 1 index row + 1 q^b FFT (x{G}) + peel at b={b_s}, S={S_peel} (same bin load) scaled x{scale:.0f}")
    return total, detail, sample, cores


def reference_transform_seconds(a, budget_s=20.0, seed=0):
    """The same estimate with the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py; /root/reference in the build
    container) doing the work: its `sampling_function` closure (synthetic_signal.py:97-101) on bounded batches of queries --
    the reference's own batch of 10 000 queries x S = 1e5 complex128 is 16 GB per worker, so batches are cut to fit in
    memory and run on a thread pool --, its `_get_qsft_query_indices` for one delay row, its `_compute_subtransform` for one
    row, and its `QSFT.transform` peel loop on a reduced instance with the same bin load (bins filled from the closed form).
    Returns None when no copy of the reference is available."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    root = ref_shim.reference_root()
    if root is None:
        return None
    ref_shim.install(root)
    import qsft_oracle as orc
    from concurrent.futures import ThreadPoolExecutor
    from synt_exp.synt_src.synthetic_signal import SyntheticSubsampledSignal as RefSignal
    from synt_exp.synt_src.synthetic_signal import generate_signal_w as ref_generate_signal_w
    from qsft.qsft import QSFT as RefQSFT

    class Sig(RefSignal):                      # the reference object without its constructor's full sampling run
        def _init_signal(self):
            self._set_Ms_and_Ds_qsft()

    cores = os.cpu_count() or 1
    R, b, n, S = a.repeat, a.b, a.n, a.sparsity
    P = R * (n + 1)
    B = Q ** b
    G = C_SUB * P
    rng_state = np.random.get_state()
    np.random.seed(seed)
    sw, locq, strengths = ref_generate_signal_w(n, Q, S, 1, 1, 0, full=False)
    sig = Sig(n=n, q=Q, query_args=query_args(R, b), signal_w=sw, locq=locq, strengths=strengths, noise_sd=0.0)
    # (1) indices of one delay row of the first lattice (object-dtype big-int arithmetic of the reference)
    t0 = time.time()
    idx_rows = sig._get_qsft_query_indices(sig.Ms[0], np.asarray(sig.Ds[0][0])[:1])
    t_idx = time.time() - t0
    idx = np.asarray(idx_rows[0][: min(B, 200_000)], dtype=object)
    # (2) sampling: the reference's closure on batches that fit in memory, all cores
    batch = int(max(1, min(10_000, 2e7 // S)))
    t0 = time.time()
    done = 0
    with ThreadPoolExecutor(cores) as pool:
        while time.time() - t0 < budget_s * 0.6 and done < len(idx):
            chunk = idx[done:done + batch * cores]
            list(pool.map(sig.sampling_function, [chunk[i:i + batch] for i in range(0, len(chunk), batch)]))
            done += len(chunk)
    t_sample = time.time() - t0
    pairs_per_s = done * S / t_sample
    est_sampling = G * B * S / pairs_per_s
    # (3) transform of one q^b row
    x = (np.random.normal(size=B) + 1j * np.random.normal(size=B))[None, :]
    t0 = time.time()
    sig._compute_subtransform(x, b)
    t_fft = time.time() - t0
    # (4) the reference's peel loop on a reduced instance with the same bin load
    b_s = max(2, b - 2)
    scale = (Q ** b) / (Q ** b_s)
    S_peel = max(50, int(S / scale))
    sw2, locq2, st2 = ref_generate_signal_w(n, Q, S_peel, 1, 1, 0, full=False)
    small = Sig(n=n, q=Q, query_args=query_args(R, b_s), signal_w=sw2, locq=locq2, strengths=st2, noise_sd=0.0)
    osig = orc.OracleSignal(n, Q, query_args(R, b_s), locq2, st2, 0.0, sw2, Ms=small.Ms, Ds=small.Ds, use_closed_form=True)
    small.Us = [[{b_s: [np.array(row) for row in osig.Us[i][j][b_s]]} for j in range(R)] for i in range(C_SUB)]
    small.transformTimes = [[{b_s: 0.0} for j in range(R)] for i in range(C_SUB)]
    t0 = time.time()
    res = RefQSFT(num_subsample=C_SUB, num_repeat=R, b=b_s, reconstruct_method_source="identity",
                  reconstruct_method_channel="nso").transform(small, verbosity=0)
    t_peel = (time.time() - t0) * scale
    np.random.set_state(rng_state)
    total = est_sampling + G * (t_idx + t_fft) + t_peel
    detail = {"sampling_pairs_per_s": pairs_per_s, "sampled_queries": done, "est_sampling_s": est_sampling,
              "index_s_per_row": t_idx, "fft_s_per_row": t_fft, "est_peel_s": t_peel,
              "peel_sample_recovered": len(res) == len(sw2), "reference_copy": os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}
    sample = (f"UNMODIFIED reference functions: sampling_function on {done} queries x S={S} in {t_sample:.1f}s on {cores} threads "
              f"(batches of {batch}; extrapolated to G*B={G * B} queries) + _get_qsft_query_indices for 1 row + "
              f"_compute_subtransform for 1 row (x{G}) + QSFT.transform at b={b_s}, S={S_peel} (same bin load) scaled x{scale:.0f}")
    return total, detail, sample, cores


# ------------------------------------------------------------------------------------------------------------
def make_inputs(a, steps, seed0=1000):
    """Host-side synthetic inputs per step (support + strengths + Ms/Ds), generated outside the timed regions."""
    from qsft_b200.synthetic_signal import generate_signal_w
    from qsft_b200.query import get_Ms_and_Ds
    out = []
    for s in range(steps):
        np.random.seed(seed0 + s)
        sw, locq, strengths = generate_signal_w(a.n, Q, a.sparsity, 1, 1, 0, full=False)
        Ms, Ds = get_Ms_and_Ds(a.n, Q, **query_args(a.repeat, a.b))
        out.append((sw, locq, strengths, Ms, Ds))
    return out


def run_ours(a):
    import torch
    import torch.distributed as td
    import qsft_b200
    from qsft_b200 import _lib, ops
    from qsft_b200.dist import DistContext
    from qsft_b200.utils import padded_ld

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)
        dist = DistContext(peel_mode=a.peel_mode)

    def barrier():
        if dist is not None:
            td.barrier()
        torch.cuda.synchronize()

    R, b, n, S = a.repeat, a.b, a.n, a.sparsity
    P = R * (n + 1)
    B = Q ** b
    G = C_SUB * P
    ld = padded_ld(n)
    total_steps = a.warmup + a.steps
    inputs = make_inputs(a, total_steps)
    qa = query_args(R, b)

    def build_signal(inp, resident):
        sw, locq, strengths, Ms, Ds = inp
        kw = {}
        if resident is not None:
            kw = {"loc_dev": resident[0], "a_dev": resident[1]}
        # Ms / Ds were drawn in make_inputs (outside the timed region) and are handed over as kwargs
        sig = qsft_b200.SyntheticSubsampledSignal(signal_w=sw, locq=locq, strengths=strengths, noise_sd=0.0, n=n, q=Q,
                                                  query_args=dict(qa), device=dev, dist=dist, Ms=Ms, Ds=Ds,
                                                  noise_rng="device", eval_impl=a.eval_impl, **kw)
        return sig

    def transform(sig, output):
        sft = qsft_b200.QSFT(num_subsample=C_SUB, num_repeat=R, b=b, reconstruct_method_source="identity",
                             reconstruct_method_channel="nso")
        return sft.transform(sig, output=output), sft

    # ---- loop A: inputs resident in HBM, CUDA-event timed -------------------------------------------------
    resident = []
    for inp in inputs:
        resident.append((ops.pad_digits(np.asarray(inp[1]).T, ld, dev),
                         torch.from_numpy(np.asarray(inp[2]).astype(np.complex64)).to(dev)))
    log(f"inputs ready; warm-up x{a.warmup}")
    out = sft = None
    for s in range(a.warmup):
        t0 = time.time()
        # like the timed loop below, the previous step's result stays referenced while the next signal is built: the
        # caching allocator then holds two U buffers BEFORE the timed region (round 1's recurring slow second step was the
        # cudaMalloc of that second buffer inside it)
        out, sft = transform(build_signal(inputs[s], resident[s]), "device")
        torch.cuda.synchronize()
        log(f"warm-up step {s}: {time.time() - t0:.2f}s")
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    ops.TIMERS.reset(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_marks = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev0.record()
    step_marks[0].record()
    # output="device_async": nothing is read back inside a step, so the host queues step s + 1 while the GPU runs step s (the
    # synchronous call leaves the GPU idle for the ~1-2 ms the host needs to set the next transform up); every transform of
    # the timed region has completed on the device when `barrier()` returns, and its outcome is checked right after
    pending = []
    for s in range(a.warmup, total_steps):
        out, sft = transform(build_signal(inputs[s], resident[s]), "device_async")
        pending.append(out)
        step_marks[s - a.warmup + 1].record()
    ev1.record()
    barrier()
    for s, h in enumerate(pending):
        st_ = h.wait()
        if st_["distinct"] != len(inputs[a.warmup + s][0]) or st_["rounds"] < 1:
            raise SystemExit(f"device-resident loop, step {s}: {st_} but the support has {len(inputs[a.warmup + s][0])} coefficients")
    device_rounds = pending[-1].stats["rounds"]
    ms_total = ev0.elapsed_time(ev1)
    per_step = [step_marks[i].elapsed_time(step_marks[i + 1]) for i in range(a.steps)]
    log("per-step device ms: " + " ".join(f"{v:.1f}" for v in per_step))
    launches = _lib.launch_count()
    kt = ops.TIMERS.totals()
    ops.TIMERS.reset(False)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / a.steps
    log(f"device-resident loop: {ms_per_step:.1f} ms/step; kernels: " + ", ".join(f"{k}={v[0] / a.steps:.1f}ms" for k, v in kt.items()))

    # ---- loop B: end to end through the public API with host buffers (H2D + D2H inside) -------------------
    # result container: host arrays (locations (K, n) int8, values (K,) complex128) -- output="arrays"; the
    # reference-compatible dict {tuple(k): complex} is timed separately below (pure-Python object construction).
    # bytes uploaded per step by ONE rank: the support digits (its 1 / N slice when the ranks stage the table together,
    # ops.pad_digits_sharded; the rest arrives over NVLink), the strengths, Ms / Ds / M^T rows
    from qsft_b200 import ops as _ops
    shared_staging = world > 1 and S >= _ops.SHARD_PACK_MIN_ROWS
    h2d = (-(-S // world) if shared_staging else S) * ld + S * 8 + C_SUB * (n * b + P * ld + b * ld)

    def e2e_loop(output):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        last_ = None
        for s_ in range(a.warmup, total_steps):
            sig_ = build_signal(inputs[s_], None)
            res_, sft_ = transform(sig_, output)
            symm_ = getattr(sig_, "_symm", None)
            xch_ = "" if symm_ is None else ("scatter" if getattr(sig_, "_U_scattered", False) else ("mcast" if symm_[3] else "p2p"))
            last_ = (res_, inputs[s_][0], sft_.last_stats, xch_)
        e1.record()
        barrier()
        ms = max(e0.elapsed_time(e1), (time.time() - t_wall0) * 1e3)
        if dist is not None:
            t_ = torch.tensor([ms], dtype=torch.float64, device=dev)
            td.all_reduce(t_, op=td.ReduceOp.MAX)
            ms = float(t_.item())
        return ms / a.steps, last_

    # untimed: first-use costs of the host path (pinned buffers; with several ranks the SECOND symmetric U buffer, which is
    # only needed once a finished signal is still referenced while the next one is built, exactly as in the loops below)
    keep = None
    for out_kind in ("arrays", "dict", "arrays"):
        sig_w = build_signal(inputs[0], None)
        keep = (transform(sig_w, out_kind), sig_w, keep[1] if keep else None)
    del keep
    e2e_ms_per_step, last = e2e_loop("arrays")
    log(f"end-to-end loop (arrays result): {e2e_ms_per_step:.1f} ms/step")
    e2e_dict_ms_per_step, last_d = e2e_loop("dict")
    log(f"end-to-end loop (dict result): {e2e_dict_ms_per_step:.1f} ms/step")
    res, sw, stats, used_symm = last
    d2h = stats["distinct"] * (n + 16 + 4)          # digit rows, complex128 means, counts (qsft_peel_distinct)
    got = dict(zip(map(tuple, res["locations"].tolist()), res["values"].tolist()))
    recovered = set(got.keys()) == set(sw.keys()) and set(last_d[0].keys()) == set(sw.keys())
    max_err = max(abs(got[k] - v) for k, v in sw.items()) if recovered else None

    if rank != 0:
        if dist is not None:
            td.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K2 evaluation) ---------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "2 x measured sustained bf16 (MEASURED_PEAKS.json), no int8 figure measured" if peaks else "2 x fallback 1.4 PF"
    sparse_mode = 0 if os.environ.get("QSFT_LATTICE_SPARSE", "2") == "0" else 2
    peak_mult = 2.0                                # dense int8 = 2 x bf16
    if "k2_eval_lattice" in kt:
        # lattice-factorised evaluation: the GEMM is 2 * (2*4^b1) * 4^b2 * (2S) * 3 limbs = 24 (dense-equivalent) int8 ops per
        # (query, support) pair; the launch also contains the operand-generation kernels.  The default
        # kernel runs it as a 2:4 structured-sparse MMA (A' has one zero per byte pair): half of those ops are executed and
        # the matching peak is the sparse int8 rate = 2 x dense int8 = 4 x bf16.
        k2_name, ops_per_pair = "k2_eval_lattice", 24.0
        if sparse_mode:
            peak_mult = 4.0
            peak_src = peak_src.replace("2 x", "4 x (2:4 sparse int8 = 2 x dense int8 = 4 x bf16)")
            k2_desc = ("k2_eval_lattice (K=S tcgen05.mma.sp int8 GEMM on CTA pairs, 2:4-sparse exact +-1/+-i operand "
                       + ("generated in tensor memory" if sparse_mode == 2 else "compressed in HBM")
                       + " x 3-limb int8 operand; 24 dense-equivalent int8 ops/pair, 12 executed)")
        else:
            k2_desc = "k2_eval_lattice (K=S tcgen05 int8 GEMM, exact +-1/+-i operand x 3-limb int8 operand; 24 int8 ops/pair)"
    else:
        k2_name, ops_per_pair = "k2_eval", 2.0 * n
        k2_desc = "k2_eval (tcgen05 int8 contraction K=n + root-of-unity epilogue; 2n int8 ops/pair)"
    k2_ms, k2_calls, k2_pairs = kt.get(k2_name, (0.0, 0, 0))
    ops_per_launch = ops_per_pair * (k2_pairs / max(1, k2_calls))      # algorithmic int8 ops of one launch
    achieved = ops_per_launch / (k2_ms / max(1, k2_calls) * 1e-3) / 1e12 if k2_ms > 0 else 0.0
    # DRAM traffic of the GEMM launch from the committed ncu capture (same shape only), else null
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", {0: "r1_k2l_traffic.json", 1: "r1_k2l_sparse_hbm_traffic.json",
                                                            2: "r1_k2l_sparse_traffic.json"}[sparse_mode])))
        if k2_name == "k2_eval_lattice" and a.gpus == 1 and (n, b, S) == (40, 10, 100_000):
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": k2_desc, "bound": "tensor",
                "achieved": achieved, "peak": peak_mult * bf16, "unit": "TFLOP/s", "frac": achieved / (peak_mult * bf16),
                "traffic": traffic, "peak_source": peak_src, "int8_ops_per_pair": ops_per_pair,
                # SURVEY 8(d) counts the plain contraction, 2 n int8 ops per (query, support) pair; the lattice-factorised
                # kernel needs 24.  `achieved` uses the kernel's own (smaller) count; the same launch at the survey's count:
                "survey_ops_per_pair": 2.0 * n, "achieved_at_survey_count": achieved * (2.0 * n) / ops_per_pair,
                "pairs_per_s": k2_pairs / (k2_ms * 1e-3) if k2_ms > 0 else None,
                "share_of_step": (k2_ms / a.steps) / ms_per_step if ms_per_step else None,
                "per_kernel_ms_per_step": {k: v[0] / a.steps for k, v in kt.items()}}
    if k2_name == "k2_eval_lattice" and sparse_mode and clocks and clocks.get("sm_mhz"):
        # second denominator: the tensor pipe's measured ISSUE rate (tools/sp_probe.cu on this pool: one N=256 + one N=128
        # sparse kind::i8 MMA with M=128, K=64 logical = 2*128*384*64 dense-equivalent ops every 192.1 cycles per SM)
        # at the SM clock actually sampled during the timed region (the board power cap holds it below the 1965 MHz maximum)
        per_clk = 2.0 * 128 * 384 * 64 / 192.1
        issue_peak = per_clk * 148 * clocks["sm_mhz"] * 1e6 / 1e12
        roofline["tensor_issue_rate"] = {"ops_per_clk_per_sm": per_clk, "sm_mhz": clocks["sm_mhz"], "peak": issue_peak,
                                         "unit": "TFLOP/s", "frac": achieved / issue_peak if issue_peak else None,
                                         "source": "tools/sp_probe.cu (profiles/README.md), clock = median nvidia-smi sample under load"}
    # the two HBM-bound kernels of the step, against the measured copy bandwidth: K3 = 16 B per bin (1 read + 1 write),
    # K4 = 8 B per bin and ROUND (one scan of U per round, what the reference does)
    hbm = peaks.get("hbm_gbs_sustained") or peaks.get("hbm_gbs", 6650.0)
    k3_ms, k3_calls, k3_elems = kt.get("k3_gwht", (0.0, 0, 0))
    if k3_ms > 0:
        gbs = 16.0 * k3_elems / (k3_ms * 1e-3) / 1e9
        roofline["k3_gwht_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                   "bytes_per_bin": 16, "ms_per_step": k3_ms / a.steps}
    k4_ms, k4_calls, k4_elems = kt.get("k4_peel", (0.0, 0, 0))
    if k4_ms > 0:
        rounds = int(device_rounds)
        gbs = 8.0 * k4_elems * rounds / (k4_ms * 1e-3) / 1e9
        roofline["k4_peel_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                   "bytes_per_bin_and_round": 8, "rounds": rounds, "ms_per_step": k4_ms / a.steps,
                                   "host_syncs_per_peel": 0}
        try:                                               # DRAM bytes of one peel from the committed ncu capture (config 5 only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2", "r2z_k4_traffic.json")))
            if (a.n, a.b, a.sparsity, a.repeat, a.gpus) == (N_DIM, B_DIM, SPARSITY, 1, 1):
                roofline["k4_peel_hbm"]["traffic"] = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
                roofline["k4_peel_hbm"]["traffic_note"] = tj["note"]
        except Exception:
            pass

    line = {
        "metric": "q-SFT transforms/sec (sample + FFT + peel)", "value": 1e3 / ms_per_step, "unit": "transforms/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8",
        "dtype_detail": "int8 x int8 -> int32 contraction (exact), limbs recombined in f64, samples / bins complex64 (f32)",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "groups_G": G, "bins_B": B, "samples": G * B,
                   "l2": "inputs larger than L2 (per-step working set >= 1 GB)",
                   "device_loop": "transforms queued back to back, no host read-back inside a step (QSFT.transform output='device_async'); every step's outcome (rounds, distinct k == support size) checked after the timed region; e2e is the synchronous call",
                   "support_recovered_exactly": recovered,
                   "max_coeff_err": max_err, "eval_impl": a.eval_impl,
                   "parallelism": (f"delay rows sharded over {a.gpus} GPU(s), U exchanged by "
                                   + ({"scatter": "K3 scatter stores into the peers' symmetric U buffers: every element to the ONE rank that owns its bin (fused all-to-all)",
                                       "mcast": "K3 multimem.st into the NVLS multicast mapping of the symmetric U buffers (fused all-gather)",
                                       "p2p": "K3 unicast peer stores into symmetric memory (fused all-gather)"}.get(used_symm)
                                      or "NCCL all-gather")
                                   + {"device": ", bin-sharded on-device peel loop: the round's finds exchanged inside the kernel over NVLink",
                                      "host": ", bin-sharded peel with one NCCL all-gather of finds per round",
                                      "": ", peel replicated on every rank (no collective)"}[dist.shard_peel(8 * G * B)])
                   if a.gpus > 1 else "single GPU"},
        "e2e": {"value": 1e3 / e2e_ms_per_step, "unit": "transforms/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "bytes_are": "per rank (every rank uploads its inputs and reads the result back)",
                "result": "host arrays (locations, values); output='arrays'",
                "value_with_reference_dict_result": 1e3 / e2e_dict_ms_per_step},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "sample_fft_gbs": (24.0 * G * B * a.steps) / ((k2_ms + k3_ms) * 1e-3) / 1e9 if (k2_ms + k3_ms) > 0 else None,
    }
    if a.gpus == 1 and not a.no_cpu_baseline:
        log("timing the CPU reference (bounded sample)")
        kind, got = "reference", None
        try:
            got = reference_transform_seconds(a)
        except Exception as exc:                # e.g. a missing module of the reference's environment on this box
            log(f"reference functions not usable here ({exc!r}): timing the oracle port instead")
        if got is None:
            kind, got = "port", cpu_reference_transform_seconds(a)
        secs, detail, sample, cores = got
        line["cpu_baseline"] = {"value": 1.0 / secs, "unit": "transforms/s", "cores": cores, "kind": kind,
                                "sample": sample, "detail": detail}
    if a.gpus == 1 and not a.no_extras:
        try:
            torch.cuda.empty_cache()
            line["extras"] = run_extras()
            # the K2L issue-rate denominator from THIS box's probe instead of the constant measured earlier on this pool
            cyc = (line["extras"].get("int8_issue_probe", {}).get("cycles_per_kstep", {}) or {}).get("grid 148 sparse K=64 A from TMEM")
            tir = roofline.get("tensor_issue_rate")
            if cyc and tir:
                per_clk = 2.0 * 128 * 384 * 64 / cyc
                peak = per_clk * 148 * tir["sm_mhz"] * 1e6 / 1e12
                tir.update({"ops_per_clk_per_sm": per_clk, "peak": peak, "frac": roofline["achieved"] / peak,
                            "cycles_per_kstep_measured_in_this_run": cyc,
                            "source": "tools/sp_probe.cu run after the measurement on this box (extras.int8_issue_probe), "
                                      "clock = median nvidia-smi sample under load"})
        except Exception as exc:            # the side checks must never cost the measurement
            line["extras"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        td.destroy_process_group()


def run_extras():
    """Untimed side measurements AFTER the main one, each in its own subprocess (a failure there cannot touch the numbers
    above; 2.5 minutes in total): the int8 tensor-pipe issue rate measured on this box (second denominator of the K2L
    roofline), the stand-alone HBM kernels, the other BASELINE configs end to end.  Nothing here feeds `value`."""
    deadline = time.time() + 150.0

    def sub(cmd, env=None, timeout=120):
        t0 = time.time()
        timeout = min(timeout, deadline - t0)
        if timeout < 15:
            return -8, "", "skipped: side-measurement time budget used up", 0.0
        try:
            r = subprocess.run(cmd, cwd=ROOT, env={**os.environ, **(env or {})}, capture_output=True, text=True, timeout=timeout)
            return r.returncode, r.stdout, r.stderr, time.time() - t0
        except subprocess.TimeoutExpired:
            return -9, "", "timeout", time.time() - t0

    out = {"note": "untimed side measurements run after the main one in subprocesses; not part of value / e2e"}
    py = sys.executable
    # 1. tools/sp_probe.cu: cycles per k-step (one N=256 + one N=128 sparse kind::i8 MMA, M=128, K=64 logical) with the
    #    exact operand resident in tensor memory -- the issue floor of the K2L main loop, measured on this box
    exe = "/tmp/qsft_sp_probe"
    rc, so, se, dt = sub(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe, "tools/sp_probe.cu"], timeout=60)
    if rc == 0:
        rc, so, se, dt = sub([exe], timeout=40)
        rates = {}
        for ln in so.splitlines():
            if "cycles / k-step" in ln and ln.startswith("grid 148"):
                rates[ln.split(":")[0].strip()] = float(ln.split(":")[1].split()[0])
        out["int8_issue_probe"] = {"cycles_per_kstep": rates, "rc": rc}
    else:
        out["int8_issue_probe"] = {"rc": rc, "stderr": se[-300:]}
    # 2. stand-alone timings of the HBM-bound kernels (K1 / K3 / K4; L2 flushed between iterations)
    rc, so, se, dt = sub([py, "tools/microbench.py", "--only", "k1,k3,k4"], timeout=60)
    try:
        out["microbench"] = json.loads(so)
    except Exception:
        out["microbench"] = {"rc": rc, "stderr": se[-300:]}
    # 3. the other BASELINE configs (and config 5 at num_repeat = 3) through the public API: best of 3 signals each
    rc, so, se, dt = sub([py, "tools/bench_configs.py"], timeout=100)
    rows = []
    for ln in so.splitlines():
        try:
            rows.append(json.loads(ln))
        except Exception:
            pass
    out["baseline_configs_e2e"] = rows if rows else {"rc": rc, "stderr": se[-300:]}
    return out


def run_reference(a):
    """Reference arm: the oracle port of the reference's CPU path, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    times = []
    detail = sample = None
    cores = os.cpu_count() or 1
    kind = "reference"
    for s in range(a.warmup + a.steps):
        got = None
        if kind == "reference":
            try:
                got = reference_transform_seconds(a, budget_s=8.0, seed=s)
            except Exception as exc:
                log(f"reference functions not usable here ({exc!r}): timing the oracle port instead")
            if got is None:
                kind = "port"
        if got is None:
            got = cpu_reference_transform_seconds(a, budget_s=8.0, seed=s)
        secs, detail, sample, cores = got
        if s >= a.warmup:
            times.append(secs)
    secs = float(np.mean(times))
    value = 1.0 / secs
    line = {"impl": "reference", "metric": "q-SFT transforms/sec (sample + FFT + peel)", "value": value,
            "unit": "transforms/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": secs * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128 / int64 (NumPy)",
            "data": "synthetic", "config": {"workload": workload_name(a)},
            "cpu_baseline": {"value": value, "unit": "transforms/s", "cores": cores, "kind": kind, "sample": sample,
                             "detail": detail},
            "e2e": {"value": value, "unit": "transforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
