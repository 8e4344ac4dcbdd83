#!/usr/bin/env python
"""bench.py — DQN transitions trained/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one Agent.Learn over one sampled batch.  Default workload = the wide
DQN config the metric's target is quoted on (BASELINE.json configs[2]: MLP
128->512->512->18, batch 8192, bf16 tcgen05 GEMMs, fp32 master weights);
N > 1 = batch-sharded data parallelism at configs[3]'s stated shape: global batch
65536 split over the ranks (strong scaling: 32768 / 16384 / 8192 rows per GPU at
N = 2 / 4 / 8), one gradient exchange per step; `--scaling weak` keeps 8192 rows
per GPU instead.  Prints ONE JSON line on rank 0; at N = 1 the line also carries
the CartPole batch-4096 and replica-sweep workloads (configs[1], configs[4]) under
"configs".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gold_b200 import synth  # noqa: E402

WORKLOADS = {
    # name: (dims, per-GPU batch, replay rows per GPU, replicas per GPU (0 = single agent), default precision)
    "wide": (synth.WIDE, 8192, 1 << 20, 0, "bf16"),
    "cartpole4096": (synth.CARTPOLE, 4096, 1 << 20, 0, "fp32"),
    "replicas": (synth.CARTPOLE, 20, 4096, 128, "fp32"),
}
TARGET_SYNC_EVERY = 100  # Hyperparameters.UpdateTargetSteps (agent.go:63)


def algorithmic_work(dims, B, replicas=1):
    """bytes/transition = 8D+13, parameter bytes/step = 28P,
    FLOPs/transition = 6*P_w + 2*(P_w - D*H1)."""
    D, H1, H2, A = dims
    P = synth.nparams(dims)
    Pw = D * H1 + H1 * H2 + H2 * A
    bytes_step = B * (8 * D + 13) + 28 * P
    flops_step = B * (6 * Pw + 2 * (Pw - D * H1))
    return bytes_step * replicas, flops_step * replicas


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            p.update({k: j[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in j})
            p["source"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions (B200_PROFILING.md).  The resident leg lasts
    a few milliseconds, far below nvidia-smi's sampling period, so NVML is polled directly from a thread
    (about every millisecond); samples are kept for the spans marked with begin()/end() only.  Falls back
    to `nvidia-smi -lms` (all samples of the run) when pynvml is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.spans, self.samples = index, [], None, None, [], []
        self._stop = False
        self._in = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # the process may see a subset of the box's GPUs: resolve the CUDA device through its UUID/visible list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def begin(self):
        self._in = True

    def end(self):
        self._in = False

    def _poll(self):
        nv = self.nv
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                self.samples.append((self._in, mhz, rs))
            except Exception:
                pass
            time.sleep(0.0005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv:
            self._stop = True
            self.t.join(timeout=2)
            inside = [x for x in self.samples if x[0]] or self.samples
            sm = [x[1] for x in inside]
            bits = 0
            for x in inside:
                bits |= x[2]
            reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "samples": len(sm),
                    "samples_total": len(self.samples), "reasons": reasons,
                    "how": "NVML polled from a thread, samples inside the timed regions (resident + e2e legs)"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons), "how": "nvidia-smi -lms 100 over the whole run"}


def pinned_array(G, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = G.lib().goldq_host_alloc(n)
    if not p:
        raise RuntimeError("goldq_host_alloc failed")
    buf = (C.c_uint8 * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


# --------------------------------------------------------------------------- reference arm
def cpu_rows_for_budget(dims, cap_rows, steps, threads, budget_s):
    """Rows per CPU step so that `steps` steps fit the time budget (bounded sample)."""
    probe = min(cap_rows, 128 if dims == synth.WIDE else 2048)
    rate, _ = cpu_learn_rate(dims, probe, 1, 1, threads)
    rows = int(rate * budget_s / max(steps, 1))
    rows = max(min(rows, cap_rows), min(cap_rows, 64))
    return (rows // 8) * 8 if rows >= 8 else rows


def cpu_learn_rate(dims, rows, steps, warmup, threads):
    """Times the oracle port of Agent.Learn (reference structure: 2*rows single-row
    Predicts + one FitBatch + multi-pass Adam) on `rows` sampled transitions."""
    from oracle import binding as O
    O.set_threads(threads)
    P = synth.nparams(dims)
    N = max(4 * rows, 4096)
    data = synth.replay(dims, N)
    th = synth.glorot_params(dims, synth.SEED_ONLINE)
    tt = synth.glorot_params(dims, synth.SEED_TARGET)
    m = np.zeros(P, np.float32); v = np.zeros(P, np.float32)
    it = 0
    times = []
    for step in range(warmup + steps):
        idx = synth.sample_indices(N, rows, step)
        t0 = time.perf_counter()
        s, a, r, s2, d = O.gather(dims[0], *data, idx)
        out = O.learn(dims, rows, th, tt, m, v, it, s, a, r, s2, d, want_debug=False)
        it = out["iter"]
        if (step + 1) % TARGET_SYNC_EVERY == 0:
            tt = th.copy()
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
    O.set_threads(1)
    return rows * len(times) / sum(times), float(np.mean(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as O
    dims, B, _, replicas, prec = WORKLOADS[args.workload]
    # torchrun exports OMP_NUM_THREADS=1, so omp_get_max_threads() would say 1 on a 24-core box: take the
    # cores this process may run on and set the team size explicitly
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(1, min(threads, 64))
    if replicas:
        rows = 20
    else:
        rows = args.cpu_rows or cpu_rows_for_budget(dims, B, args.steps + 2, threads, args.cpu_budget)
    rate, sec = cpu_learn_rate(dims, rows, args.steps, min(args.warmup, 2), threads)
    line = {
        "impl": "reference", "metric": "dqn_transitions_trained_per_sec", "value": rate, "unit": "transitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, dims, rows_per_gpu(args, B, replicas), replicas, prec),
        "cpu_baseline": {"value": rate, "unit": "transitions/s", "cores": threads, "kind": "port",
                         "sample": f"{rows}-row learn steps of the same workload (oracle port of the Gorgonia "
                                   f"learn step: per-event Predicts + FitBatch + multi-pass Adam), OpenMP over rows; "
                                   f"the Go reference cannot run here (no Go toolchain)"},
        "e2e": {"value": rate, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def scaling_of(args):
    """N = 1: nothing to scale ("weak" by convention).  N > 1, wide workload: strong by default (configs[3]:
    global batch 65536 over the ranks); the other workloads keep a fixed per-GPU batch (weak)."""
    if args.gpus > 1 and args.workload == "wide":
        return args.scaling or "strong"
    return "weak"


def rows_per_gpu(args, B, replicas):
    if not replicas and scaling_of(args) == "strong":
        if args.global_batch % args.gpus:
            raise SystemExit("--global-batch must divide evenly over the GPUs")
        return args.global_batch // args.gpus
    return B


def workload_config(args, dims, B, replicas, prec, name=None):
    n = args.gpus
    name = name or args.workload
    cfg = {"workload": {"wide": "wide DQN MLP 128->512->512->18, learn step (BASELINE configs[2]; N>1: configs[3] data-parallel)",
                        "cartpole4096": "CartPole DQN 4->24->24->2 learn step, batch 4096 (BASELINE configs[1])",
                        "replicas": "independent CartPole DQN replicas, batch 20 each (BASELINE configs[4])"}[name],
           "dims": list(dims), "batch_per_gpu": B, "global_batch": B * n * max(replicas, 1),
           "replay_rows_per_gpu": WORKLOADS[name][2], "precision": prec,
           "parallelism": (f"dp{n}" if not replicas else f"replicas{replicas}x{n}"),
           "target_sync_every": TARGET_SYNC_EVERY,
           "l2": "inputs larger than L2: rows are drawn at random from a replay ring far above 126 MB"
                 if not replicas and dims == synth.WIDE else "replay rows drawn at random each step; weights stay L2-resident by design"}
    return cfg


# --------------------------------------------------------------------------- own arm
def build_handle(G, args, dims, B, N, replicas, prec, rank, world, local_rank):
    kw = dict(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
              replay_capacity=N, device=local_rank,
              precision=G.PREC_BF16 if prec == "bf16" else G.PREC_FP32)
    if replicas:
        kw["n_replicas"] = replicas
    else:
        kw["global_batch"] = B * world
    h = G.Handle(**kw)
    R = max(replicas, 1)
    th = np.concatenate([synth.glorot_params(dims, synth.SEED_ONLINE + 1000 * r) for r in range(R)])
    tt = np.concatenate([synth.glorot_params(dims, synth.SEED_TARGET + 1000 * r) for r in range(R)])
    h.set_params(G.NET_ONLINE, th)
    h.set_params(G.NET_TARGET, tt)
    # synthetic replay, pushed in chunks (host RAM stays bounded)
    chunk = min(N, 1 << 16)
    for c in range(0, N, chunk):
        n = min(chunk, N - c)
        parts = [synth.replay(dims, n, seed=synth.SEED_DATA + 7919 * (c // chunk) + 104729 * r + 15485863 * rank)
                 for r in range(R)]
        h.replay_push(*[np.concatenate([p[k] for p in parts]) for k in range(5)])
    h.sync()
    return h


class Workload:
    """One handle + the three measurements of it: resident steps (value), host-buffer steps (e2e),
    per-kernel breakdown (roofline)."""

    def __init__(self, G, args, name, rank, world, local_rank, dist, torch):
        self.G, self.args, self.name, self.rank, self.world = G, args, name, rank, world
        self.dist, self.torch = dist, torch
        dims, B, N, replicas, prec = WORKLOADS[name]
        if name == args.workload:
            B = rows_per_gpu(args, B, replicas)
            prec = args.precision or prec
            N = args.replay_rows or N
        self.dims, self.B, self.N, self.replicas, self.prec = dims, B, N, replicas, prec
        self.R = max(replicas, 1)
        self.rows = B * self.R                     # transitions per step on this GPU
        self.h = build_handle(G, args, dims, B, N, replicas, prec, rank, world, local_rank)
        if world > 1 and not replicas:
            uid = [G.Handle.dp_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            self.h.dp_init(uid[0], rank, world)
        # every step graph (32/28/../8/4/2/1 steps, host-index step) is instantiated here, before any warm-up, so
        # that --warmup is honoured as given and no instantiation lands in a timed region
        self.h.learn_prepare()
        self.seed0 = 1_000_003 * (rank + 1)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run_resident(self, first, count):
        """steps [first, first+count): replay indices drawn on the device (nothing crosses PCIe),
        enqueued through goldq_learn_many in chunks that end at the target-sync cadence."""
        h, i = self.h, first
        while i < first + count:
            n = min(first + count - i, TARGET_SYNC_EVERY - (i % TARGET_SYNC_EVERY), self.args.chunk)
            h.learn_many(n, self.seed0 + i)
            i += n
            if i % TARGET_SYNC_EVERY == 0:
                h.sync_target()

    def resident(self, K, W, sampler=None):
        """value: K steps with inputs resident in HBM; CUDA events on the handle's stream, max over ranks."""
        h = self.h
        self.run_resident(0, W)
        h.sync()
        self.barrier()
        l0 = h.launch_count()
        if sampler:
            sampler.begin()
        h.timer_start()
        self.run_resident(W, K)
        ms = h.timer_stop()
        if sampler:
            sampler.end()
        self.barrier()
        launches = h.launch_count() - l0
        ms = self.max_over_ranks(ms)
        return self.rows * self.world * K / (ms * 1e-3), ms, launches

    def e2e(self, K, sampler=None):
        """The same metric through the public C ABI with HOST buffers.  Every step: B fresh transitions
        from ONE pinned host buffer (goldq_replay_stage_packed: one H2D copy on the library's copy stream,
        started two steps ahead like an actor thread would; goldq_replay_commit appends the batch behind the
        step), the sampled index list from host memory (read by the device over PCIe), the step (goldq_learn),
        the loss read back (goldq_last_loss: the step's store into mapped host memory + sync).  Host wall clock."""
        G, h, dims, B, R, N = self.G, self.h, self.dims, self.B, self.R, self.N
        D = dims[0]
        nbytes = h.packed_bytes(B)
        bufs = []
        for k in range(3):
            arr, ptr = pinned_array(G, (nbytes,), np.uint8)
            parts = [synth.replay(dims, B, seed=4242 + k + 31 * r) for r in range(R)]
            fresh = [np.concatenate([p[j] for p in parts]) for j in range(5)]
            h.pack(*fresh, out=arr)
            bufs.append(ptr)
        idx_host = []
        for i in range(8):
            arr, ptr = pinned_array(G, (R, B), np.int32)
            arr[...] = np.stack([synth.sample_indices(N, B, 7 * i + r) for r in range(R)])
            idx_host.append(ptr)
        h2d = nbytes + R * B * 4
        d2h = 4 * R

        def step(i):
            # batch i+1 was staged during step i-1 and batch i+2 is staged now (three staging sets): the copies
            # overlap the steps' kernels; every byte still crosses PCIe inside the timed region, one batch per step
            h.learn_raw(idx_host[i % len(idx_host)])          # index list (host memory) + the step
            h.replay_commit()                                 # batch i+1 -> replay memory, behind step i
            h.replay_stage_packed(B, bufs[(i + 2) % 3])       # H2D of batch i+2 (copy stream)
            loss = h.last_loss()                              # the step's loss (mapped host memory) + sync
            if (i + 1) % TARGET_SYNC_EVERY == 0:
                h.sync_target()
            return loss

        h.replay_push_packed(B, bufs[0])
        h.replay_stage_packed(B, bufs[1])
        for i in range(3):
            step(i)
        self.barrier()
        if sampler:
            sampler.begin()
        t0 = time.perf_counter()
        for i in range(3, 3 + K):
            last = step(i)
        self.barrier()
        sec = self.max_over_ranks(time.perf_counter() - t0)
        if sampler:
            sampler.end()
        return self.rows * self.world * K / sec, h2d, d2h, last

    def breakdown(self):
        """[(label, us)] of one eager, serialised learn step (mean of 10), after 3 warm-up steps."""
        h = self.h
        for i in range(3):
            h.step_breakdown(self.seed0 + 10_000_000 + i)
        reps = [h.step_breakdown(self.seed0 + 20_000_000 + i) for i in range(10)]
        labels = [l for l, _ in reps[0]]
        kms = np.array([[t for _, t in r] for r in reps]).mean(axis=0)
        return labels, kms

    def close(self):
        self.h.close()


def wide_kernel_flops(dims, B):
    """algorithmic FLOPs of each tensor-core launch of the wide step (2 per MAC; online and target share
    the forward launch)"""
    D_, H1_, H2_, A_ = dims
    fwd = 2 * 2.0 * B * (D_ * H1_ + H1_ * H2_ + H2_ * A_)
    return {"fused forward L1-L3 online+target (tcgen05)": fwd,
            "fused forward L1-L3 online+target, CTA pairs (tcgen05 cta_group::2)": fwd,
            "fwd L1 online+target (tcgen05)": 2 * 2.0 * B * D_ * H1_,
            "fwd L2 online+target (tcgen05)": 2 * 2.0 * B * H1_ * H2_,
            "fwd L3 online+target (tcgen05)": 2 * 2.0 * B * H2_ * A_,
            "dW3 (tcgen05)": 2.0 * B * H2_ * A_, "dW2 split-K (tcgen05)": 2.0 * B * H1_ * H2_,
            "dX -> dZ1 (tcgen05)": 2.0 * B * H1_ * H2_, "dW1 split-K (tcgen05)": 2.0 * B * D_ * H1_}


def roofline_of(w, ms_per_step, pk, floor_us):
    G, h, dims, B, R = w.G, w.h, w.dims, w.B, w.R
    bytes_step, flops_step = algorithmic_work(dims, B, R)
    step_s = ms_per_step * 1e-3
    labels, kms = w.breakdown()
    breakdown = {labels[i]: round(float(kms[i]) * 1e3, 2) for i in range(len(labels))}   # us per launch
    if h.path == G.PATH_WIDE:
        kflops = wide_kernel_flops(dims, B)
        cand = [i for i, l in enumerate(labels) if l in kflops]
        top = max(cand, key=lambda i: kms[i])        # dominant = the tensor-core launch with the most time
        k_us = float(kms[top]) * 1e3
        how = "one eager step with a CUDA event after every launch (includes the launch gap of an isolated launch)"
        if labels[top].startswith("fused forward"):
            # average launch duration of that kernel over 20 back-to-back launches between two CUDA events
            # on the handle's stream (goldq_time_forward): the figure `achieved` is computed from
            k_us = h.time_forward_us(20)
            how = "20 back-to-back launches of the kernel between two CUDA events on the handle's stream"
        ach = kflops[labels[top]] / (k_us * 1e-6) / 1e12
        peak = pk["bf16_tflops"]      # burst figure: this kernel is timed alone, between events
        # DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
        # (dram__bytes_read.sum + dram__bytes_write.sum); only valid for the configuration profiled
        traffic = NCU_TRAFFIC.get((labels[top], tuple(dims), B))
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": NCU_TRAFFIC_SOURCE if traffic else None,
                "kernel": labels[top], "kernel_us": k_us, "kernel_us_how": how,
                "kernel_us_isolated_launch": float(kms[top]) * 1e3,
                "flops_per_launch": kflops[labels[top]],
                "peak_source": pk["source"] + " bf16_tflops (burst: kernel timed alone between CUDA events)",
                "step": {"flops_per_step": flops_step, "achieved_tflops": flops_step / step_s / 1e12,
                         "frac_of_sustained_peak": flops_step / step_s / 1e12 / pk["bf16_tflops_sustained"]},
                "kernels_us": breakdown}
    top = int(np.argmax(kms))
    ach = bytes_step / (kms[top] * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": None, "kernel": labels[top], "kernel_us": float(kms[top]) * 1e3,
            "bytes_per_launch": bytes_step, "peak_source": pk["source"] + " hbm_gbs",
            "launch_floor_us": floor_us,
            "frac_of_launch_floor": (floor_us / (step_s * 1e6)) if floor_us else None,
            "note": "launch-latency bound by construction: the step's algorithmic bytes take < 0.1 us at HBM "
                    "speed; launch_floor_us = one empty kernel per graph node, replayed from a 16-node graph",
            "kernels_us": breakdown}


def run_own(args):
    import torch
    import torch.distributed as dist

    from gold_b200 import _capi as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # (NCCL_DEBUG is left as the caller set it: stdout was diverted to stderr in main(), so NCCL's
        # banner cannot reach the one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pk = peaks()
    K, W = args.steps, args.warmup
    w = Workload(G, args, args.workload, rank, world, local_rank, dist, torch)
    h = w.h
    sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_clocks) else None
    if sampler:
        sampler.start()
    value, ms, launches = w.resident(K, W, sampler)
    Ke = max(3, min(K, args.e2e_steps))
    e2e_value, h2d, d2h, last = w.e2e(Ke, sampler)
    clocks = sampler.stop() if sampler else None
    floor_us = h.launch_floor_us() if rank == 0 else None
    roof = roofline_of(w, ms / K, pk, floor_us)

    # ---------------- CPU baseline (rank 0, N=1 only) ---------------------------------
    cpu = None
    dims, B, replicas, prec = w.dims, w.B, w.replicas, w.prec
    if rank == 0 and world == 1 and not args.no_cpu:
        reps = 5 if dims == synth.WIDE else 20
        rows = 20 if replicas else (args.cpu_rows or cpu_rows_for_budget(dims, B, reps + 1, 1, 15.0))
        rate, sec = cpu_learn_rate(dims, rows, reps, 1, 1)
        cpu = {"value": rate, "unit": "transitions/s", "cores": 1, "kind": "port",
               "sample": f"{reps} learn steps of {rows} rows of the same workload, single thread (the Gorgonia VM is "
                         f"single-threaded); oracle port, the Go reference cannot run here"}
    path = {1: "generic", 2: "narrow", 3: "wide"}[h.path]
    mode = {0: None, 1: "ncclAllReduce + update kernel",
            2: "gradient exchange over NVLink peer memory (CUDA IPC) inside the update kernel"}[h.dp_mode()]
    w.close()

    # ---------------- the other single-GPU BASELINE configs, on the same record -----------
    extra = None
    if rank == 0 and world == 1 and args.workload == "wide" and not args.no_extra:
        extra = {}
        for name, key in (("cartpole4096", "configs[1]"), ("replicas", "configs[4] (one GPU's share: 128 of 1024 replicas)")):
            x = Workload(G, args, name, 0, 1, local_rank, dist, torch)
            Kx = max(K, 200)                      # microsecond steps: 200 of them are still a few ms
            v, msx, lx = x.resident(Kx, max(W, 3))
            ev, h2dx, d2hx, _ = x.e2e(max(3, min(Kx, args.e2e_steps)))
            rf = roofline_of(x, msx / Kx, pk, floor_us)
            extra[name] = {"baseline_config": key, "value": v, "unit": "transitions/s", "us_per_step": msx / Kx * 1e3,
                           "steps": Kx, "gpu_launches": int(lx), "e2e": {"value": ev, "h2d_bytes_per_step": h2dx,
                                                                          "d2h_bytes_per_step": d2hx},
                           "config": workload_config(args, x.dims, x.B, x.replicas, x.prec, name=name), "roofline": rf}
            x.close()
    elif world > 1 and args.workload == "wide" and not args.no_extra:
        # configs[4]: independent CartPole replicas sharded over the GPUs, 128 per GPU (1024 on 8 B200), no
        # collective on the data path: every rank trains its own replicas, the timing is the max over ranks
        x = Workload(G, args, "replicas", rank, world, local_rank, dist, torch)
        Kx = max(K, 200)
        v, msx, lx = x.resident(Kx, max(W, 3))
        x.close()
        if rank == 0:
            extra = {"replicas": {"baseline_config": f"configs[4]: {128 * world} independent replicas over {world} GPUs (128 per GPU)",
                                  "value": v, "unit": "transitions/s", "us_per_step": msx / Kx * 1e3, "steps": Kx,
                                  "gpu_launches": int(lx), "n_gpus": world, "replicas_total": 128 * world,
                                  "scaling": "weak", "collective": None,
                                  "config": workload_config(args, x.dims, x.B, x.replicas, x.prec, name="replicas")}}

    if rank == 0:
        line = {
            "metric": "dqn_transitions_trained_per_sec", "value": value, "unit": "transitions/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": scaling_of(args), "vs_baseline": None, "dtype": "bf16" if prec == "bf16" else "f32",
            "data": "synthetic", "config": workload_config(args, dims, B, replicas, prec),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke,
                    "what": "per step: one fresh batch from ONE pinned host buffer (goldq_replay_stage_packed: H2D on the "
                            "copy stream, started two steps ahead; goldq_replay_commit behind the step) + goldq_learn "
                            "with host indices + goldq_last_loss readback (sync); host wall clock, max over ranks"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "path": path, "last_loss": [float(x) for x in last[:4]], "allreduce": mode,
            "configs": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None

# ff_forward2_kernel<2, 8>, wide config, B=8192: dram__bytes_read.sum 5.812736 MB + dram__bytes_write.sum 0 per launch
# (one `ncu --set full` capture of the end-of-round build; its activations and mask words stay in the 126 MB L2)
NCU_TRAFFIC = {("fused forward L1-L3 online+target, CTA pairs (tcgen05 cta_group::2)", (128, 512, 512, 18), 8192):
               5812736}
NCU_TRAFFIC_SOURCE = "profiles/r02_step_kernels_ncu_full_raw.csv"


def emit(line):
    """The ONE JSON line goes to the real stdout; everything libraries print meanwhile (NCCL's
    version banner, torchrun chatter) has been diverted to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="wide", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--replay-rows", type=int, default=0)
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--chunk", type=int, default=64, help="learn steps per goldq_learn_many call")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="N > 1, wide workload: strong (default) = --global-batch rows over the ranks; weak = 8192 rows per GPU")
    ap.add_argument("--global-batch", type=int, default=65536, help="configs[3]: global batch of the data-parallel step")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[1] / configs[4] workloads on the N=1 line")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample nvidia-smi during the timed region")
    ap.add_argument("--cpu-budget", type=float, default=60.0, help="seconds of CPU work for --impl reference")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
