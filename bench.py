#!/usr/bin/env python
"""Benchmark of the hot path: batched inside + outside (expected rule counts) chart pass.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one JSON line on rank 0)
    python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host cores (CPU)

Workload (BASELINE.json configs[2], the configuration the metric's target is quoted on): PCFG with N = 256
nonterminals, V = 32 terminals, sequences of length 128, batch 512 per GPU, inside + outside.  A "step" is one
pass of the hot path over one batch: inside pass (log Z for every sequence) followed by the outside pass
(gradients of sum log Z = expected rule counts) and, for N > 1 GPUs, the NCCL all-reduce of the count tensors.
Weak scaling: every rank owns its own batch of 512 sequences; no collective touches chart data.

Metric: inside+outside sequences / second (whole job).  `value` times the device-resident path (inputs in HBM),
`e2e` the public API (`SequentialRBN.log_inside_batch` on a model whose Parameters live on the GPU) with the token
batch in pinned host memory and the loss read back to the host every step.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, V, L, B per GPU, with outside pass)
    # batch 512 per GPU, run as micro-batches of 128: the split-sums of 128 sequences (136 GB of fp16 N x N blocks) fit
    # HBM beside the charts, so the outside pass reads them back instead of recomputing them
    "c3": dict(N=256, V=32, L=128, B=512, micro=128, outside=True,
               name="PCFG N=256 L=128 batch 512/GPU inside+outside (BASELINE configs[2])"),
    "c2": dict(N=64, V=32, L=64, B=1024, outside=False,
               name="discrete RBN N=64 L=64 batch 1024/GPU inside only (BASELINE configs[1])"),
    # configs[4]: global batch 8192 over 8 GPUs = 1024 sequences per GPU, run as micro-batches of 512 (chart + outside
    # chart of 512 sequences are 88 GB at N = 512); counts accumulate over the micro-batches, one all-reduce per step
    "c5": dict(N=512, V=32, L=128, B=1024, micro=512, outside=True,
               name="PCFG N=512 L=128 batch 1024/GPU (8192 over 8 GPUs) inside+outside, EM E-step (BASELINE configs[4])"),
    # configs[0]: the reference's own CPU-runnable case (plot_PCFG.py-style): latency-bound, one sequence
    "c1": dict(N=10, V=19, L=20, B=1, outside=True,
               name="discrete PCFG N=10 V=19, single length-20 sequence, inside+outside (BASELINE configs[0]; latency-bound)"),
    # configs[3]: continuous RBN, multivariate-normal transitions (separate code path: run_gauss)
    "c4": dict(D=8, L=64, B=256, outside=True, gauss=True,
               name="Gaussian RBN D=8 L=64 batch 256/GPU, inside marginal likelihood + gradient (BASELINE configs[3])"),
    # the same with float32 charts (the reference's default dtype; csrc/gauss_kernels.cu, RBN_GAUSS_F32)
    "c4f32": dict(D=8, L=64, B=256, outside=True, gauss=True, f32=True,
                  name="Gaussian RBN D=8 L=64 batch 256/GPU, inside marginal likelihood + gradient, float32 charts (BASELINE configs[3] at the reference's default dtype)"),
    "tiny": dict(N=64, V=32, L=16, B=8, outside=True, name="tiny self-test"),
}


def flops_per_sequence(N, L, outside):
    """Algorithmic FLOPs: split-sum 2 N^2 (L^3-L)/6, rule contraction N^3 L (L-1);
    inside+outside = 3x inside."""
    f_split = 2.0 * N * N * (L ** 3 - L) / 6.0
    f_rule = float(N) ** 3 * L * (L - 1)
    mult = 3.0 if outside else 1.0
    return mult * (f_split + f_rule), mult * f_rule, f_rule


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p.get("bf16_tflops_sustained", 1394.7), burst=p.get("bf16_tflops", 1670.0),
                    hbm=p.get("hbm_gbs", 6553.0), source="measured")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------------
CPU_SAMPLE_OUT = {}        # inputs / outputs of the last cpu_sample() run: the bench's parity check reuses them


def _err_metrics(d, ref, floor=1e-4):
    import numpy as np
    d, ref = np.asarray(d, dtype=np.float64).ravel(), np.asarray(ref, dtype=np.float64).ravel()
    mx = max(float(np.abs(ref).max()), 1e-300)
    big = np.abs(ref) >= floor * mx
    return (float(np.abs(d - ref).max() / mx), float(np.linalg.norm(d - ref) / max(float(np.linalg.norm(ref)), 1e-300)),
            float((np.abs(d - ref)[big] / np.abs(ref)[big]).max()))


def parity_check(wl, batch_logz0=None):
    """The CUDA path on exactly the sample the cpu_baseline leg just ran through the oracle (sequence 0 of the benched
    batch): log Z and, entry by entry, the gradients (max-norm, relative L2, per-entry relative above 1e-4 of the max)."""
    import numpy as np
    import torch
    from rbn_b200 import inside_logZ
    if not CPU_SAMPLE_OUT:
        return None
    W, T, pr, tok, le = CPU_SAMPLE_OUT["inputs"]
    out = CPU_SAMPLE_OUT["outputs"]
    Wt, Tt, pt = (torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=wl["outside"]) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.as_tensor(tok), torch.as_tensor(le))
    res = dict(sample="sequence 0 of the benched batch through the CUDA path (batch of 1) vs the fp64 oracle run of cpu_baseline")
    if wl["outside"]:
        lz.sum().backward()
        lz_ref, gW, gT, gP = (x.numpy() for x in out)
        errs = [_err_metrics(a.grad.cpu().numpy(), b) for a, b in ((Wt, gW), (Tt, gT), (pt, gP))]
        res.update(grad_max_rel=max(e[0] for e in errs), grad_l2_rel=max(e[1] for e in errs),
                   grad_entry_rel=max(e[2] for e in errs), entry_floor="entries >= 1e-4 * max|ref|")
    else:
        lz_ref = out.detach().numpy()
    lzv = lz.detach().double().cpu().numpy()
    res["logZ_rel"] = float(np.max(np.abs(lzv - lz_ref) / np.abs(lz_ref)))
    if batch_logz0 is not None and tok.shape[1] == wl["L"]:
        res["logZ_rel_in_batch"] = float(abs(batch_logz0 - lz_ref[0]) / abs(lz_ref[0]))   # same sequence inside the benched batch
    res["tolerance"] = "logZ 1e-4, gradients 1e-3 (north_star)"
    res["ok"] = bool(res["logZ_rel"] <= 1e-4 and res.get("grad_max_rel", 0) <= 1e-3 and res.get("grad_l2_rel", 0) <= 1e-3)
    return res


def workload_config(wl, B, world):
    """The `config` object of the JSON line: the workload only, identical for this repo's arm and the reference arm."""
    if wl.get("gauss"):
        return dict(workload=wl["name"], D=wl["D"], L=wl["L"], batch_per_gpu=B, global_batch=world * B, parallelism=f"dp{world}")
    return dict(workload=wl["name"], N=wl["N"], V=wl["V"], L=wl["L"], batch_per_gpu=B, global_batch=world * B,
                parallelism=f"dp{world}")


def cpu_sample(wl, seconds_budget=25.0):
    """Time the fp64 oracle (oracle/inside_oracle.py, a restatement of rbnet's inside pass + autograd outside) on a
    bounded sample of the workload: whole sequences of the workload's N, with the length truncated so that one
    sequence fits the time budget; the per-sequence time at full length is extrapolated with the algorithmic FLOP
    ratio and stated as such."""
    import torch
    from oracle import inside_oracle as O
    N, V, L = wl["N"], wl["V"], wl["L"]
    W, T, pr = O.synthetic_grammar(N, V, seed=0)
    # pick the sample length from a quick probe at L/4
    Lp = max(4, L // 4)
    tok, le = O.synthetic_tokens(1, Lp, V, seed=1)
    t0 = time.perf_counter()
    run = (lambda tk, ln: O.flat_inside_with_grads(W, T, pr, tk[:, :, None], ln)) if wl["outside"] else \
          (lambda tk, ln: O.flat_inside(torch.as_tensor(W), torch.as_tensor(T), torch.as_tensor(pr), torch.as_tensor(tk)[:, :, None], ln))
    out = run(tok, le)
    t_probe = time.perf_counter() - t0
    f_probe = flops_per_sequence(N, Lp, wl["outside"])[0]
    Ls = Lp
    for cand in (L, (3 * L) // 4, L // 2, (3 * L) // 8):
        if t_probe * flops_per_sequence(N, cand, wl["outside"])[0] / f_probe <= seconds_budget:
            Ls = cand
            break
    if Ls != Lp:
        tok, le = O.synthetic_tokens(1, Ls, V, seed=1)
        t0 = time.perf_counter()
        out = run(tok, le)
        t_s = time.perf_counter() - t0
    else:
        t_s = t_probe
    scale = flops_per_sequence(N, L, wl["outside"])[0] / flops_per_sequence(N, Ls, wl["outside"])[0]
    t_full = t_s * scale
    sample = (f"1 sequence (sequence 0 of the benched batch" + ("" if Ls == L else f", first {Ls} of {L} tokens; time scaled x{scale:.2f} by algorithmic FLOPs")
              + f"), N={N}, L={Ls}, fp64 torch restatement of rbnet inside{'+autograd outside' if wl['outside'] else ''}, "
              f"{t_s:.2f}s measured")
    CPU_SAMPLE_OUT.clear()
    CPU_SAMPLE_OUT.update(inputs=(W, T, pr, tok, le), outputs=out)
    return 1.0 / t_full, t_full, sample, torch.get_num_threads()


def run_reference_gauss(args, wl):
    """--impl reference for configs[3]: the fp64 torch composition of rbnet.multivariate_normal's algebra (oracle/
    gauss_oracle.py) with autograd as the outside pass, whole sequences of the workload, on the host cores."""
    import torch
    from oracle import gauss_oracle as GO
    torch.set_num_threads(os.cpu_count() or 1)
    D, L = wl["D"], wl["L"]
    syn = GO.synthetic_gauss(D, 4, L, seed=10)
    dt = torch.float32 if wl.get("f32") else torch.float64          # c4f32: the reference's default dtype
    tt = {k: torch.tensor(v, dtype=dt, requires_grad=(k != "y")) for k, v in syn.items()}
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        lz = GO.gauss_inside(tt["y"][s % 4:s % 4 + 1], [L], tt["cov_term"], tt["cov_left"], tt["cov_right"], tt["prior_mean"],
                             tt["prior_cov"], tt["log_struct"])
        lz.sum().backward()
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    v = 1.0 / t
    cores = torch.get_num_threads()
    sample = f"1 sequence per step, D={D}, L={L}, {'fp32' if wl.get('f32') else 'fp64'} torch composition of rbnet.multivariate_normal + autograd"
    print(json.dumps(dict(metric="inside+gradient sequences/sec", value=v, unit="seq/s", n_gpus=args.gpus, steps=args.steps,
                          warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="f32" if wl.get("f32") else "f64", data="synthetic", impl="reference",
                          config=workload_config(wl, args.batch or wl["B"], int(os.environ.get("WORLD_SIZE", "1"))),
                          cpu_baseline=dict(value=v, unit="seq/s", cores=cores, kind="port", sample=sample),
                          e2e=dict(value=v, unit="seq/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)), flush=True)


def literal_reference_available(wl):
    """The UNMODIFIED reference (rbnet's own Python loop nest + torch.autograd) can be timed where it is importable
    (/root/reference, or the copy oracle/vendor_reference.py makes at build time) and where one pass finishes in seconds:
    configs[0].  At C3 it needs ~4 h per sequence: there the vectorised port is the CPU arm."""
    if wl.get("gauss") or wl["N"] > 16 or wl["L"] > 24:
        return False
    try:
        from oracle.reference_loader import reference_available
        return reference_available()
    except Exception:
        return False


def literal_reference_sample(wl, reps, threads):
    """seq/s of rbnet.SequentialRBN.inside() + backward() (tier T0) on sequence 0 of the workload, `threads` host threads."""
    import numpy as np
    import torch
    from oracle.reference_loader import load_reference
    from rbn_b200 import synth
    ref = load_reference()
    torch.set_num_threads(threads)
    N, V, L = wl["N"], wl["V"], wl["L"]
    rng = np.random.default_rng(0)          # the same weights build_model() gives this repo's model
    rp = ref["pcfg"]
    cell = rp.DiscreteCell(variable=rp.DiscreteNonTermVar(N), weights=rng.random((2, N)),
                           transitions=[rp.DiscreteTerminalTransition(weights=rng.random((V, N))),
                                        rp.DiscreteBinaryNonTerminalTransition(weights=rng.random((N, N, N)))])
    prior = rp.DiscretePrior(struc_weights=np.ones(1), prior_weights=[rng.random(N)])
    model = ref["sequential"].SequentialRBN(cells=[cell], prior=prior)
    tok, _ = synth.synthetic_tokens(1, L, V, seed=1)
    seq = [[int(y)] for y in tok[0]]
    times, z = [], None
    for _ in range(reps):
        model.zero_grad()
        t0 = time.perf_counter()
        Z = model.inside(sequence=seq)
        (-torch.log(Z)).backward()
        times.append(time.perf_counter() - t0)
        z = float(Z.detach())
    return 1.0 / (sum(times) / len(times)), z


def run_reference_literal(args, wl):
    import torch
    cores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        literal_reference_sample(wl, 1, cores)
    v_all, z = literal_reference_sample(wl, args.steps, cores)
    v_one, _ = literal_reference_sample(wl, args.steps, 1)
    sample = (f"rbnet (unmodified, {('/root/reference' if os.path.isdir('/root/reference/rbnet') else 'oracle/_ref copy')}) "
              f"SequentialRBN.inside() + backward(), 1 sequence per step, N={wl['N']} L={wl['L']}; "
              f"{v_all:.2f} seq/s on {cores} threads, {v_one:.2f} seq/s on 1 thread (the loop nest is Python-bound)")
    line = dict(metric="inside+outside sequences/sec", value=v_all, unit="seq/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / v_all, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=workload_config(wl, args.batch or wl["B"], int(os.environ.get("WORLD_SIZE", "1"))),
                cpu_baseline=dict(value=v_all, unit="seq/s", cores=cores, kind="reference", sample=sample, value_1thread=v_one),
                e2e=dict(value=v_all, unit="seq/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0, Z=z)
    print(json.dumps(line), flush=True)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if wl.get("gauss"):
        return run_reference_gauss(args, wl)
    if literal_reference_available(wl):
        return run_reference_literal(args, wl)
    import torch
    torch.set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1: use the box's cores anyway
    vals = []
    sample, cores = "", torch.get_num_threads()
    for s in range(args.warmup + args.steps):
        v, t_full, sample, cores = cpu_sample(wl, seconds_budget=20.0)
        if s >= args.warmup:
            vals.append((v, t_full))
    v = sum(x[0] for x in vals) / len(vals)
    t = sum(x[1] for x in vals) / len(vals)
    line = dict(metric="inside+outside sequences/sec" if wl["outside"] else "inside sequences/sec", value=v,
                unit="seq/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=workload_config(wl, args.batch or wl["B"], int(os.environ.get("WORLD_SIZE", "1"))),
                cpu_baseline=dict(value=v, unit="seq/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit="seq/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="the literal reference (pure Python loop nest) needs hours per sequence at this size (BASELINE.md section 2; it "
                     "IS the timed arm for --workload c1); this arm times oracle/, its vectorised fp64 restatement, on all host "
                     "cores, one sequence per step")
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------------
def build_model(N, V, seed, device):
    """A single-cell discrete RBN of the AbstractedPCFG shape with random dense weights (the public API object)."""
    import numpy as np
    from rbn_b200 import pcfg, sequential
    rng = np.random.default_rng(seed)
    cell = pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(N), weights=rng.random((2, N)),
                             transitions=[pcfg.DiscreteTerminalTransition(weights=rng.random((V, N))),
                                          pcfg.DiscreteBinaryNonTerminalTransition(weights=rng.random((N, N, N)))])
    prior = pcfg.DiscretePrior(struc_weights=np.ones(1), prior_weights=[rng.random(N)])
    return sequential.SequentialRBN(cells=[cell], prior=prior).to(device)


def run_ours(args, wl, secondary=False):
    """This repo's arm.  `secondary`: called a second time inside a multi-GPU run for the C5 line (bounded steps; the
    process group already exists); returns the line instead of printing it."""
    import gc
    import numpy as np
    import torch
    import torch.distributed as dist
    from rbn_b200 import _lib, inside_logZ, synth
    from rbn_b200.distributed import CountReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    if secondary:       # bounded: a C5 step takes ~10 s; the small workloads take milliseconds (3 warm-ups, 5 steps)
        steps, warmup = (min(args.steps, 2), min(args.warmup, 1)) if wl["N"] >= 512 else (5, 3)
    else:
        steps, warmup = args.steps, args.warmup

    N, V, L, B = wl["N"], wl["V"], wl["L"], (wl["B"] if secondary else (args.batch or wl["B"]))
    W, T, pr = synth.synthetic_grammar(N, V, seed=0)
    tok_np, len_np = synth.synthetic_tokens(B, L, V, seed=1 + rank)
    Wd = torch.tensor(W, dtype=torch.float32, device=dev, requires_grad=wl["outside"])
    Td = torch.tensor(T, dtype=torch.float32, device=dev, requires_grad=wl["outside"])
    pd = torch.tensor(pr, dtype=torch.float32, device=dev, requires_grad=wl["outside"])
    tok_d = torch.tensor(tok_np, dtype=torch.int32, device=dev)
    len_d = torch.tensor(len_np, dtype=torch.int32, device=dev)
    len_h = torch.tensor(len_np, dtype=torch.int32)           # host copy of the lengths: the pass then never synchronises

    micro = min(B, args.micro or wl.get("micro", B))

    def device_step():
        """One pass of the hot path over the batch, inputs resident in HBM.  Micro-batches run inside + outside back to
        back (the second pass reuses the first one's workspace); every micro-batch's count tensors go into an
        asynchronous all-reduce that overlaps the next micro-batch (N > 1), and are summed at the end of the step."""
        if args.streams > 1:
            return device_step_streams()
        red = CountReducer()
        outs = []
        for b0 in range(0, B, micro):
            lz = inside_logZ(Wd, Td, pd, tok_d[b0:b0 + micro], len_h[b0:b0 + micro], chunk_items=args.chunk)
            if wl["outside"]:
                red.add(torch.autograd.grad(lz.sum(), (Wd, Td, pd)))
            outs.append(lz.detach())
            del lz                           # drops the pass (its workspace goes back to the allocator for the next one)
        grads = red.result() if wl["outside"] else None
        return (outs[0] if len(outs) == 1 else torch.cat(outs)), grads

    side = [torch.cuda.Stream(device=dev) for _ in range(max(0, args.streams))] if args.streams > 1 else []

    def device_step_streams():
        """Experiment (--streams S, single GPU): the micro-batches dealt over S launch streams, so that the ramp-up and the
        last wave of one micro-batch's per-diagonal launches overlap another's."""
        cur = torch.cuda.current_stream()
        outs, gsum = {}, [None] * len(side)
        for st in side:
            st.wait_stream(cur)
        for k, b0 in enumerate(range(0, B, micro)):
            st = side[k % len(side)]
            with torch.cuda.stream(st):
                lz = inside_logZ(Wd, Td, pd, tok_d[b0:b0 + micro], len_h[b0:b0 + micro], chunk_items=args.chunk)
                if wl["outside"]:
                    g = torch.autograd.grad(lz.sum(), (Wd, Td, pd))
                    j = k % len(side)
                    gsum[j] = list(g) if gsum[j] is None else [a.add_(b) for a, b in zip(gsum[j], g)]
                outs[k] = lz.detach()
                del lz
        for st in side:
            cur.wait_stream(st)
        grads = None
        if wl["outside"]:
            parts = [g for g in gsum if g is not None]
            grads = tuple(sum(p[i] for p in parts[1:]) + parts[0][i] if len(parts) > 1 else parts[0][i] for i in range(3))
        o = [outs[k] for k in sorted(outs)]
        return (o[0] if len(o) == 1 else torch.cat(o)), grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps, profile_last=False):
        """nsteps calls of fn between two CUDA events (barrier + synchronize on both sides), max over ranks.  With
        `profile_last` the LAST of the timed steps also brackets every kernel launch of the library with CUDA events on
        the launching stream (rbn_profile_*): the per-kernel times come from inside the timed region, and the event
        records (two per launch, ~3600 launches per step) tax one step instead of all of them."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for k in range(nsteps):
            if profile_last and k == nsteps - 1:
                _lib.profile_enable(True)
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / nsteps, out

    # ---- device-resident arm -------------------------------------------------------------------------------------------
    for _ in range(warmup):
        device_step()
    _lib.profile_read(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    prof_steps = steps if args.profile_all else 1
    if args.profile_all:
        _lib.profile_enable(True)
    ms_step, (lz, grads) = timed(device_step, steps, profile_last=not args.profile_all)
    clocks = sampler.stop() if rank == 0 else None
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    lz_host = lz.detach().double().cpu().numpy()
    # sequence 0 of the benched batch is sequence 0 of the headline fixture (tests/golden/headline_c3.npz, written by the
    # fp64 oracle): its log Z inside the full batch, checked on every run without the 30 s oracle
    fixture_check = None
    fpath = os.path.join(ROOT, "tests", "golden", "headline_c3.npz")
    if (N, V, L) == (256, 32, 128) and rank == 0 and os.path.exists(fpath):
        ref0 = float(np.load(fpath)["logZ"][0])
        fixture_check = dict(logZ0=float(lz_host[0]), logZ0_oracle=ref0, abs_err=float(lz_host[0] - ref0),
                             rel_err=float(abs(lz_host[0] - ref0) / abs(ref0)))
    count_check = None
    if grads is not None:
        # expected-count identity over the whole job: sum_{bca} gW * W = number of binary rule applications = sum (len - 1)
        count_check = dict(binary_rules=float((grads[0].double() * Wd.detach().double()).sum()),
                           expected=float(world * (len_np - 1).sum()))
    del lz, grads
    from rbn_b200 import release_workspaces
    release_workspaces()
    gc.collect()
    torch.cuda.empty_cache()

    # ---- end-to-end arm: public API, host tokens in pinned memory, loss back to the host --------------------------------
    model = build_model(N, V, seed=0, device=dev)
    tok_pin = torch.tensor(tok_np, dtype=torch.int32).pin_memory()
    len_pin = torch.tensor(len_np, dtype=torch.int32).pin_memory()
    params = [p for p in model.parameters()]

    def e2e_step():
        """The call a user makes: SequentialRBN.log_inside_batch on a pinned host batch (the library copies it to the device,
        asynchronously), loss.backward() through the Parameters, gradients all-reduced, the loss read back to the host."""
        if wl["outside"]:
            for p in params:
                p.grad = None
        total = None
        for b0 in range(0, B, micro):
            lz = model.log_inside_batch(tokens=tok_pin[b0:b0 + micro], lengths=len_pin[b0:b0 + micro])
            loss = -lz.sum()
            if wl["outside"]:
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
            del lz, loss
        if wl["outside"] and world > 1:
            from rbn_b200.distributed import allreduce_counts
            allreduce_counts([p.grad for p in params])
        return float(total)         # device -> host read of the step's result

    if args.skip_e2e:
        ms_e2e = float('nan')
    else:
        for _ in range(min(warmup, 2)):
            e2e_step()
        ms_e2e, _ = timed(e2e_step, steps)
    h2d = tok_pin.numel() * 4 + len_pin.numel() * 4
    d2h = 8
    del model, params
    release_workspaces()
    gc.collect()
    torch.cuda.empty_cache()

    if rank != 0:
        return None

    f_step, f_rule_step, f_rule_fwd = (x * B for x in flops_per_sequence(N, L, wl["outside"]))
    peaks = load_peaks()
    kms = {k: v[0] / prof_steps for k, v in prof.items() if v[1]}     # event times: summed over the profiled step(s)
    kms["rule_gemm"] = kms.get("rule_gemm", 0.0) + kms.pop("rule_finalize", 0.0)   # the max-shift epilogue of K-split tiles
    roofline = dict(step_tflops=f_step / (ms_step * 1e-3) / 1e12, step_frac=f_step / (ms_step * 1e-3) / 1e12 / peaks["tflops"],
                    step_peak=f"{peaks['source']} bf16_tflops_sustained {peaks['tflops']} TFLOP/s (kernels timed inside a long step)",
                    algorithmic_flops_per_step=f_step,
                    kernels_ms_per_step={k: round(v, 3) for k, v in kms.items()})
    # Every kernel class of the tensor-core path against the roofline that bounds it: algorithmic
    # bytes or flops per step / summed CUDA-event time of the class.  The top-level entry is the class with the most time.
    if N >= 64 and prof["split_sum"][1]:
        spans = B * (L * (L - 1) // 2)                         # spans of width >= 2 per pass
        splits = B * ((L ** 3 - L) // 6)
        units = max(1, prof["rule_gemm"][1])                   # (diagonal, chunk) units of the inside passes
        k1_passes = prof["split_sum"][1] / units               # 1 = every split-sum kept for the outside pass, 2 = all recomputed
        y_bytes = 2.0 * N * N * spans
        row_bytes = 2.0 * splits * 2 * N                       # fp16 child rows streamed once per diagonal
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"n{N}", {})
        names = dict(split_sum="k_tc_splitsum (split-sum Y = sum_j c_j L_j^T R_j, tcgen05 + TMA store)",
                     rule_gemm="k_tc_gemm2<rule> (rule contraction, tcgen05 cta_group::2)",
                     gy_gemm="k_tc_gy2 (gY = G Wk^T, tcgen05 cta_group::2, TMA store)",
                     count_gemm="k_tc_gemm2<counts> (expected rule counts gW += Y^T G, tcgen05 cta_group::2)",
                     child_grad="k_tc_childgrad (child gradients, tcgen05 + TMA reduce-add into the outside chart)")
        per = {}

        def entry(cls, bound, amount, peak, unit):
            ms = kms.get(cls, 0.0)
            if ms <= 0:
                return
            launches = prof[cls][1] / steps                    # (launches are counted on every step)
            ach = amount / (ms * 1e-3) / (1e12 if unit == "TFLOP/s" else 1e9)
            tr = traffic.get(cls, {}).get("dram_bytes_per_span")
            per[cls] = dict(bound=bound, kernel=names[cls], ms_per_step=round(ms, 3), achieved=round(ach, 1), peak=peak, unit=unit,
                            frac=round(ach / peak, 3), launches_per_step=launches, avg_launch_ms=ms / launches,
                            algorithmic_per_launch=amount / launches,
                            traffic=(tr * spans * (k1_passes if cls == "split_sum" else 1.0) / launches) if tr else None)
        # the rule / count GEMMs do 2 N^3 flops per 2 N^2 bytes of Y: N flop per byte, against a ridge point of
        # peak flops / peak bytes (~213 flop/B on the measured peaks): HBM-bound for N <= 192, tensor-bound from N = 256
        gemm_hbm = N < peaks["tflops"] * 1e3 / peaks["hbm"]
        entry("split_sum", "hbm", k1_passes * (y_bytes + row_bytes), peaks["hbm"], "GB/s")
        if gemm_hbm:
            entry("rule_gemm", "hbm", y_bytes, peaks["hbm"], "GB/s")
        else:
            entry("rule_gemm", "tensor", f_rule_fwd, peaks["tflops"], "TFLOP/s")
        if wl["outside"]:
            # gY is write-only output: N flop per byte WRITTEN, and pure writes reach ~0.6 of the copy bandwidth on this part
            # (memset 3.9 TB/s against 6.5): HBM-bound up to N ~ 355, tensor-bound at N = 512
            if N < peaks["tflops"] * 1e3 / (0.6 * peaks["hbm"]):
                entry("gy_gemm", "hbm", y_bytes, peaks["hbm"], "GB/s")
            else:
                entry("gy_gemm", "tensor", f_rule_fwd, peaks["tflops"], "TFLOP/s")
            if gemm_hbm:
                entry("count_gemm", "hbm", y_bytes, peaks["hbm"], "GB/s")
            else:
                entry("count_gemm", "tensor", f_rule_fwd, peaks["tflops"], "TFLOP/s")
            entry("child_grad", "hbm", y_bytes + row_bytes + 2.0 * splits * 2 * N * 4, peaks["hbm"], "GB/s")
        if per:
            top = max(per, key=lambda k: per[k]["ms_per_step"])
            t = per[top]
            roofline.update(bound=t["bound"], kernel=t["kernel"], achieved=t["achieved"], peak=t["peak"], unit=t["unit"],
                            frac=t["frac"], traffic=t["traffic"],
                            traffic_source=("profiles/r02_traffic.json (ncu --set full, one launch, scaled per span)" if t["traffic"] else None),
                            launches=int(t["launches_per_step"] * steps), avg_launch_ms=t["avg_launch_ms"],
                            algorithmic_per_launch=t["algorithmic_per_launch"],
                            peak_source=f"{peaks['source']} " + ("hbm_gbs" if t["bound"] == "hbm" else "bf16_tflops_sustained"),
                            selected="the kernel class with the largest share of the step")
        roofline["by_kernel"] = per
        roofline["split_sum_passes"] = round(k1_passes, 3)
        # the step as a whole against HBM: Y written by every split-sum pass and read by the rule / count GEMM, gY
        # written and read once, child rows streamed by K1 and K5, outside-chart reduce-adds counted as read + write
        step_bytes = k1_passes * (y_bytes + row_bytes) + y_bytes
        if wl["outside"]:
            step_bytes += y_bytes + 2 * y_bytes + row_bytes + 2 * (2.0 * splits * 2 * N * 4)
        ach = step_bytes / (ms_step * 1e-3) / 1e9
        roofline["step_hbm"] = dict(algorithmic_bytes_per_step=step_bytes, achieved=round(ach, 1), peak=peaks["hbm"], unit="GB/s",
                                    frac=round(ach / peaks["hbm"], 3),
                                    note="every kernel of this path streams the N^2-per-span split-sum intermediate through HBM; "
                                         "the step is bandwidth-bound, see DESIGN.md section 3.4")
    launches = int(sum(v[1] for v in prof.values()))
    cpu_v = parity = None
    cpu_kind = "port"
    if world == 1 and not args.skip_cpu and not secondary:
        cpu_v, cpu_t, cpu_samp, cores = cpu_sample(wl)
        parity = parity_check(wl, batch_logz0=float(lz_host[0]))
        if literal_reference_available(wl):      # configs[0]: the literal reference is runnable: it is the CPU baseline
            port_v = cpu_v
            cores = os.cpu_count() or 1
            cpu_v, _ = literal_reference_sample(wl, 5, cores)
            one, _ = literal_reference_sample(wl, 5, 1)
            cpu_kind = "reference"
            cpu_samp = (f"rbnet (unmodified) SequentialRBN.inside() + backward(), sequence 0, 5 repetitions: {cpu_v:.2f} seq/s on "
                        f"{cores} threads, {one:.2f} on 1 thread; the vectorised fp64 port (oracle/) does {port_v:.1f} seq/s")
    line = dict(metric="inside+outside sequences/sec" if wl["outside"] else "inside sequences/sec",
                value=world * B / (ms_step * 1e-3), unit="seq/s", n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16",
                data="synthetic", config=workload_config(wl, B, world),
                impl_notes=dict(micro_batch=micro, l2="inputs larger than L2 (chart + split-sum buffers are GBs)",
                                accumulate="fp32 (tcgen05, segmented)", operands="fp16 mantissas + per-cell exponents",
                                split_sum_cache="inside pass keeps Y in HBM for the outside pass where it fits (RbnDesc.cache_bytes)"),
                e2e=dict(value=world * B / (ms_e2e * 1e-3), unit="seq/s", ms_per_step=ms_e2e, h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=d2h),
                gpu_launches=launches, roofline=roofline, clocks=clocks,
                logZ_mean=float(np.mean(lz_host)), count_check=count_check, fixture_check=fixture_check)
    if cpu_v is not None:
        line["cpu_baseline"] = dict(value=cpu_v, unit="seq/s", cores=cores, kind=cpu_kind, sample=cpu_samp)
    if parity is not None:
        line["parity_check"] = parity
    return line


def _brief(line):
    """The part of a workload's line that goes under `other_workloads`."""
    r = line.get("roofline") or {}
    return dict(workload=line["config"]["workload"], metric=line["metric"], value=line["value"], unit=line["unit"],
                ms_per_step=line["ms_per_step"], steps=line["steps"], warmup=line["warmup"], e2e=line["e2e"]["value"],
                roofline=dict(bound=r.get("bound"), kernel=r.get("kernel"), frac=r.get("frac"), step_frac=r.get("step_frac")),
                kernels_ms_per_step=r.get("kernels_ms_per_step"), gpu_launches=line["gpu_launches"])


def finish_ours(args, wl):
    """Print this arm's ONE JSON line.  With more than one GPU the line also carries, under `secondary`, the C5 step
    (BASELINE configs[4]: N = 512, 1024 sequences per GPU, EM E-step + count all-reduce) measured in the same run; on one
    GPU, under `other_workloads`, short measurements of the other BASELINE configurations (configs[0], [1], [3])."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = run_ours(args, wl)
    if world == 1 and args.workload == "c3" and not args.no_secondary and not args.batch and line is not None:
        others = {}
        for name in ("c1", "c2", "c4", "c4f32"):          # never let a side measurement take the headline line down with it
            try:
                r = run_gauss(args, WORKLOADS[name], ret=True) if WORKLOADS[name].get("gauss") else run_ours(args, WORKLOADS[name], secondary=True)
                if r is not None:
                    others[name] = _brief(r)
            except Exception as e:              # noqa: BLE001
                others[name] = dict(error=f"{type(e).__name__}: {e}"[:300])
        line["other_workloads"] = others
    if world > 1 and args.workload == "c3" and not args.no_secondary:
        try:
            sec = run_ours(args, WORKLOADS["c5"], secondary=True)
        except Exception as e:                  # noqa: BLE001  (deterministic failures hit every rank alike: no rank is left waiting)
            sec = dict(error=f"{type(e).__name__}: {e}"[:300])
        if line is not None and sec is not None:
            line["secondary"] = sec
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


def run_gauss(args, wl, ret=False):
    """BASELINE configs[3]: Gaussian chart pass (csrc/gauss_kernels.cu), fp64 charts (584 B per cell; the parity-pinned
    mode) or, for wl["f32"], float32 charts (296 B per cell: SURVEY.md 8(d)'s 292 B at the reference's fp32 plus the fp64
    log weight).  HBM roofline per SURVEY.md 8(d): a pass re-reads two child cells per split, 2 (L^3 - L)/6 cells per
    sequence; inside + outside = forward reads + backward reads and adjoint read-modify-writes of the same cells (3x)."""
    import numpy as np
    import torch
    from rbn_b200 import synth
    from rbn_b200 import _lib
    from rbn_b200.gaussian import GaussianRBN
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    D, L, B = wl["D"], wl["L"], (wl["B"] if ret else (args.batch or wl["B"]))
    steps_g, warm_g = (5, 3) if ret else (args.steps, args.warmup)
    syn = synth.synthetic_gauss(D, B, L, seed=10 + rank)
    y_np = syn["y"]
    f32 = bool(wl.get("f32"))
    ydt = torch.float32 if f32 else torch.float64
    model = GaussianRBN(D, cov_term=syn["cov_term"], cov_left=syn["cov_left"], cov_right=syn["cov_right"],
                        prior_mean=syn["prior_mean"], prior_cov=syn["prior_cov"], struct_weights=(0.4, 0.6),
                        chart_dtype=ydt).to(dev)
    params = list(model.parameters())
    y_d = torch.tensor(y_np, dtype=ydt, device=dev)
    y_pin = torch.tensor(y_np, dtype=ydt).pin_memory()

    def step(y):
        for p in params:
            p.grad = None
        lz = model.log_inside_batch(y)
        (-lz.sum()).backward()
        if world > 1:
            from rbn_b200.distributed import allreduce_counts
            allreduce_counts([p.grad for p in params])
        return lz

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out

    for _ in range(warm_g):
        step(y_d)
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, lz = timed(lambda: step(y_d), steps_g)
    clocks = sampler.stop() if rank == 0 else None
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    ms_e2e, _ = timed(lambda: float(step(y_pin.to(dev, non_blocking=True)).sum()), steps_g)
    if rank != 0:
        return
    peaks = load_peaks()
    cell_bytes = 4 * (2 + D + D * D) if f32 else 8 * (1 + D + D * D)
    splits = (L ** 3 - L) // 6
    alg_bytes = 3.0 * B * 2 * splits * cell_bytes             # child-cell reads fwd + bwd + adjoint updates
    diag_ms = prof["simt_inside"][0] + prof["simt_outside"][0]      # k_gauss_diag + k_gauss_bwd_diag
    diag_launches = prof["simt_inside"][1] + prof["simt_outside"][1]
    ach = alg_bytes * steps_g / (diag_ms * 1e-3) / 1e9 if diag_ms > 0 else None
    cpu = None
    if world == 1 and not args.skip_cpu and not ret:
        from oracle import gauss_oracle as GO      # cpu_baseline leg only
        t0 = time.perf_counter()
        n_cpu = 0
        tt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=(k != "y")) for k, v in syn.items()}
        while time.perf_counter() - t0 < 15.0:
            lzc = GO.gauss_inside(torch.tensor(y_np[n_cpu % B:n_cpu % B + 1]), [L], tt["cov_term"], tt["cov_left"], tt["cov_right"],
                                  tt["prior_mean"], tt["prior_cov"], tt["log_struct"])
            lzc.sum().backward()
            n_cpu += 1
        dt = time.perf_counter() - t0
        cpu = dict(value=n_cpu / dt, unit="seq/s", cores=torch.get_num_threads(), kind="port",
                   sample=f"{n_cpu} sequences of the same workload, fp64 torch restatement on rbnet.multivariate_normal's algebra, {dt:.1f}s")
    line = dict(metric="inside+gradient sequences/sec", value=world * B / (ms_step * 1e-3), unit="seq/s", n_gpus=world,
                steps=steps_g, warmup=warm_g, ms_per_step=ms_step, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32" if f32 else "f64", data="synthetic",
                config=workload_config(wl, B, world),
                impl_notes=dict(l2=f"chart + adjoint chart ({2 * B * (L * (L + 1) // 2) * cell_bytes / 1e6:.0f} MB) are larger than L2",
                                log_weights="fp64 in both modes"),
                e2e=dict(value=world * B / (ms_e2e * 1e-3), unit="seq/s", ms_per_step=ms_e2e,
                         h2d_bytes_per_step=int(y_pin.numel() * y_pin.element_size()),
                         d2h_bytes_per_step=8),
                gpu_launches=int(sum(v[1] for v in prof.values())),
                roofline=dict(bound="hbm", kernel="k_gauss_diag / k_gauss_bwd_diag (per-diagonal cell pass)", achieved=ach,
                              peak=peaks["hbm"], unit="GB/s", frac=(ach / peaks["hbm"]) if ach else None, traffic=None,
                              peak_source=f"{peaks['source']} hbm_gbs", launches=int(diag_launches),
                              avg_launch_ms=(diag_ms / diag_launches) if diag_launches else None,
                              algorithmic_bytes_per_step=alg_bytes,
                              kernels_ms_per_step={k: round(v[0] / steps_g, 3) for k, v in prof.items() if v[1]}),
                clocks=clocks, logZ_mean=float(lz.detach().double().mean()))
    if cpu:
        line["cpu_baseline"] = cpu
    if ret:
        return line
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU (default: the workload's)")
    ap.add_argument("--chunk", type=int, default=0, help="spans per launch group (0 = library default)")
    ap.add_argument("--streams", type=int, default=1, help="experiment: deal the micro-batches over this many launch streams")
    ap.add_argument("--micro", type=int, default=0, help="sequences per micro-batch (default: the workload's)")
    ap.add_argument("--no-secondary", action="store_true", help="N > 1: skip the C5 line under `secondary`")
    ap.add_argument("--profile-all", action="store_true", help="bracket the kernels of EVERY timed step with events (default: the last one)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: skip the end-to-end arm")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: skip the CPU baseline")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif wl.get("gauss"):
        run_gauss(args, wl)
    else:
        finish_ours(args, wl)


if __name__ == "__main__":
    main()
