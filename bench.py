#!/usr/bin/env python
"""Headline benchmark: one G+D training step of td-vc-gan (config/conv_enc-stage1.yaml as shipped:
B=16/GPU, T=8960 samples @16 kHz, 100 speakers) on N B200s, in audio-seconds per second.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

Contract: W untimed warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the
launching stream, max over ranks; rank 0 prints ONE JSON line.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "td-vc-gan_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "train_audio_sec_per_sec"
UNIT = "audio-s/s"
SR = 16000
# model + train sections of config/conv_enc-stage1.yaml (the reference's file does not travel to the GPU box)
MODEL = dict(ratios=(10, 8, 2, 2), channels=(256, 128, 64, 32, 16), content_dim=128, cond_dim=128, nspk=100,
             num_disc=3, d_layers=4, d_base=16)
TRAIN = dict(no_conv=False, lambda_rec=0, lambda_idt=5, lambda_feat=2, lambda_spec=5, lambda_wave=0, lambda_latcls=0,
             lambda_cont_emb=10, lambda_corrupted=1, lambda_converted=0, lambda_f0=0, jitter_amp=0,
             batch_size=16, max_segment=8960, lr=1e-4, betas=(0.8, 0.99))
# BASELINE.json configs: train sections of the shipped YAMLs (lambda_f0 forced to 0: torchcrepe is unavailable
# offline) and the algorithmic FLOPs of one REFERENCE step at B=16 (analytic 2*MAC per conv,
# every pass the reference runs -- the passes this implementation shares or skips still count as work delivered)
CONFIGS = {
    "stage1": dict(train=TRAIN, gflop=3695.0,
                   workload="conv_enc-stage1.yaml as shipped: one D step + one G step (reference: G fwd x3, D fwd x5, "
                            "encoder(corrupted), LSGAN + feature-matching + mel + contrastive losses, both backward passes, "
                            "AdamW on G and D)"),
    "stage2_1": dict(train=dict(TRAIN, no_conv=True, lambda_idt=20), gflop=2299.0,
                     workload="conv_enc-stage2_1.yaml as shipped (no_conv: reconstruction; reference: G fwd x2, D fwd x4, "
                              "encoder(corrupted))"),
    "stage2_1_latcls": dict(train=dict(TRAIN, no_conv=True, lambda_idt=20, lambda_latcls=1), gflop=2299.0 + 15.0,
                            workload="conv_enc-stage2_1.yaml with lambda_latcls=1: latent classifier step + gradient "
                                     "reversal term (BASELINE config 2)"),
    "stage2_2": dict(train=dict(TRAIN, lambda_rec=10, lambda_idt=1), gflop=5159.0,
                     workload="conv_enc-stage2_2.yaml as shipped (conversion + reverse/cycle pass G(fake) -> rec, D(rec))"),
}
STEP_GFLOP_B16 = CONFIGS["stage1"]["gflop"]
WORKLOAD = CONFIGS["stage1"]["workload"]


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_models(device):
    import torch
    from model.discriminator import CollaborativeMultibandDiscriminator
    from model.generator import Generator
    torch.manual_seed(0)      # identical replicas on every rank
    m = MODEL
    G = Generator(list(m["ratios"]), list(m["channels"]), 0, m["nspk"], m["cond_dim"], m["content_dim"], 3, 0, "conv",
                  norm_layer=(None, None, None), weight_norm=("weight_norm",) * 3, bot_cond="target", enc_cond=None,
                  dec_cond="target", output_content_emb=True)
    D = CollaborativeMultibandDiscriminator(m["num_disc"], m["nspk"], m["d_layers"], m["d_base"], 4, 4, 128, "target")
    return G.to(device), D.to(device)


def synth_batch(B, T, nspk, seed):
    """SURVEY.md 8(d) synthetic batch (host tensors, fp32)."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float32) / SR
    f0 = torch.rand(B, 1, 1, generator=g) * 200 + 100
    x = 0.05 * torch.sin(2 * math.pi * f0 * t) + 0.01 * torch.randn(B, 1, T, generator=g)
    xc = x + 0.005 * torch.randn(B, 1, T, generator=g)
    lab_s = torch.randint(0, nspk, (B,), generator=g)
    perm = torch.randperm(B, generator=g)
    lab_t, f0_t = lab_s[perm], f0[perm]

    def excite(f):
        ph = torch.rand(1, generator=g) * 2 * math.pi
        return 0.1 * torch.sin(2 * math.pi * f * t + ph) + 0.003 * torch.randn(B, 1, T, generator=g)
    return {"signal_real": x, "signal_corrupted": xc, "label_src": lab_s, "label_tgt": lab_t,
            "c_f0_conv": excite(f0_t), "c_f0_src": excite(f0)}


def to_device(batch, device, pinned=None):
    import torch
    out = {}
    nbytes = 0
    for k, v in batch.items():
        src = pinned[k] if pinned is not None else v
        out[k] = src.to(device, non_blocking=True)
        nbytes += v.numel() * v.element_size()
    return out, nbytes


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tdvc import _lib, ops
    from tdvc.dp import GradAverager, broadcast_parameters
    from tdvc.optim import FusedAdamW
    from tdvc.train_step import GraphedTrainStep, TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    ops.set_precision(args.precision)
    ops.set_branch_streams(args.branch_streams)
    conf = CONFIGS[args.config]
    hp = conf["train"]
    step_gflop = conf["gflop"]
    B, T, nspk = TRAIN["batch_size"], TRAIN["max_segment"], MODEL["nspk"]
    G, D = build_models(dev)
    broadcast_parameters(G); broadcast_parameters(D)
    oG = FusedAdamW(G.parameters(), TRAIN["lr"], TRAIN["betas"])
    oD = FusedAdamW(D.parameters(), TRAIN["lr"], TRAIN["betas"])
    Cm = oC = None
    if hp["lambda_latcls"] != 0:
        from model.latent_classifier import LatentClassifier
        Cm = LatentClassifier(nspk, MODEL["content_dim"]).to(dev)
        broadcast_parameters(Cm)
        oC = FusedAdamW(Cm.parameters(), TRAIN["lr"], (0.9, 0.999), weight_decay=0.0)     # torch.optim.Adam, train.py:192
    hook = GradAverager() if world > 1 else None
    ts = TrainStep(G, D, hp, oG, oD, nspk, grad_hook=hook, C=Cm, optimizer_C=oC)
    if world > 1 and args.overlap and not args.graph:
        # eager launches: bucketed gradient all-reduce issued from autograd hooks, overlapped with the backward.  Under
        # CUDA-graph replay (the default) the all-reduces run between graph segments instead: capturing the hook-driven
        # NCCL calls into the step graph deadlocked on this stack (torch 2.11 / NCCL 2.28.9, 2 x B200, measured).
        ts.enable_overlap(bucket_mb=args.bucket_mb)
    host = synth_batch(B, T, nspk, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident, _ = to_device(host, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    last = {}
    h2d = [0]
    d2h = [0]
    nbytes_in = sum(v.numel() * v.element_size() for v in host.values())
    if args.graph:
        # the whole step (both backward passes, all-reduce, optimisers) is one CUDA graph; eager warm-up inside
        n0 = lib.tdvc_launch_count()
        f0 = [lib.tdvc_flop_count(i) for i in range(6)]
        graphed = GraphedTrainStep(ts, resident, warmup=2)
        n_cap = lib.tdvc_launch_count() - n0
        launches_per_step = n_cap / 3.0          # 2 eager warm-up steps + 1 captured step
        fam_gflop = [(lib.tdvc_flop_count(i) - f0[i]) / 3.0 / 1e9 for i in range(6)]

        def step_resident():
            last.update(graphed.step())

        def step_e2e():
            graphed.load(pinned)                                  # H2D of this step's inputs (pinned host memory)
            out = graphed.step()
            losses = torch.stack([out["d_loss"].reshape(()), out["g_loss"].reshape(())]).cpu()   # D2H of the result
            h2d[0], d2h[0] = nbytes_in, losses.numel() * losses.element_size()
            last["host_losses"] = losses
    else:
        launches_per_step = None
        fam_gflop = None

        def step_resident():
            last.update(ts.step(resident))

        def step_e2e():
            batch, nb = to_device(host, dev, pinned)
            out = ts.step(batch)
            losses = torch.stack([out["d_loss"].reshape(()), out["g_loss"].reshape(())]).cpu()
            h2d[0], d2h[0] = nb, losses.numel() * losses.element_size()
            last["host_losses"] = losses

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.tdvc_launch_count()
    if sampler:
        sampler.start()
    if args.profile:
        torch.cuda.profiler.start()      # ncu --profile-from-start off: the profiled range is exactly the timed steps
    ms = timed(step_resident, args.steps)
    if args.profile:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    launches = launches_per_step if launches_per_step is not None else (lib.tdvc_launch_count() - n0) / max(1, args.steps)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": ms / args.steps, "gpu_launches": launches}))
        if world > 1:
            dist.destroy_process_group()
        return
    # end-to-end arm: host buffers in, loss scalars out, copies inside the timed region
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    audio_s = world * B * T / SR
    value = audio_s * args.steps / (ms / 1e3)
    e2e = audio_s * args.steps / (ms_e2e / 1e3)

    out = None
    fams = None
    if args.graph:
        # every rank replays (the step holds collectives when N > 1: a rank-0-only replay would wait for its peers
        # forever); rank 0's trace is the one reported
        try:
            fams = kernel_families(graphed.step, fam_gflop)
        except Exception as e:          # never lose the headline line to the trace
            fams = {"error": repr(e)[:200]}
    if rank == 0:
        pk = peaks()
        flop_dom = dominant_kernel_roofline(dev, B, T, pk, args)
        ms_step = ms / args.steps
        roof = time_dominant_roofline(fams, pk, ms_step, step_gflop, flop_dom)
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32" if args.precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": {"workload": conf["workload"], "name": args.config,
                       "batch_per_gpu": B, "segment_samples": T, "sample_rate": SR, "speakers": nspk,
                       "lambda_f0": "0 (torchcrepe unavailable offline, SURVEY 8c)", "precision": args.precision,
                       "l2": "no explicit flush: one step streams >6 GB of activations, far larger than the 126 MB L2",
                       "parallelism": f"dp{world}", "step_gflop_algorithmic": step_gflop * world,
                       "grad_allreduce": ("none (1 GPU)" if world == 1 else
                                          (f"bucketed ({args.bucket_mb} MB), issued from autograd hooks, overlapped with the backward"
                                           if (args.overlap and not args.graph) else
                                           "one flat all-reduce per network between CUDA-graph segments")),
                       "cuda_graph": bool(args.graph)},
            "clocks": clocks,
            "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": d2h[0],
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(round(launches)),
            "step_tflops_algorithmic": round(step_gflop * world / (ms / args.steps), 3),
            "step_frac_of_sustained_tensor_peak": round(step_gflop / (ms / args.steps) / pk["tc_sustained"], 5),
            "roofline": roof,
            "roofline_flop_dominant_kernel": flop_dom,
            "kernel_families": fams,
            "losses": {k: float(v) for k, v in last.items() if k in ("d_loss", "g_loss")},
        }
        if world == 1 and not args.no_inference:
            try:
                out["inference"] = inference_sweep(G, dev, args)
            except Exception as e:      # never lose the headline line to the auxiliary measurement
                out["inference"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(steps=1, warmup=0, hp=hp)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


# conv kernel families = the counters of tdvc_flop_count(): a kernel name (CUPTI, template arguments kept) -> the family whose
# 2*MAC counter its launches feed (csrc: g_flops[...]).  The chain epilogues (second template argument >= 3) are instances of
# the ws / fwdh / fwd kernels but count as their own family; the finalize passes of the weight gradients belong to them.
FAMILY_NAME = {0: "conv_tc_fwd_k + conv_tc_fwdh_k (tcgen05, one tile per CTA, streamed weights: C >= 128 layers, short sequences, their data gradients)",
               1: "conv_tc_ws_k (tcgen05, weight-stationary persistent)", 2: "conv_tc_wt_k (tcgen05, stacked cond_var.0)",
               3: "conv_tc_wgrad2_k + conv_tc_wgrad2s_k + finalize (tcgen05 weight gradients)",
               4: "fp32 CUDA-core convs (1-channel stems, 8-channel excitation pyramid, FIR filters, their gradients)",
               5: "MRF chain epilogues of conv_tc_ws_k / conv_tc_fwdh_k (EPI 3-6, tcgen05)"}


def family_of(name):
    """Family index (FAMILY_NAME) of a kernel name, or None for kernels that are not convolutions."""
    import re
    m = re.search(r"conv_tc_(ws|fwdh|fwd)_k<\s*\d+,\s*(\d+)", name)
    if m:
        if int(m.group(2)) >= 3:
            return 5
        return 1 if m.group(1) == "ws" else 0
    if "conv_tc_wt_k" in name:
        return 2
    if re.search(r"conv_tc_wgrad\w*_k|wgrad2s?_finalize\w*_k|wgrad_finalize\w*_k", name):
        return 3
    if re.search(r"(^|:)(conv_fwd_k|conv_tr_k|conv_wgrad_k|narrow_wgrad_k|stem_wgrad_k)", name):
        return 4
    return None


def kernel_families(replay, fam_gflop):
    """Per-kernel-family busy time of ONE CUDA-graph replay of the step (CUPTI through torch.profiler, after the timed
    region; kernels overlap slightly under programmatic dependent launch, so the sum can exceed the wall time), next to
    the multiply-accumulate work the library handed to each conv family (tdvc_flop_count deltas of one step)."""
    import collections
    import re
    import torch
    from torch.profiler import ProfilerActivity, profile
    replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if not evs:
        return {"error": "no CUDA activities in the trace"}
    # kernels run back to back in one stream; with programmatic dependent launch a kernel is resident (waiting in
    # griddepcontrol.wait) while its predecessor still runs, so its CUPTI duration overlaps the predecessor's.  A kernel is
    # charged the part of its interval that follows the end of everything before it: the charges add up to the busy span.
    evs.sort(key=lambda e: e.time_range.start)
    busy, cnt = collections.Counter(), collections.Counter()
    frontier = evs[0].time_range.start
    idle = 0.0
    inst = []
    for e in evs:
        name = re.sub(r"\(.*", "", e.name)
        name = re.sub(r"^void ", "", name)[:60]
        st, en = e.time_range.start, e.time_range.end
        if st > frontier:
            idle += (st - frontier) / 1e3
        busy[name] += max(0.0, en - max(st, frontier)) / 1e3
        inst.append((max(0.0, en - max(st, frontier)), name, len(inst)))
        frontier = max(frontier, en)
        cnt[name] += 1
    t0 = evs[0].time_range.start
    t1 = frontier
    total = sum(busy.values())
    top = [{"kernel": k, "launches": cnt[k], "ms": round(v, 3), "share_of_busy": round(v / total, 4)}
           for k, v in sorted(busy.items(), key=lambda kv: -kv[1])[:24]]
    conv = {}
    for k, v in busy.items():
        fam = family_of(k)
        if fam is not None:
            d = conv.setdefault(fam, {"family": FAMILY_NAME[fam], "id": fam, "launches": 0, "ms": 0.0})
            d["launches"] += cnt[k]
            d["ms"] += v
    for fam, d in conv.items():
        d["ms"] = round(d["ms"], 3)
        if fam_gflop is not None:
            d["gflop_per_step"] = round(fam_gflop[fam], 2)
            d["tflops"] = round(fam_gflop[fam] / d["ms"], 2) if d["ms"] > 0 else None
    slowest = [{"kernel": n, "us": round(d, 1), "position": i} for d, n, i in sorted(inst, reverse=True)[:30]]
    return {"span_ms": round((t1 - t0) / 1e3, 3), "busy_ms": round(total, 3), "idle_ms": round(idle, 3),
            "activities": len(evs), "top": top, "slowest_launches": slowest,
            "conv_families": [conv[k] for k in sorted(conv, key=lambda k: -conv[k]["ms"])]}


# families whose tensor-core launches were ALL captured by the final-build ncu metric pass (profiles/r2/final_ncu_conv_metrics.json:
# -k regex:conv_tc_ws_k|conv_tc_fwdh_k|conv_tc_wgrad2s_k|conv_tc_wt_k); the others hold conv_tc_fwd_k / conv_tc_wgrad2_k launches too
NCU_COMPLETE = (1, 2)


def ncu_family_traffic(fam_id):
    """`traffic` of a conv family: dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the family's launches
    inside one step, from the committed ncu capture; None (with a note on what was captured) when the capture does not
    hold every launch of the family."""
    try:
        with open(os.path.join(REPO, "profiles", "r2", "final_ncu_conv_metrics.json")) as f:
            inst = json.load(f)["instances"]
    except Exception:
        return {"traffic": None}
    n = nbytes = tens = us = 0.0
    for name, k in inst.items():
        if family_of(name) == fam_id:
            n += k["launches"]
            nbytes += k["launches"] * k["dram_bytes_per_launch"]
            us += k["launches"] * k["us_per_launch"]
            tens += k["launches"] * k["us_per_launch"] * k["tensor_pipe_pct_of_active"]
    src = "profiles/r2/final_ncu_conv_metrics.json (ncu, one eager step of the 21.3 ms/step build)"
    if n == 0:
        return {"traffic": None, "traffic_note": f"no launch of this family in {src}"}
    got = {"launches": int(n), "dram_bytes_per_launch": round(nbytes / n), "tensor_pipe_pct_of_active": round(tens / us, 2)}
    if fam_id in NCU_COMPLETE:
        return {"traffic": got["dram_bytes_per_launch"], "tensor_pipe_pct_of_active_ncu": got["tensor_pipe_pct_of_active"],
                "traffic_source": f"{src}: dram__bytes_read.sum + dram__bytes_write.sum, mean over the family's {int(n)} launches"}
    return {"traffic": None, "traffic_note": f"{src} holds only part of this family's launches: {got}"}


def time_dominant_roofline(fams, pk, ms_step, step_gflop, flop_dom):
    """`roofline` describes what bounds the STEP: the conv kernel family with the largest share of the step's time, its
    achieved rate = (2*MAC handed to it per step) / (its busy time in one replay) against the SUSTAINED dense bf16 peak
    (it runs inside a long step); the whole-step fraction and the FLOP-dominant launch are reported beside it."""
    peak = pk["tc_sustained"]
    base = {"bound": "tensor", "peak": peak, "unit": "TFLOP/s", "traffic": None,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step), {pk['src']}",
            "step": {"achieved": round(step_gflop / ms_step, 3), "frac": round(step_gflop / ms_step / peak, 5),
                     "what": "algorithmic FLOPs of the whole reference step / step time"}}
    try:
        fam = fams["conv_families"][0]
        ach = fam["tflops"]
        base.update(achieved=ach, frac=round(ach / peak, 5), kernel=fam["family"], launches_per_step=fam["launches"],
                    ms_per_step=fam["ms"], gflop_per_step=fam["gflop_per_step"],
                    share_of_step_time=round(fam["ms"] / ms_step, 4),
                    how="2*MAC handed to the family per step (tdvc_flop_count) / its busy time in one CUDA-graph replay "
                        "(CUPTI, taken inside bench.py after the timed region)")
        base.update(ncu_family_traffic(fam.get("id")))
    except Exception:
        base.update(achieved=flop_dom["achieved"], frac=round(flop_dom["achieved"] / peak, 5), kernel=flop_dom["kernel"],
                    how="family trace unavailable: the FLOP-dominant launch timed alone")
    return base


def dominant_kernel_roofline(dev, B, T, pk, args):
    """The launch that carries most of the step's FLOPs: the 9 FiLM `cond_var.0` convolutions (136->136, k=3,
    'same') of the full-rate decoder stage (143 of 465 GF of a G forward).  bf16 mode runs them as
    ONE tcgen05 launch (N = 9*144) writing the packed bf16 operand of cond_var.2; fp32 mode as 9 CUDA-core launches.
    Timed alone, CUDA events around a CUDA-graph replay of `reps` launches on the capturing stream (so no host
    launch gaps); operands 41 MB in, 371 MB out per launch (> L2).  achieved = algorithmic FLOPs / launch time
    against the measured dense bf16 tensor-core burst peak."""
    import torch
    from tdvc import ops
    from tdvc import _lib
    from tdvc._lib import ACT_LRELU
    Cc, n, K = MODEL["cond_dim"] + 8, 9, 3
    flops = 2.0 * B * T * Cc * Cc * K * n
    reps = 5
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        if args.precision == "bf16":
            Cg = 144
            cp = (torch.randn(B, T, Cg, device=dev) * 0.5).to(torch.bfloat16)
            outs = [torch.empty(B, T, n * Cg, device=dev, dtype=torch.bfloat16) for _ in range(2)]
            if ops._STACKED_COND:
                # what _MRFCondPath.forward launches: weights stacked densely (9 x 136 rows) as the M operand
                R0 = (n - 1) * Cc + Cg
                w0p = (torch.randn(K, R0, Cg, device=dev) * 0.05).to(torch.bfloat16)
                b0p = torch.zeros(R0, device=dev)
                lib = _lib.load()

                def launch(i):
                    _lib.check(lib.tdvc_conv1d_tc_fwd_stacked(cp.data_ptr(), w0p.data_ptr(), b0p.data_ptr(), outs[i % 2].data_ptr(),
                                                              B, Cg, 0, Cg, T, T, K, 1, -1, R0, n, Cc, ACT_LRELU, 0.2, T, n * Cg,
                                                              0, 0, Cg, ops._st()), "stacked")
                kname = ("conv_tc_wt_k<ACT=lrelu> (tcgen05 bf16, M=128 stacked weight rows resident in smem, N=256 time steps; "
                         "9 cond_var.0 convs in one launch)")
            else:
                w0p = (torch.randn(K, n * Cg, Cg, device=dev) * 0.05).to(torch.bfloat16)
                b0p = torch.zeros(n * Cg, device=dev)

                def launch(i):
                    ops._tc_conv(xp=cp, wp=w0p, bias=b0p, B=B, Tp=T, Tout=T, K=K, dilation=1, t_off=-1, Cp_total=Cg, groups=1,
                                 a_ch_off=0, a_ch_stride=0, Cinp_g=Cg, Cout_g=n * Cg, Coutp_g=n * Cg, bias_stride=0,
                                 out_act=ACT_LRELU, out_slope=0.2, out_packed=1, yp=outs[i % 2], tp_out=T, cp_out=n * Cg,
                                 out_halo=0, out_ch_off=0, out_ch_stride=0)
                kname = "conv_tc_ws_k<ACT=lrelu,EPI=0,OUT=packed,MASK=0> (tcgen05 bf16, 9 cond_var.0 convs in one launch)"
            launches_per_rep = 1
        else:
            x = torch.randn(B, Cc, T, device=dev)
            ws = [torch.randn(Cc, Cc, K, device=dev) * 0.05 for _ in range(n)]
            bs = torch.zeros(Cc, device=dev)

            def launch(i):
                for w in ws:
                    ops.conv1d(x, w, bs, padding=1)
            kname = "conv_fwd_k<8,32> (fp32 CUDA cores, 9 launches)"
            launches_per_rep = n
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            launch(0)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for i in range(reps):
                    launch(i)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ach = flops / (ms * 1e-3) / 1e12
    peak = pk["tc_burst"]
    traffic = None
    try:   # per-launch DRAM bytes of this kernel from the committed ncu --set full capture, if present
        with open(os.path.join(REPO, "profiles", "r1_dominant_kernel_ncu.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    return {"bound": "tensor", "achieved": round(ach, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 5),
            "traffic": traffic, "kernel": kname, "shape": f"B={B} T={T} Cin={Cc} Cout={n}x{Cc} K={K}",
            "flops_per_launch": flops / launches_per_rep, "ms_per_launch": round(ms / launches_per_rep, 4),
            "algorithmic_bytes_per_launch": B * T * (Cc * 2 + n * Cc * 2) if args.precision == "bf16" else None,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone), {pk['src']}"}


def inference_sweep(G, dev, args, seconds=4.0, batches=(1, 4, 16, 64, 128)):
    """BASELINE config 5: batched `G(signal, c_tgt, c_var=excitation)` (generate_with_target.py:169) on synthetic 4 s
    utterances, forward only (no_grad), weight norm + bf16 operand packing folded ONCE (ops.inference_cache), one CUDA
    graph per batch size; RTF = wall time / audio seconds.  Batch sizes are bounded (no max-fit search on a shared box)."""
    import torch
    from tdvc import ops
    T = int(seconds * SR)
    rows = []
    side = torch.cuda.Stream()
    with torch.no_grad(), ops.inference_cache():
        for B in batches:
            host = synth_batch(B, T, MODEL["nspk"], seed=4321)
            x, cv = host["signal_real"].to(dev), host["c_f0_conv"].to(dev)
            c_tgt = torch.zeros(B, MODEL["nspk"], device=dev).scatter_(1, host["label_tgt"].to(dev).view(-1, 1), 1.0)
            try:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        G(x, c_tgt, c_var=cv)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    y = G(x, c_tgt, c_var=cv)
                g.replay()
                torch.cuda.synchronize()
                reps = 5 if B <= 16 else 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
            except torch.OutOfMemoryError:
                rows.append({"batch": B, "error": "out of memory"})
                break
            ms = e0.elapsed_time(e1) / reps
            audio = B * seconds
            rows.append({"batch": B, "ms_per_batch": round(ms, 3), "rtf": round(ms / 1e3 / audio, 8),
                         "x_realtime": round(audio / (ms / 1e3), 1), "tflops_algorithmic": round(51.95 * audio / ms, 2),
                         "finite": bool(torch.isfinite(y).all())})
            del g, y, x, cv
            torch.cuda.empty_cache()
    best = max((r for r in rows if "x_realtime" in r), key=lambda r: r["x_realtime"])
    return {"utterance_s": seconds, "gflop_per_audio_s": 51.95, "weight_norm": "folded once (ops.inference_cache)",
            "sweep": rows, "best": best, "rtf": best["rtf"], "x_realtime": best["x_realtime"], "batch": best["batch"]}


def cpu_baseline(steps, warmup, hp=None, batch=None):
    """The reference algorithm (CPU oracle port, fp32, all host threads) on the SAME workload as the CUDA arm: full G+D
    steps (forward, both backward passes, torch AdamW updates) at B = 16, T = 8960; audio-s/s = B*0.56 s / step time.
    About 7 s per step on 16 cores, so the bounded sample is simply `steps` whole steps."""
    import torch
    from oracle.step import make_models, oracle_step, step_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = dict(hp if hp is not None else TRAIN)
    hp["faithful_waste"] = True      # the reference back-propagates the sub-scale heads into G in the D step
    Bs = int(batch or TRAIN["batch_size"])
    cfg = dict(MODEL, B=Bs, T=TRAIN["max_segment"], seed=6)
    sdG, sdD = make_models(cfg, torch.float32)
    for v in list(sdG.values()) + list(sdD.values()):
        v.requires_grad_(True)
    oG = torch.optim.AdamW(list(sdG.values()), TRAIN["lr"], TRAIN["betas"])
    oD = torch.optim.AdamW(list(sdD.values()), TRAIN["lr"], TRAIN["betas"])
    data = step_batch(cfg, hp, torch.float32)

    def one():
        out = oracle_step(cfg, hp, torch.float32, sdG, sdD, data)
        for k, v in sdD.items():
            v.grad = out["D_grad"][k]
        for k, v in sdG.items():
            v.grad = out["G_grad"][k]
        oD.step(); oG.step()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    v = Bs * TRAIN["max_segment"] / SR / dt
    return {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} full G+D step(s) (fwd, both bwd, AdamW) at B={Bs}, T=8960, fp32, {warmup} warm-up; "
                      f"{dt:.2f} s/step",
            "s_per_step": round(dt, 3), "batch": Bs}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores -- the oracle port
    (`kind: "port"`): the reference itself is pure Python/PyTorch and /root/reference does not exist on the GPU box -- at
    the CUDA arm's own workload (same config, B = 16 per step, same metric)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    conf = CONFIGS[args.config]
    steps = max(1, args.steps)
    warm = min(args.warmup, 1)
    cb = cpu_baseline(steps=steps, warmup=warm, hp=conf["train"])
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": warm, "ms_per_step": round(cb["s_per_step"] * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": {"workload": conf["workload"], "name": args.config, "batch_per_gpu": cb["batch"],
                      "segment_samples": TRAIN["max_segment"], "sample_rate": SR, "speakers": MODEL["nspk"],
                      "lambda_f0": "0 (torchcrepe unavailable offline, SURVEY 8c)", "precision": "fp32",
                      "device": f"host CPU, {cb['cores']} threads (rank 0 only; the GPUs are not used by this arm)",
                      "sample": f"each step is one full G+D iteration at B={cb['batch']} (the CUDA arm's batch per GPU)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    # stdout carries exactly one JSON line: anything a library prints there (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TDVC_PRECISION", "bf16"), choices=["fp32", "bf16"],
                    help="bf16 = tcgen05 tensor-core path (2e-2 parity, the headline); fp32 = exact CUDA-core path (1e-5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", type=int, default=int(os.environ.get("TDVC_DP_OVERLAP", "1")),
                    help="N > 1 with --no-graph: 1 = bucketed gradient all-reduce overlapped with the backward, 0 = one "
                         "flat all-reduce per network after its backward (what the CUDA-graph path always does)")
    ap.add_argument("--bucket-mb", type=float, default=16.0)
    ap.add_argument("--no-inference", action="store_true", help="skip the BASELINE config 5 inference sweep")
    ap.add_argument("--config", default="stage1", choices=sorted(CONFIGS),
                    help="which shipped train config the step follows (BASELINE.json configs 1, 2, 4; the wavlm configs "
                         "need the WavLM front end, which is outside this path: unmeasured)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel from Python each step instead of replaying one CUDA graph")
    ap.add_argument("--branch-streams", type=int, default=int(os.environ.get("TDVC_BRANCH_STREAMS", "0")),
                    help="1: run the 3 independent MRF branches on forked CUDA streams")
    ap.add_argument("--profile", action="store_true",
                    help="profiling aid (ncu): honour --warmup as given, skip the e2e / roofline / cpu legs; "
                         "numbers printed in this mode are not bench values")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3 and not args.profile:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
