#!/usr/bin/env python
"""bench.py — visual-encoder clips/sec on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a path (libsblk)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)
  torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, clip batch sharded (weak scaling)

A step = one forward of the hot path (Conv3d frontend -> ResNet-18 trunk -> 6-layer transformer Encoder,
always-on dropout included, exactly what `Transformer.forward` runs before the decoder,
SBL_Multilingual_Lip_reading/transformer/transformer.py:34-38) over one synthetic LRW-shaped batch
(BASELINE.json configs[1]: 32 clips x 29 frames x 88 x 88 per GPU).

  value   device-timed (CUDA events on the launching stream), inputs resident in HBM, L2 flushed between steps
  e2e     same metric through the host-facing call: pinned host fp32 clips in, H2D + forward + D2H of the encoder
          output every step (double-buffered), wall clock bracketed by synchronize
  roofline        dominant kernel: algorithmic FLOPs per launch / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline    the CPU oracle (port of the reference algorithm; same torch primitives) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "visual_encoder_clips_per_sec"
UNIT = "clips/s"
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def flops_per_clip(t, layers):
    """SURVEY.md §8d / BASELINE.md §3: 2*MACs of Conv3d + 20 ResNet convs + linear_in + per-layer linears + bmm."""
    return t * (60712960 + 571604992 + 524288 + layers * 6291456) + layers * 2048 * t * t


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            if "bf16_tflops" in d and "hbm_gbs" in d:
                d["_source"] = "measured"
                return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_setup(layers):
    import torch
    from oracle import visual_encoder_oracle as O
    from sbl_for_multilingual_lip_reading_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {}
    sd.update(synth.frontend_state_dict(1, prefix="visual_frontend."))
    sd.update(synth.encoder_state_dict(2, layers, prefix="encoder."))
    return O, sd, cores


def cpu_oracle_step(O, sd, x):
    import torch
    with torch.no_grad():
        gen = torch.Generator().manual_seed(0)
        n, t = x.shape[0], x.shape[2]
        mask = (torch.rand((n * t, 512), generator=gen) >= 0.5).float() * 2.0  # always-on dropout(0.5), x2 scaling
        return O.visual_encoder_forward(x, sd, dropout_mask=mask)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from sbl_for_multilingual_lip_reading_b200 import synth
    O, sd, cores = cpu_oracle_setup(args.layers)
    sample = args.ref_clips if args.ref_clips > 0 else (4 if args.steps <= 100 else 2 if args.steps <= 300 else 1)
    x = synth.synthetic_clips(sample, args.frames, seed=7)
    for _ in range(max(args.warmup, 1)):
        cpu_oracle_step(O, sd, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(O, sd, x)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    desc = f"{sample} clips x {args.frames} frames per step (bounded sample of the {args.batch}-clip batch)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    if (args.batch, args.frames) == (32, 29):
        name, shape = "BASELINE configs[1]", "LRW-shaped"
    elif (args.batch, args.frames) == (8, 40):
        name, shape = "BASELINE configs[2] (batch 64 sharded across 8 B200: 8 clips per GPU)", "LRW-1000-shaped"
    else:
        name, shape = "custom shape", "LRW-style"
    return {
        "workload": (f"{name}: full visual encoder forward (Conv3d frontend + ResNet-18 trunk + "
                     f"{args.layers}-layer transformer Encoder), {shape} {args.frames}x88x88 gray clips, "
                     f"batch {args.batch} per B200"),
        "clips_per_gpu": args.batch, "frames": args.frames, "encoder_layers": args.layers,
        "global_batch": args.batch * world, "parallelism": f"dp{world}",
        "l2": "L2 flushed (256 MiB write) between timed steps; e2e inputs stream from pinned host memory",
    }


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run on (and first-touch the pinned host buffers from) the CPUs of the NUMA node the GPU hangs
    off, so the per-step host->device copy does not cross the socket interconnect.  Returns the node or None."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference)")
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner ignores
    # NCCL_DEBUG_FILE) is sent to stderr by pointing fd 1 at fd 2 for the whole run; the line goes to the saved fd
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout; stdout carries exactly ONE JSON line, so NCCL's log goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/sblk_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    from sbl_for_multilingual_lip_reading_b200 import ops, sharding, synth
    from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
    from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan, VisualEncoderPlan
    from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend

    ops.init()
    B, T, L = args.batch, args.frames, args.layers
    fe = visual_frontend(None)
    fe.load_state_dict(synth.frontend_state_dict(1))
    enc = Encoder(512, L, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
    enc.load_state_dict(synth.encoder_state_dict(2, L))
    fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
    torch.manual_seed(1234 + rank)

    # Throughput plan: two-stage software pipeline (the encoder stack of batch i-1 co-runs with the clip prep + stem of
    # batch i, runner.PipelinedVisualEncoderPlan); shapes it does not cover and --no-pipeline use the one-batch plan.
    pipelined, pipeline_note = False, "off (--no-pipeline)"
    plan = None
    if not args.no_pipeline:
        try:
            plan = PipelinedVisualEncoderPlan(fe, enc, B, T, device=dev, pdl=not args.no_pdl)
            pipelined = True
            pipeline_note = (f"2-stage software pipeline: step i = encoder stack of batch i-1 (8-CTA clusters) next to clip "
                             f"prep + Conv3d stem of batch i on {plan.head_sm_limit} SMs, then the trunk of batch i at "
                             f"full width; value/e2e are steady-state throughput, `latency` is the unpipelined plan")
        except Exception as e:  # noqa: BLE001  (shape outside the plan's envelope: the one-batch plan is timed instead)
            pipeline_note = f"off ({e})"
    plan_lat = VisualEncoderPlan(fe, enc, B, T, device=dev, slots=2, pdl=not args.no_pdl)
    if plan is None:
        plan = plan_lat

    def make_plan(**kw):
        if pipelined:
            return PipelinedVisualEncoderPlan(fe, enc, B, T, device=dev, pdl=not args.no_pdl, **kw)
        return VisualEncoderPlan(fe, enc, B, T, device=dev, slots=2, pdl=not args.no_pdl, **kw)

    # synthetic inputs: a pool of distinct batches (host pinned for e2e; device copies for the device-timed run)
    pool_n = 4
    host_pool = [synth.synthetic_clips(B, T, seed=100 + 17 * rank + i).pin_memory() for i in range(pool_n)]
    dev_pool = [h.to(dev) for h in host_pool]
    out_host = [torch.empty((B, T, 512), dtype=torch.float32).pin_memory() for _ in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * B, T, 512), dtype=torch.float32, device=dev) if world > 1 else None
    # output gathering: one-shot peer-memory kernel over NVLink (sharding.P2PGather) or, with --gather nccl, NCCL
    p2p, gather_mode = None, (args.gather if world > 1 else None)
    if world > 1 and args.gather == "p2p":
        err = None
        try:
            p2p = sharding.P2PGather(B * T * 512, dev)
        except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC not permitted in this container)
            err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # the choice must be the same on every rank
        if int(ok.item()) == 0:
            p2p, gather_mode = None, f"nccl (peer-memory gather unavailable: {err})"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather_outputs(out):
        if p2p is not None:
            # output gathering (the DataParallel `gather` of the reference, train.py:115): one kernel that stores this
            # rank's block into every peer's buffer over NVLink and waits for all peers' blocks
            p2p(out.view(-1))
        elif gathered is not None:
            dist.all_gather_into_tensor(gathered, out)

    gather_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    pending = [None]

    def one_device_step(i, timed, plan=plan, defer_gather=pipelined and world > 1 and not args.inline_gather):
        """inputs already in HBM; L2 flushed before the timed part; returns (start, end) events.  With the pipelined plan
        the step's output is the previous batch's: one frontend pass + one encoder pass per step.  Multi-GPU: every step
        contains exactly one output gather inside its event pair — in stream after the replay, or (pipelined plan) the
        gather of the PREVIOUS step's output on a side stream next to this step's replay, so that a rank waiting for
        its peers' blocks keeps computing (the output buffer it reads is not written by this replay)."""
        s = i % plan.slots
        with torch.cuda.stream(plan.compute):
            plan.x[s].copy_(dev_pool[i % pool_n])
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(plan.compute)
            if defer_gather and pending[0] is not None:
                gather_stream.wait_event(e0)
                with torch.cuda.stream(gather_stream):
                    gather_outputs(pending[0])
            out = plan.forward_device(s)
            if defer_gather:
                plan.compute.wait_stream(gather_stream)
                pending[0] = out
            else:
                gather_outputs(out)
            e1.record(plan.compute)
        return e0, e1

    def flush_pending_gather():
        """the last deferred gather (every rank issues the same number of gathers)"""
        if pending[0] is not None:
            with torch.cuda.stream(plan.compute):
                gather_outputs(pending[0])
            pending[0] = None

    sampler = ClockSampler(local_rank)

    # ---- device-timed run -------------------------------------------------------------------
    for i in range(args.warmup):
        one_device_step(i, False)
    barrier()
    sampler.start()
    evs = [one_device_step(args.warmup + i, True) for i in range(args.steps)]
    barrier()
    sampler.stop()
    flush_pending_gather()
    barrier()
    dev_ms = sharding.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs), dev)
    value = world * B * args.steps / (dev_ms * 1e-3)

    # ---- one-batch latency of the unpipelined plan (same timing rules; extra key, not the metric) ----------
    latency = None
    if pipelined:
        for i in range(3):
            one_device_step(i, False, plan_lat, False)
        barrier()
        lat_n = min(args.steps, 20)
        evs_l = [one_device_step(3 + i, True, plan_lat, False) for i in range(lat_n)]
        barrier()
        lat_ms = sharding.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs_l), dev) / lat_n
        latency = {"ms_per_step": lat_ms, "clips_per_s": world * B / (lat_ms * 1e-3), "steps": lat_n,
                   "plan": "VisualEncoderPlan (one batch per replay, nothing overlapped across batches)"}

    # ---- end-to-end run (host buffers, H2D + D2H inside the timed region) --------------------
    def e2e_steps(k):
        for i in range(k):
            plan.submit_host(host_pool[i % pool_n], out_host[i % 2])
        if pipelined:   # k inputs in, the k outputs OF THOSE inputs out: the last batch's encoder runs here
            plan.drain(out_host[k % 2])
        plan.synchronize()

    e2e_steps(max(args.warmup, 3))
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    sampler.stop()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s
    checksum = float(out_host[(args.steps - 1) % 2].double().abs().mean())  # the D2H result is really read

    # ---- end-to-end through the fused uint8 input pipeline (an ADDITIONAL key: `e2e` above keeps the
    # reference's fp32 model boundary): the host ships the loader's raw uint8 frames [B, T, 96, 96] (9.2 KB/frame
    # instead of 31 KB) and /255, ColorNormalize, centre crop are done by the clip-prep kernel ----------------------
    e2e_u8 = None
    if not args.no_u8:
        plan8 = make_plan(u8_input=(T, 96, 96))
        host_u8 = [synth.synthetic_u8_clips(B, T, seed=300 + 17 * rank + i).pin_memory() for i in range(pool_n)]

        def u8_steps(k):
            for i in range(k):
                plan8.submit_host(host_u8[i % pool_n], out_host[i % 2])
            if pipelined:
                plan8.drain(out_host[k % 2])
            plan8.synchronize()

        u8_steps(max(args.warmup, 3))
        barrier()
        t0 = time.perf_counter()
        u8_steps(args.steps)
        torch.cuda.synchronize(dev)
        u8_s = sharding.max_over_ranks(time.perf_counter() - t0, dev)
        e2e_u8 = {"value": world * B * args.steps / u8_s, "unit": UNIT, "h2d_bytes_per_step": B * T * 96 * 96,
                  "d2h_bytes_per_step": B * T * 512 * 4, "ms_per_step": 1e3 * u8_s / args.steps,
                  "input": "raw uint8 gray frames [B,T,96,96] from pinned host memory; /255, ColorNormalize, 88x88 "
                           "centre crop fused into the clip-prep kernel (bit-identical to the fp32 path)",
                  "result_checksum": float(out_host[(args.steps - 1) % 2].double().abs().mean())}

    # ---- per-kernel roofline (rank 0): eager traced passes, PDL off so every launch is timed alone --------
    line_extra = {}
    if rank == 0:
        peaks = load_peaks()
        prev_pdl = ops.set_pdl(False)
        prev_cl = enc.stack_cluster_size
        if pipelined:
            enc.stack_cluster_size = plan.enc_cluster   # the launch the timed region uses (one launch, one cluster size)
        agg = {}
        passes = 5
        with torch.no_grad():
            for rep in range(passes + 1):
                sink = []
                flush.zero_()
                with ops.trace(sink):
                    out, = enc(fe(dev_pool[rep % pool_n]), [T] * B)
                torch.cuda.synchronize(dev)
                if rep == 0:
                    continue  # warm-up pass
                for r in sink:
                    key = (r["name"], r["tag"])
                    a = agg.setdefault(key, {"ms": 0.0, "launches": 0, "flops": r["flops"], "bytes": r["bytes"]})
                    a["ms"] += r["start"].elapsed_time(r["end"])
                    a["launches"] += 1
        ops.set_pdl(bool(prev_pdl))
        enc.stack_cluster_size = prev_cl
        total_ms = sum(a["ms"] for a in agg.values()) / passes
        rows = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        breakdown = []
        for (name, tag), a in rows[:12]:
            per = a["ms"] / a["launches"]
            breakdown.append({"kernel": name, "case": tag, "launches_per_step": a["launches"] // passes,
                              "avg_us": round(per * 1e3, 2), "share": round(a["ms"] / passes / total_ms, 4),
                              "tflops": round(a["flops"] / (per * 1e-3) / 1e12, 1) if a["flops"] else None,
                              "gbs": round(a["bytes"] / (per * 1e-3) / 1e9, 1)})
        traffic_table = {}
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic_table = json.load(f)
            except Exception:
                traffic_table = {}

        def roofline_of(row):
            (name, tag), a = row
            per_s = a["ms"] / a["launches"] * 1e-3
            if a["flops"]:
                ach, peak = a["flops"] / per_s / 1e12, float(peaks["bf16_tflops"])
                roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak}
            else:
                ach, peak = a["bytes"] / per_s / 1e9, float(peaks["hbm_gbs"])
                roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}
            roof.update({"traffic": traffic_table.get(f"{name}|{tag}"), "kernel": name, "case": tag,
                         "avg_launch_us": per_s * 1e6, "launches_per_step": a["launches"] // passes,
                         "peak_source": peaks["_source"] + " (MEASURED_PEAKS.json burst figure: kernel timed alone)"})
            return roof

        roof = roofline_of(rows[0])
        if pipelined and rows[0][0][0] == "sblk_encoder_stack_fwd":
            # in the pipelined plan the stack runs on the side stream next to the next batch's prep + stem: the largest
            # kernel group of the step's critical path is reported as well
            roof["note"] = ("latency-bound dependent chain (25 GEMM stages on 64 SMs); overlapped with the next batch's "
                            "clip prep + stem by the pipelined plan, see roofline_critical_path and path")
            nxt = [r for r in rows if r[0][0] != "sblk_encoder_stack_fwd"]
            if nxt:
                line_extra["roofline_critical_path"] = roofline_of(nxt[0])
        fpc = flops_per_clip(T, L)
        path_tf = value / world * fpc / 1e12
        line_extra["roofline"] = roof
        line_extra["path"] = {"flops_per_clip": fpc, "achieved_tflops_per_gpu": path_tf,
                              "frac_of_bf16_peak": path_tf / float(peaks["bf16_tflops"]),
                              "frac_of_bf16_sustained": path_tf / float(peaks.get("bf16_tflops_sustained",
                                                                                  peaks["bf16_tflops"])),
                              "eager_traced_ms_per_step": total_ms}
        line_extra["breakdown"] = breakdown

        # ---- CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same workload --------------
        if world == 1 and not args.no_cpu_baseline:
            O, sd, cores = cpu_oracle_setup(L)
            sample = 4
            xs = host_pool[0][:sample].clone()
            cpu_oracle_step(O, sd, xs)
            reps, t0 = 0, time.perf_counter()
            while True:
                cpu_oracle_step(O, sd, xs)
                reps += 1
                if time.perf_counter() - t0 > args.cpu_seconds or reps >= 200:
                    break
            dt = time.perf_counter() - t0
            line_extra["cpu_baseline"] = {
                "value": sample * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{reps} forwards of {sample} clips x {T} frames (first {sample} clips of the batch), "
                          f"oracle/visual_encoder_oracle.py, fp32, torch {torch.__version__} CPU, {dt:.1f} s"}

    if rank == 0:
        clocks = sampler.summary()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(args, world), host_numa_node_rank0=numa_node,
                           gather=gather_mode if world == 1 or not (pipelined and not args.inline_gather) else
                           f"{gather_mode}, one gather per step inside its event pair: the previous step's output, on a "
                           f"side stream next to this step's replay",
                           pipeline=pipeline_note),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * T * 88 * 88 * 4,
                    "d2h_bytes_per_step": B * T * 512 * 4, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "timing": "wall clock, synchronize on both sides, double-buffered H2D/compute/D2H",
                    "result_checksum": checksum},
            "e2e_u8": e2e_u8,
            "latency": latency,
            "gpu_launches": plan.launches_per_forward * args.steps,
            "gpu_launches_per_step": plan.launches_per_forward,
        }
        line.update(line_extra)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU (BASELINE configs[1]: 32)")
    ap.add_argument("--frames", type=int, default=29)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="time the one-batch plan instead of the two-stage software pipeline")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU output gathering inside the step: one-shot peer-memory kernel (default) or NCCL")
    ap.add_argument("--inline-gather", action="store_true",
                    help="multi-GPU, pipelined plan: gather each step's output in stream after the replay instead of on a "
                         "side stream next to the following step's replay")
    ap.add_argument("--no-u8", action="store_true", help="skip the fused uint8-input end-to-end measurement (e2e_u8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-clips", type=int, default=0, help="clips per reference-arm step (0 = auto)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and args.gpus != world:
        if args.gpus > 1 and world == 1:
            raise SystemExit(f"bench.py --gpus {args.gpus}: launch with torch.distributed.run "
                             f"--nproc-per-node {args.gpus} (one rank per GPU)")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
