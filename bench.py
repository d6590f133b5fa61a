#!/usr/bin/env python
"""bench.py -- XLSTM-HVED hot path (ViL-mLSTM block pair + S-MVAE fusion) on B200, fwd+bwd, volumes/s.

One "step" = one pass of the hot path, forward and backward, over a batch of synthetic 128^3 4-modality volumes
(BASELINE.json configs[1], "single ViL mLSTM block microbench on bottleneck tokens (bidirectional, fwd+bwd)",
plus the S-MVAE product-of-experts fusion / sampling / KL of the same volumes, which north_star puts on the path):

  per volume:  bottleneck feature (32,16,16,16) = 4096 tokens of dim 32 -> ViLBlock(TOP_LEFT) -> ViLBlock(BOT_RIGHT)
               (bf16 tensor-core operands, fp32 accumulation), backward to the input and all 28 parameter tensors;
               4 latent levels (348,160 elements, 4 modality posteriors each) -> PoE(all modalities) -> z -> KL,
               backward to the posteriors.
  N > 1:       the batch is sharded (weak scaling, fixed volumes per GPU); the parameter gradients of the two blocks
               are summed with one flat-bucket NCCL all-reduce per step.

Prints ONE JSON line (see the task contract): value = device-resident throughput, e2e = same through the public API
with host (pinned) inputs copied every step and the step's outputs (y, dx, z, KL) copied back, roofline for the dominant
kernel (CUDA-event timed, separate pass; HBM roof on MINIMUM bytes and tensor roof side by side for the cell kernels),
sustained = the same step looped for >= 2 s, config3 = BASELINE configs[2]'s fusion (4 levels x 15 missing-modality subsets
in one launch), conv_path = the conv-path kernels of SURVEY 8f rank 1 alone (K6 - K10: norm + LeakyReLU, spatial gate, depthwise /
1x1x1 / dense 3^3 convolutions at (8, 4, 128^3), rank 0, outside every timed region),
gpu_eager_reference = the reference's own PyTorch classes on the same GPU (the competitor a user has today),
cpu_baseline = the reference's own classes (baseline/_ref, when it travelled with the repo) or else the oracle port of its
algorithm, on the host cores (bounded sample).  `--impl reference` times that CPU path as its own arm.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVELS = ((1, 64), (2, 32), (4, 16), (8, 8))      # (latent channels, edge) per level for a 128^3 volume
DIM, SPATIAL = 32, (16, 16, 16)
S_TOK = 4096
NH, DH = 4, 16
SUBSET_FULL = (0, 1, 2, 3)


# ------------------------------------------------------------------------------------------------ helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (recipe in B200_PROFILING.md).  The timed region of the
    default run is ~25 ms, shorter than one `nvidia-smi` call, so the samples come from NVML directly (pynvml, one query per
    ~2 ms on a polling thread); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None
        self.sm, self.reasons, self.sm_max = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(reasons_fn(self.handle))
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def bind_to_gpu_numa_node(local_rank, world):
    """Pin this rank (and therefore the pinned host buffers it allocates afterwards: first-touch) to the NUMA node of its GPU.
    The node comes from /sys/bus/pci/devices/<bdf>/numa_node when the platform reports it, else from the usual HGX layout
    (GPUs split evenly over the nodes in index order).  Returns a description for the JSON line."""
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        nodes = []
    if len(nodes) < 2:
        return {"numa_nodes": len(nodes), "bound": False}
    node, how = None, "index"
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        v = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if v >= 0:
            node, how = v, "pci"
    except Exception:
        pass
    if node is None:
        n_gpu = max(torch.cuda.device_count(), world)
        node = nodes[min(len(nodes) - 1, local_rank * len(nodes) // n_gpu)]
    cpus = set()
    try:
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        # prefer this node for every later allocation of the process (set_mempolicy(MPOL_PREFERRED)); pinned buffers are
        # allocated after this call
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))      # x86-64 set_mempolicy, MPOL_PREFERRED
        return {"numa_nodes": len(nodes), "bound": True, "node": node, "cpus": len(use), "via": how}
    except Exception as e:
        return {"numa_nodes": len(nodes), "bound": False, "error": f"{type(e).__name__}: {e}"}


def randomise_params(block, seed):
    """utils.init_weights semantics (utils.py:191-215): xavier-normal Linear weights, N(0,1) biases."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in block.named_parameters():
            if name.endswith(("proj_up.weight", "proj_down.weight", "igate.weight", "fgate.weight")):
                std = math.sqrt(2.0 / (p.shape[0] + p.shape[1]))
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif name.endswith(("igate.bias", "fgate.bias")):
                p.copy_(torch.randn(p.shape, generator=g))


def synth_inputs(B, seed, device, pin=False):
    """Seeded synthetic inputs with the statistics observed at the real bottleneck / posteriors."""
    g = torch.Generator().manual_seed(seed)
    mk = (lambda *s: torch.empty(*s).pin_memory()) if pin else (lambda *s: torch.empty(*s))
    x = mk(B, DIM, *SPATIAL).copy_(torch.randn(B, DIM, *SPATIAL, generator=g))
    mus, lvs = [], []
    for C, d in LEVELS:
        mus.append(mk(4, B, C, d, d, d).copy_(1.3 * torch.randn(4, B, C, d, d, d, generator=g)))
        lvs.append(mk(4, B, C, d, d, d).copy_((1.4 * torch.randn(4, B, C, d, d, d, generator=g)).clamp_(-50, 50)))
    if device is not None:
        x = x.to(device)
        mus = [m.to(device) for m in mus]
        lvs = [l.to(device) for l in lvs]
    return x, mus, lvs


# ------------------------------------------------------------------------------------------------ GPU arm
class HotPath:
    """The hot path of B volumes on one GPU, through the package's public API."""

    def __init__(self, B, device, world, two_streams=False):
        import xlstm_hved_b200 as xh
        self.xh, self.B, self.device, self.world = xh, B, device, world
        self.two_streams = two_streams
        self.side = torch.cuda.Stream(device=device) if two_streams else None
        self.blk_f = xh.ViLBlock(DIM, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT)
        self.blk_r = xh.ViLBlock(DIM, xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT)
        randomise_params(self.blk_f, 1)
        randomise_params(self.blk_r, 2)
        self.blk_f.to(device)
        self.blk_r.to(device)
        self.params = list(self.blk_f.parameters()) + list(self.blk_r.parameters())
        # N > 1: ONE flat-bucket all-reduce of the parameter gradients per step (xlstm_hved_b200.dist.FlatGradBucket)
        self.bucket = xh.dist.FlatGradBucket(self.params, average=True) if world > 1 else None
        g = torch.Generator().manual_seed(7)
        self.gy = torch.randn(B, DIM, *SPATIAL, generator=g).to(device)          # upstream gradient of the block output
        self.gz = [torch.randn(1, B, C, d, d, d, generator=g).to(device) for C, d in LEVELS]
        # device-side (5,B,...) expert tensors: slab 0 is the prior (zeros), slabs 1..4 are filled per step.
        # Two slots: the e2e loop copies step k+1's inputs on a copy stream while step k computes.
        self.slots = []
        for _ in range(2):
            self.slots.append(dict(
                mu5=[torch.zeros(5, B, C, d, d, d, device=device) for C, d in LEVELS],
                lv5=[torch.zeros(5, B, C, d, d, d, device=device) for C, d in LEVELS],
                x=torch.zeros(B, DIM, *SPATIAL, device=device),
                ready=torch.cuda.Event(), free=torch.cuda.Event()))
        self.copy_stream = torch.cuda.Stream(device=device)
        self.n_lat = sum(t.numel() for t in self.gz)
        self.kld = torch.zeros(4, device=device)                                 # per-level KL sums of a step
        self.kld_w = torch.tensor([0.5 / (B * C * d ** 3) / 4 for C, d in LEVELS], device=device)
        self.kld_scales = [[0.2 * 0.5 / (B * C * d ** 3) / 4] for C, d in LEVELS]     # d loss / d kld_sum per level (weight 0.2)

    def load(self, x, mus, lvs, slot=0):
        """Copy one batch into the device buffers of `slot` (H2D when the sources are pinned host tensors) on the copy
        stream; step(slot) waits for it, and the next load into the same slot waits for that step to finish."""
        sl = self.slots[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(sl["free"])
            sl["x"].copy_(x, non_blocking=True)
            for l in range(4):
                sl["mu5"][l][1:].copy_(mus[l], non_blocking=True)
                sl["lv5"][l][1:].copy_(lvs[l], non_blocking=True)
            sl["ready"].record(self.copy_stream)

    def capture(self):
        """Record the step of each slot (same public-API calls, static device buffers) into a CUDA graph: the ~30 launches
        of a step otherwise cost more host time than the kernels take on the device."""
        for slot, sl in enumerate(self.slots):
            sl["graph"] = self.xh.GraphedStep(lambda slot=slot: self._body(slot), warmup=1)

    def step(self, slot=0, graphed=False):
        sl = self.slots[slot]
        torch.cuda.current_stream().wait_event(sl["ready"])
        if graphed:
            out = sl["graph"]()
        else:
            out = self._body(slot)
        if self.world > 1:
            self.bucket.reduce()             # gather -> one NCCL all-reduce -> averaged gradients written back into p.grad
        sl["free"].record()
        return out

    def _body(self, slot):
        """One step.  The two halves of the path are independent (the S-MVAE fusion works on the latent posteriors, the ViL
        pair on the bottleneck feature): with two_streams the fusion is enqueued on a side stream, forked from and joined
        back into the main stream, so that its bandwidth-bound kernels can fill SMs the cell kernels leave partly empty
        (captured as a fork / join inside the CUDA graph)."""
        main = torch.cuda.current_stream()
        if self.two_streams:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                fused, kld_total = self._smvae(slot)
        else:
            fused, kld_total = self._smvae(slot)
        y, x = self._vil(slot)
        if self.two_streams:
            main.wait_stream(self.side)
            for t in [kld_total] + [f[2] for f in fused]:
                t.record_stream(main)
        # what the model consumes downstream / upstream of the path: y (-> decoder), z per level (-> decoder), dx (-> encoder
        # backward) and the KL term of the loss
        sl = self.slots[slot]
        sl["outs"] = dict(y=y.detach(), dx=x.grad, z=[f[2] for f in fused], loss=kld_total + y.detach()[0, 0, 0])
        return sl["outs"]["loss"]

    def _smvae(self, slot):
        ops = self.xh.ops
        sl = self.slots[slot]
        mu5, lv5 = sl["mu5"], sl["lv5"]
        # ---- S-MVAE: fusion + sampling + KL in one launch per level, backward in one launch per level
        self.kld.zero_()
        # eps ~ N(0,1) for all four levels (RA_HVED.py:743-744) from torch's generator: one draw over a flat buffer, viewed per level
        flat = torch.empty(self.n_lat, device=self.device).normal_()
        noises, off = [], 0
        for gzl in self.gz:
            noises.append(flat[off:off + gzl.numel()].view(gzl.shape))
            off += gzl.numel()
        levels = list(zip(mu5, lv5))
        # the four latent levels ride in one launch; slab 0 is the model's constant prior (mu = 0, logvar = 0,
        # RA_HVED.py:576-580): declared, not read
        fused = ops.poe_fwd_levels(levels, [SUBSET_FULL], noises=noises, kld_out=self.kld.view(4, 1), standard_prior=True)
        ops.poe_bwd_levels(levels, [SUBSET_FULL], noises=noises, g_zs=self.gz, kld_scales=self.kld_scales, standard_prior=True)
        kld_total = torch.dot(self.kld, self.kld_w)                            # mean KL over the 4 levels (train.py:236-239)
        return fused, kld_total

    def _vil(self, slot):
        sl = self.slots[slot]
        # ---- ViL block pair on the NCDHW feature (token view, no transposed copies), forward + backward
        x = sl["x"].detach().requires_grad_()
        tok = x.reshape(self.B, DIM, -1).transpose(-1, -2)
        y = self.blk_r(self.blk_f(tok))
        for p in self.params:
            p.grad = None
        y.backward(self.gy.reshape(self.B, DIM, -1).transpose(-1, -2))
        return y, x


def run_gpu(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local, world)          # before any pinned allocation
    if world > 1:
        # NCCL writes its version banner (and whatever NCCL_DEBUG asks for) to stdout by default: stdout carries the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    from xlstm_hved_b200 import _lib
    lib = _lib.load_library()
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    hp = HotPath(B, device, world, two_streams=args.streams == 2)
    x, mus, lvs = synth_inputs(B, 1000 + rank, device)
    hp.load(x, mus, lvs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(W):
        hp.step()
    ms_eager = timed(hp.step, K)
    graphed = not args.no_graph
    if graphed:
        hp.load(x, mus, lvs, slot=1)
        torch.cuda.synchronize()
        try:
            hp.capture()
            for _ in range(W):
                hp.step(graphed=True)
        except Exception as e:      # a driver that refuses the capture must not cost the measurement: fall back to eager launches
            print(f"bench.py: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            graphed = False
            torch.cuda.synchronize()
    run_step = (lambda: hp.step(graphed=True)) if graphed else hp.step
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(run_step, K)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host (pinned) inputs copied in every step, result read back every step
    hx, hmus, hlvs = synth_inputs(B, 2000 + rank, None, pin=True)
    h2d = hx.numel() * 4 + sum(m.numel() * 4 for m in hmus) + sum(l.numel() * 4 for l in hlvs)
    # outputs read back every step: y (-> decoder), z of the four levels (-> decoder), dx (-> encoder backward), the loss term
    out_host = dict(y=torch.empty(B, DIM, *SPATIAL).pin_memory(), dx=torch.empty(B, DIM, *SPATIAL).pin_memory(),
                    z=[torch.empty(1, B, C, d, d, d).pin_memory() for C, d in LEVELS], loss=torch.empty(1).pin_memory())
    d2h = sum(t.numel() * 4 for t in (out_host["y"], out_host["dx"], out_host["loss"], *out_host["z"]))
    d2h_stream = torch.cuda.Stream(device=device)
    done = [torch.cuda.Event(), torch.cuda.Event()]

    state = {"k": 0}
    hp.load(hx, hmus, hlvs, slot=0)

    def e2e_step():
        k = state["k"]
        hp.load(hx, hmus, hlvs, slot=(k + 1) & 1)          # next step's inputs stream in while this step computes
        if k >= 2:
            torch.cuda.current_stream().wait_event(done[k & 1])     # step k-2 (same slot) has handed its results over
        hp.step(slot=k & 1, graphed=graphed)
        outs = hp.slots[k & 1]["outs"]
        # results leave on their own stream (full duplex with the next step's H2D); the step is complete for the host when
        # its results have landed
        d2h_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(d2h_stream):
            out_host["y"].copy_(outs["y"].transpose(-1, -2).reshape(B, DIM, *SPATIAL), non_blocking=True)
            out_host["dx"].copy_(outs["dx"], non_blocking=True)
            for hz, dz in zip(out_host["z"], outs["z"]):
                hz.copy_(dz, non_blocking=True)
            out_host["loss"].copy_(outs["loss"].reshape(1), non_blocking=True)
            done[k & 1].record(d2h_stream)
        if k > 0:
            done[(k - 1) & 1].synchronize()                # the previous step's results are on the host
        state["k"] = k + 1

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, K)
    torch.cuda.synchronize()
    hp.load(x, mus, lvs, slot=0)
    torch.cuda.synchronize()

    # ---- sustained: the same device-resident step looped for >= 2 s (the 20-step figure above is a burst of ~25 ms)
    sustained = None
    if not args.no_sustained:
        n_sus = max(K, int(math.ceil(2000.0 / (ms_total / K))))
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        ms_sus = timed(run_step, n_sus)
        c2 = sampler2.stop() if rank == 0 else None
        sustained = {"value": round(world * B * n_sus / (ms_sus * 1e-3), 2), "unit": "volumes/s", "steps": n_sus,
                     "seconds": round(ms_sus * 1e-3, 3), "ms_per_step": round(ms_sus / n_sus, 4), "clocks": c2}

    # ---- config3 (BASELINE configs[2]): the S-MVAE fusion of all 15 missing-modality subsets, 4 latent levels, ONE launch,
    # inference (no sampling), for this rank's B volumes; posteriors larger than L2
    config3 = None
    if True:
        from xlstm_hved_b200 import ops as xops
        sl0 = hp.slots[0]
        levels3 = list(zip(sl0["mu5"], sl0["lv5"]))
        for _ in range(W):
            xops.poe_fwd_levels(levels3, xops.SUBSETS_MODALITIES, standard_prior=True)
        ms_c3 = timed(lambda: xops.poe_fwd_levels(levels3, xops.SUBSETS_MODALITIES, standard_prior=True), K)
        n_lat_rank = sum(C * d ** 3 for C, d in LEVELS) * B
        bytes_c3 = n_lat_rank * (32 + 15 * 8)
        gbs = bytes_c3 / (ms_c3 / K * 1e-3) / 1e9
        peaks0 = load_peaks()
        config3 = {"workload": "configs[2]: PoE fusion under all 15 missing-modality subsets x 4 latent levels, one launch, inference",
                   "value": round(world * B * K / (ms_c3 * 1e-3), 1), "unit": "volumes/s (x 15 subsets each)", "ms_per_launch": round(ms_c3 / K, 4),
                   "roofline": {"kernel": "poe_fwd<15 subsets>", "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks0["hbm"], "unit": "GB/s",
                                "frac": round(gbs / peaks0["hbm"], 4), "algorithmic_bytes_per_launch": bytes_c3,
                                "note": "32 B in (4 modality posteriors; the constant prior is declared, not read) + 15 x 8 B out per latent element"}}

    conv_path = None
    if rank == 0 and not args.no_conv_path:
        conv_path = conv_path_block(device, load_peaks())

    # ---- per-kernel attribution (separate pass, CUDA events on the launching stream)
    nk = lib.xhved_profile_kernel_count()
    names = [lib.xhved_profile_kernel_name(i).decode() for i in range(nk)]
    ms_arr, cnt_arr = (ctypes.c_float * nk)(), (ctypes.c_int * nk)()
    lib.xhved_profile_enable(1)
    lib.xhved_profile_read(ms_arr, cnt_arr, nk)
    for _ in range(K):
        hp.step()
    lib.xhved_profile_read(ms_arr, cnt_arr, nk)
    lib.xhved_profile_enable(0)
    kern = {names[i]: (ms_arr[i], cnt_arr[i]) for i in range(nk) if cnt_arr[i]}
    launches = sum(c for _, c in kern.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    tokens = B * S_TOK
    tokens_heads = tokens * NH
    E = 2 * DIM
    L = 128
    # ALGORITHMIC work per launch: minimum HBM bytes for every kernel and, for the cell kernels, the
    # useful (causal-half) tensor FLOP.  The roof that binds a kernel is the one with the larger bound time.
    DHP, NE = max(DH, 16), max(DH, 16) + 16
    st_tok = DHP * NE * 4 / L                         # one fp32 (or bf16 hi+lo) state per chunk, per token-head
    flops = {
        "mlstm_chunk_grad": tokens_heads * (5 * L * DH + 6 * DH * DH),      # S, dP, dQ, dK, dV causal halves + 3 inter products
        "mlstm_chunk_out": tokens_heads * (2 * L * DH + 2 * DH * DH),       # S, PV causal halves + q.[C|n]
        "mlstm_chunk_state": tokens_heads * (2 * DH * DH),
        "mlstm_chunk_rstate": tokens_heads * (2 * DH * DH),
    }
    n_lat = sum(C * d ** 3 for C, d in LEVELS) * B
    # ViL kernels, per token: MINIMUM bytes in the sense of SURVEY 8d (fp32 x / y / dy / dx, bf16 for every intermediate:
    # K2 + K3 forward = 12 C + 16 E + 8 NH = 1440 B at C = 32); impl_token_bytes = what the kernels move today
    per_token_bytes = {
        "vil_pre_fwd": 4 * DIM + 3 * E * 2 + 8 * 4 + 2 * E * 2,                       # x in; q,k,v tiles, gates, act, z out
        "vil_post_fwd": E * 2 + 2 * E * 2 + 4 * DIM + 4 * DIM,                        # h, act, z, x in; y out
        "vil_post_bwd": 4 * DIM + E * 2 + 2 * E * 2 + E * 2 + 2 * E * 2,              # dy, h, act, z in; dh, d_act, dz out
        "vil_pre_bwd_a": E * 2 + 3 * E * 2 + 8 * 4 + E * 2 + 2 * E * 2,               # xm, dq,dk,dv, dgates, d_act in; dconv, dxmv out
        "vil_pre_bwd_b": 4 * DIM + 4 * DIM + 3 * E * 2 + 4 * DIM,                     # x, dy, dconv, dxmv, dz in; dx out
    }
    ACT_B = 2                                                                         # bytes per element of the saved act / z / xm
    impl_token_bytes = {
        "vil_pre_fwd": 4 * DIM + 3 * E * 2 + 8 * 4 + 3 * E * ACT_B,                   # ... + xm saved for the backward
        "vil_post_fwd": E * 2 + 2 * E * ACT_B + 4 * DIM + 4 * DIM,
        "vil_post_bwd": 4 * DIM + E * 2 + 2 * E * ACT_B + E * 2 + 2 * E * 2,
        "vil_pre_bwd_a": E * ACT_B + 3 * E * 2 + 8 * 4 + E * 2 + 2 * E * 2,
        "vil_pre_bwd_b": 4 * DIM + 4 * DIM + 3 * E * 2 + 4 * DIM,
    }
    # cell kernels, per token-head: MINIMUM bytes in the sense of SURVEY 8d -- bf16 q,k,v (dh) in, bf16 h (dq,dk,dv) out, fp32
    # gates / stabiliser / normaliser; carried states, hi/lo pairs and fp32 gradient rows are implementation traffic and do
    # NOT count (impl_bytes below keeps them for reference)
    per_tokenhead_bytes = {
        "mlstm_chunk_state": 4 * DHP + 8,                                             # k, v tiles, i/f gates in (the chunk state is an intermediate)
        "mlstm_chunk_out": 6 * DHP + 8 + 2 * DHP + 8,                                 # q,k,v, gates in; h, m, den out
        "mlstm_chunk_rstate": 6 * DHP + 12,                                           # q, dh, h tiles, f, m, den in
        "mlstm_chunk_grad": 10 * DHP + 16 + 6 * DHP + 8,                              # q,k,v,h,dh, gates, m, den in; bf16 dq,dk,dv, di, dc out = 280 B at dhp 16
    }
    impl_tokenhead_bytes = {
        "mlstm_chunk_state": 4 * DHP + 8 + st_tok,
        "mlstm_state_scan": 2 * st_tok,
        "mlstm_chunk_out": 6 * DHP + 8 + st_tok + 2 * DHP + 8,
        "mlstm_chunk_rstate": 6 * DHP + 12 + st_tok,
        "mlstm_chunk_grad": 10 * DHP + 16 + 2 * st_tok + 6 * DHP + 8,                 # + hi/lo states C and R
        "mlstm_gate_finish": 12,
    }
    bytes_per_launch = {k: v * tokens for k, v in per_token_bytes.items()}
    bytes_per_launch.update({k: v * tokens_heads for k, v in per_tokenhead_bytes.items()})
    impl_bytes_per_launch = {k: v * tokens_heads for k, v in impl_tokenhead_bytes.items()}
    impl_bytes_per_launch.update({k: v * tokens for k, v in impl_token_bytes.items()})
    # PoE per latent element: 4 x (mu, logvar) in (the constant prior is never read) + noise; mu^, logvar^, z out;
    # backward: the same 32 B + noise + g_z in, 32 B of gradients out.  One launch per step covers the 4 latent levels.
    bytes_per_step = {"poe_fwd": n_lat * (32 + 4 + 12), "poe_bwd": n_lat * (32 + 4 + 4 + 32)}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if not os.path.exists(traffic_file):
        traffic_file = os.path.join(ROOT, "profiles", "traffic_r01.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}

    def roofline_of(name):
        ms, cnt = kern[name]
        if name in bytes_per_step:
            ach = bytes_per_step[name] / (ms / K * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm"], 4), "traffic": traffic.get(name), "algorithmic_bytes_per_step": bytes_per_step[name],
                    "ms_per_step": round(ms / K, 5), "peak_source": peaks["src"], "note": "one launch per step covers the four latent levels"}
        if name not in bytes_per_launch:
            return {"kernel": name, "bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                    "avg_launch_ms": round(ms / cnt, 5),
                    "note": "moves intermediates only (no algorithmic bytes of its own): pure latency"}
        sec = ms / cnt * 1e-3
        nbytes, nflop = bytes_per_launch[name], flops.get(name, 0)
        hbm = {"kernel": name, "bound": "hbm", "achieved": round(nbytes / sec / 1e9, 1), "peak": peaks["hbm"], "unit": "GB/s",
               "frac": round(nbytes / sec / 1e9 / peaks["hbm"], 4), "traffic": traffic.get(name), "algorithmic_bytes_per_launch": int(nbytes),
               "avg_launch_ms": round(ms / cnt, 5), "peak_source": peaks["src"]}
        if name in impl_bytes_per_launch:
            hbm["implementation_bytes_per_launch"] = int(impl_bytes_per_launch[name])
            hbm["note_bytes"] = "algorithmic = minimum bytes (bf16 operands / intermediates / results, fp32 gates and x / y); implementation bytes add carried states, the saved xm and whatever is still fp32 between kernels"
        if not nflop:
            return hbm
        # cell kernels: both roofs side by side -- HBM on minimum bytes (frac) and bf16 tensor peak on useful FLOP (tensor_frac)
        tf = nflop / sec / 1e12
        hbm["tensor_frac"] = round(tf / peaks["tf"], 5)
        hbm["tensor_achieved_tflops"] = round(tf, 2)
        hbm["tensor_peak_tflops"] = peaks["tf"]
        hbm["algorithmic_flop_per_launch"] = nflop
        binding = "hbm" if nflop / (peaks["tf"] * 1e12) <= nbytes / (peaks["hbm"] * 1e9) else "tensor"
        hbm["note"] = "arithmetic intensity %.0f FLOP/B vs ridge %.0f: the binding roof is %s; both fractions are reported" % (
            nflop / nbytes, peaks["tf"] * 1e12 / (peaks["hbm"] * 1e9), binding)
        if binding == "tensor":
            hbm.update({"bound": "tensor", "achieved": round(tf, 3), "peak": peaks["tf"], "unit": "TFLOP/s", "hbm_frac": hbm["frac"],
                        "frac": round(tf / peaks["tf"], 5)})
        return hbm

    order = sorted(kern.items(), key=lambda kv: -kv[1][0])
    roof = roofline_of(order[0][0])
    secondary = {k: {kk: vv for kk, vv in roofline_of(k).items() if kk in ("bound", "achieved", "unit", "frac", "tensor_frac", "avg_launch_ms", "ms_per_step")}
                 for k, _ in order[1:]}
    shares = {k: round(v[0] / sum(m for m, _ in kern.values()), 4) for k, v in order}

    # "mLSTM kernel % of bf16 tensor peak" (second half of BASELINE.json's metric): the cell kernels alone, at the shipped
    # head dim and at the scaled widths of SURVEY.md 8d (f_maps 16 / 32 -> DH 64 / 128)
    cell = None
    if world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_cell
        cell = {}
        for cfg in ((32, 4, 4096, 16), (16, 4, 4096, 64), (16, 4, 4096, 128)):
            r = bench_cell.run(*cfg, iters=5)
            ko = r["kernels"]["mlstm_chunk_out"]
            cell[f"DH{cfg[3]}"] = {"shape_B_NH_S_DH": list(cfg), "chunk_out_ms": ko["ms"], "chunk_out_useful_tflops": ko["useful_tflops"],
                                   "chunk_out_executed_tflops": ko["executed_tflops"],
                                   "chunk_out_executed_frac_of_burst_peak": ko["executed_frac_of_bf16_burst_peak"],
                                   "fwd_ms": r["fwd_ms"], "bwd_ms": r["bwd_ms"],
                                   "chunk_out_flop_per_byte": ko["flop_per_byte"], "ridge_flop_per_byte": r["ridge_flop_per_byte"],
                                   "chunk_out_hbm_frac_on_required_bytes": ko["hbm_frac_on_required_bytes"],
                                   "reference_form_equiv_tflops_fwd": r["reference_form_equiv_tflops_fwd"],
                                   "reference_form_equiv_frac_of_bf16_burst_peak_fwd": r["reference_form_equiv_frac_of_bf16_burst_peak_fwd"],
                                   "reference_form_equiv_frac_of_bf16_burst_peak_fwd_bwd": r.get("reference_form_equiv_frac_of_bf16_burst_peak_fwd_bwd")}
            kg = r["kernels"].get("mlstm_chunk_grad")
            if kg:
                cell[f"DH{cfg[3]}"].update({"chunk_grad_ms": kg["ms"], "chunk_grad_useful_tflops": kg["useful_tflops"],
                                            "chunk_grad_executed_tflops": kg["executed_tflops"],
                                            "chunk_grad_executed_frac_of_burst_peak": kg["executed_frac_of_bf16_burst_peak"],
                                            "chunk_grad_hbm_frac_on_required_bytes": kg["hbm_frac_on_required_bytes"]})
        cell["note"] = ("useful = causal-half FLOP of the chunkwise algorithm, executed = full 128x128 tiles on the tensor pipe; the chunkwise "
                        "form (L = 128) has 20-90 executed FLOP per byte it must move, below the ridge of the measured peaks, so HBM is the "
                        "binding roof at every head dim (hbm_frac_on_required_bytes); reference_form_equiv = the FLOP the reference's O(S^2) "
                        "parallel form needs for the same problem / our time")
    eager_ref = gpu_eager_reference(device) if world == 1 and not args.no_eager_ref else None
    torch.cuda.empty_cache()
    cpu = cpu_reference_arm(steps=args.cpu_steps, warmup=1) if world == 1 and not args.no_cpu else None
    vols = world * B * K
    line = {
        "metric": "128^3 4-modality volumes/sec fwd+bwd (ViL-mLSTM block pair + S-MVAE fusion hot path)",
        "value": round(vols / (ms_total * 1e-3), 2), "unit": "volumes/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 tensor-core operands, fp32 accumulate/gates/PoE", "data": "synthetic (seeded, bottleneck statistics)",
        "config": {"workload": "configs[1]: bidirectional ViLBlock pair on bottleneck tokens (dim 32, S=4096, NH=4, DH=16) fwd+bwd "
                               "+ S-MVAE PoE/reparam/KL fwd+bwd over the 4 latent levels of the same volumes",
                   "volumes_per_gpu": B, "global_batch": world * B, "tokens_per_volume": S_TOK,
                   "latent_elements_per_volume": sum(C * d ** 3 for C, d in LEVELS),
                   "l2_policy": f"inputs larger than L2 (PoE posteriors {h2d / 2**20:.0f} MiB per step > 126 MiB)",
                   "launch": "CUDA graph replay of the step captured through the public API" if graphed else "eager (one Python call per op)",
                   "streams": args.streams,
                   "eager_ms_per_step": round(ms_eager / K, 4),
                   "collective": "1 flat-bucket NCCL all-reduce (xlstm_hved_b200.dist.FlatGradBucket: gather, all-reduce, average, write "
                                 "back) of the 28 ViL parameter gradients per step" if world > 1 else "none"},
        "e2e": {"value": round(vols / (ms_e2e * 1e-3), 2), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / K, 4), "numa": numa,
                "note": "every step copies its inputs from pinned host memory (copy stream, double-buffered against the previous "
                        "step's compute) and copies back what the model consumes around the path: y and dx (B,32,16,16,16), z of the four "
                        "latent levels and the loss term (own stream, full duplex with the next step's inputs).  The gradients w.r.t. the "
                        "posteriors (d_mu, d_logvar: 8 x the z bytes) and the 28 parameter gradients stay on the device, where the encoder "
                        "backward / the optimizer consume them.  PCIe-bound"},
        "sustained": sustained, "config3": config3, "conv_path": conv_path, "gpu_eager_reference": eager_ref,
        "gpu_launches": launches, "kernel_ms_per_step": {k: round(v[0] / K, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])},
        "kernel_ms_per_step_total": round(sum(v[0] for v in kern.values()) / K, 4), "kernel_time_share": shares, "roofline": roof, "roofline_secondary": secondary,
        "clocks": clocks, "mlstm_cell_tensor_peak": cell,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU / reference arm
def conv_path_block(device, peaks, n_vol=8, iters=10, warm=3):
    """SURVEY 8f rank 1 (the callers either side of the path): the conv path's normalisation / gate / depthwise kernels (K6, K7, K8)
    alone, at the first encoder level's shape for `n_vol` volumes -- (n_vol, 4, 128^3) fp32 = 268 MB per tensor, larger than L2.
    CUDA events on the launching stream, rank-local (no collective).  HBM fractions are on algorithmic bytes."""
    from xlstm_hved_b200 import ops as xops
    g = torch.Generator(device=device).manual_seed(3)
    shape = (n_vol, 4, 128, 128, 128)
    x = torch.rand(shape, device=device, generator=g)
    dy = torch.randn(shape, device=device, generator=g)
    elems = x.numel()

    def t(fn):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / iters

    def hbm(ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / peaks["hbm"], 4), "algorithmic_bytes": nbytes}

    out = {"workload": f"conv path of XLSTM_HVED at the first encoder level, {n_vol} volumes: tensors ({n_vol}, 4, 128^3) fp32 (> L2)",
           "peak": peaks["hbm"], "unit": "GB/s"}
    y, mean, rstd = xops.norm_act_fwd(x, slope=0.01)
    out["norm_act_fwd"] = hbm(t(lambda: xops.norm_act_fwd(x, slope=0.01)), 8 * elems)
    out["norm_act_bwd"] = hbm(t(lambda: xops.norm_act_bwd(x, dy, mean, rstd, slope=0.01)), 12 * elems)
    out["pytorch_instance_norm_leaky_relu_fwd_ms"] = round(t(lambda: torch.nn.functional.leaky_relu_(torch.nn.functional.instance_norm(x), 0.01)), 3)
    w3 = torch.randn(4, 27, device=device, generator=g) * 0.2
    out["dwconv3_fwd"] = hbm(t(lambda: xops.dwconv3_fwd(x, w3)), 8 * elems)
    out["dwconv3_bwd"] = hbm(t(lambda: xops.dwconv3_bwd(x, w3, dy)), 16 * elems)          # dgrad: dy in, dx out; wgrad: x, dy in
    w1, b1 = torch.randn(4, 4, device=device, generator=g) * 0.5, torch.randn(4, device=device, generator=g)
    out["pwconv_fwd"] = hbm(t(lambda: xops.pwconv_fwd(x, w1, b1)), 8 * elems)
    out["pwconv_bwd"] = hbm(t(lambda: xops.pwconv_bwd(x, w1, dy, want_db=True)), 16 * elems)       # dgrad: dy in, dx out; wgrad: x, dy in
    xh16, dyh16 = x.half(), dy.half()
    out["pwconv_fwd_fp16"] = hbm(t(lambda: xops.pwconv_fwd(xh16, w1, b1)), 4 * elems)
    out["pwconv_bwd_fp16"] = hbm(t(lambda: xops.pwconv_bwd(xh16, w1, dyh16, want_db=True)), 8 * elems)
    conv16 = torch.nn.Conv3d(4, 4, 1).to(device).half()
    xg16 = xh16[:1].clone().requires_grad_()

    def stock16():
        conv16.zero_grad(set_to_none=True)
        torch.autograd.grad(conv16(xg16), [xg16] + list(conv16.parameters()), dyh16[:1])
    out["pytorch_conv1x1x1_fp16_fwd_bwd_ms_ONE_volume"] = round(t(stock16), 3)
    del xh16, dyh16
    wd = torch.randn(4, 4, 3, 3, 3, device=device, generator=g) * 0.1
    fl3 = 2.0 * 27 * 4 * elems                                                               # 27 taps x 4 input channels per output element
    ms_f, ms_b = t(lambda: xops.conv3_fwd(x, wd, b1)), t(lambda: xops.conv3_bwd(x, wd, dy, want_db=True))
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    out["conv3_fwd"] = {**hbm(ms_f, 8 * elems), "TFLOP/s": round(fl3 / (ms_f * 1e-3) / 1e12, 2),
                        "frac_of_nominal_fp32": round(fl3 / (ms_f * 1e-3) / 1e12 / fp32_nominal, 4)}
    out["conv3_bwd"] = {**hbm(ms_b, 16 * elems), "TFLOP/s": round(2 * fl3 / (ms_b * 1e-3) / 1e12, 2),
                        "frac_of_nominal_fp32": round(2 * fl3 / (ms_b * 1e-3) / 1e12 / fp32_nominal, 4)}
    w7 = torch.randn(4, 343, device=device, generator=g) * 0.02
    dgate = dy[:, :1].contiguous()
    gate = xops.gate7_fwd(x, w7)
    fl = 2.0 * 343 * elems                                                                   # one FMA per tap and input element
    ms_f, ms_b = t(lambda: xops.gate7_fwd(x, w7)), t(lambda: xops.gate7_bwd(x, w7, gate, dgate))
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12                                              # nominal CUDA-core fp32 (no measured figure)
    out["gate7_fwd"] = {"ms": round(ms_f, 4), "TFLOP/s": round(fl / (ms_f * 1e-3) / 1e12, 2), "frac_of_nominal_fp32": round(fl / (ms_f * 1e-3) / 1e12 / fp32_peak, 4)}
    out["gate7_bwd"] = {"ms": round(ms_b, 4), "TFLOP/s": round(2 * fl / (ms_b * 1e-3) / 1e12, 2),
                        "frac_of_nominal_fp32": round(2 * fl / (ms_b * 1e-3) / 1e12 / fp32_peak, 4)}
    out["nominal_fp32_TFLOP/s"] = round(fp32_peak, 1)
    return out


def cpu_hot_path_one_volume(params_f, params_r, x, mus, lvs, gy, gzs):
    """The reference's algorithm for the same hot path on the CPU (oracle port: fp32, O(S^2) parallel cell with the
    reference's op sequence, PoE as buildingblocks.py:853-866), forward + backward through autograd."""
    from oracle import restate
    x = x.clone().requires_grad_()
    tok = x.reshape(1, DIM, -1).transpose(-1, -2)
    y = restate.vil_block(tok, params_f, reverse=False, cell=restate.mlstm_parallel_reference_cost)
    y = restate.vil_block(y, params_r, reverse=True, cell=restate.mlstm_parallel_reference_cost)
    loss = (y * gy.reshape(1, DIM, -1).transpose(-1, -2)).sum()
    for l in range(4):
        mm = mus[l].clone().requires_grad_()
        ll = lvs[l].clone().requires_grad_()
        mu5 = torch.cat([torch.zeros_like(mm[:1]), mm], 0)
        lv5 = torch.cat([torch.zeros_like(ll[:1]), restate.clip_logvar(ll)], 0)
        a, b = restate.poe(mu5, lv5, SUBSET_FULL)
        z = restate.reparametrize(a, b, torch.randn_like(a))
        loss = loss + (z * gzs[l][0]).sum() + 0.2 * restate.kl_to_prior(a, b) / 4
    loss.backward()
    return float(loss)


def reference_hot_path_one_volume(ns, blk_f, blk_r, x, mus, lvs, gy, gzs):
    """The same hot path through the UNMODIFIED reference classes on the CPU: vision_lstm.ViLBlock (which builds the S x S
    decay matrix, vision_lstm.py:48-130), buildingblocks.ProductOfExperts, RA_HVED.reparametrize / clip, loss.KL_divergence."""
    x = x.clone().requires_grad_()
    tok = x.reshape(1, DIM, -1).transpose(-1, -2)
    y = blk_r(blk_f(tok))
    loss = (y * gy.reshape(1, DIM, -1).transpose(-1, -2)).sum()
    experts = ns.buildingblocks.ProductOfExperts()
    for l in range(4):
        mm = mus[l].clone().requires_grad_()
        ll = lvs[l].clone().requires_grad_()
        mu5 = torch.cat([torch.zeros_like(mm[:1]), mm], 0)                      # RA_HVED.py:576-580
        lv5 = torch.cat([torch.zeros_like(ll[:1]), ns.RA_HVED.clip(ll)], 0)
        a, b = experts(mu5, lv5, SUBSET_FULL)
        z = ns.RA_HVED.reparametrize(a, b, False)
        loss = loss + (z * gzs[l][0]).sum() + 0.2 * ns.loss.KL_divergence(a, b) / 4
    loss.backward()
    return float(loss)


def gpu_eager_reference(device, iters=3):
    """The competitor a user of the reference has on the same GPU: the reference's OWN classes (baseline/_ref, unmodified) in
    PyTorch eager on the B200 -- vision_lstm.ViLBlock pair (O(S^2) parallel cell), buildingblocks.ProductOfExperts,
    RA_HVED.reparametrize / clip, loss.KL_divergence -- fwd + bwd, one volume per step as the reference trains (train.py:50).
    A baseline leg: nothing of it is on the product path."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "RA_HVED.py")):
        return {"unavailable": "no reference tree travelled with the repo (baseline/_ref)"}
    try:
        os.environ["XHVED_REFERENCE"] = ref_dir
        from oracle import ref_loader
        import xlstm_hved_b200 as xh
        ns = ref_loader.load_reference()
        vl = ns.vision_lstm
        blocks = []
        for seed, md, rd in ((1, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT, vl.SequenceTraversal.ROWWISE_FROM_TOP_LEFT),
                             (2, xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT, vl.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT)):
            mirror = xh.ViLBlock(DIM, md)
            randomise_params(mirror, seed)
            rb = vl.ViLBlock(dim=DIM, direction=rd)
            rb.load_state_dict(mirror.state_dict(), strict=True)
            blocks.append(rb.to(device))
        x, mus, lvs = synth_inputs(1, 1000, device)
        g = torch.Generator().manual_seed(7)
        gy = torch.randn(1, DIM, *SPATIAL, generator=g).to(device)
        gzs = [torch.randn(1, 1, C, d, d, d, generator=g).to(device) for C, d in LEVELS]
        step = lambda: reference_hot_path_one_volume(ns, blocks[0], blocks[1], x, mus, lvs, gy, gzs)
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"value": round(1e3 / ms, 2), "unit": "volumes/s", "ms_per_volume": round(ms, 3), "volumes_per_step": 1, "iters": iters,
                "what": "reference's own vision_lstm.ViLBlock pair + ProductOfExperts / reparametrize / clip / KL_divergence, fp32 eager, "
                        "fwd+bwd through autograd, on this GPU (includes the loss read-back the helper does per volume)",
                "peak_mem_gb": round(torch.cuda.max_memory_allocated(device) / 2 ** 30, 2)}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def cpu_reference_arm(steps, warmup):
    """CPU timing of the path on the host cores: the reference's own classes when a reference tree travelled with the repo
    (baseline/_ref, kind "reference"), else the oracle port of the same algorithm (kind "port")."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    import xlstm_hved_b200 as xh
    keys = xh.ops.VIL_PARAM_KEYS
    x, mus, lvs = synth_inputs(1, 1000, None)
    g = torch.Generator().manual_seed(7)
    gy = torch.randn(1, DIM, *SPATIAL, generator=g)
    gzs = [torch.randn(1, 1, C, d, d, d, generator=g) for C, d in LEVELS]
    mirrors = []
    for seed, direction in ((1, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT), (2, xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT)):
        blk = xh.ViLBlock(DIM, direction)
        randomise_params(blk, seed)
        mirrors.append(blk)
    step, kind, what = None, "port", "oracle port of the reference's O(S^2) parallel cell + PoE"
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isfile(os.path.join(ref_dir, "RA_HVED.py")):
        try:
            os.environ["XHVED_REFERENCE"] = ref_dir
            from oracle import ref_loader
            ns = ref_loader.load_reference()
            vl = ns.vision_lstm
            real = []
            for blk, direction in zip(mirrors, (vl.SequenceTraversal.ROWWISE_FROM_TOP_LEFT, vl.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT)):
                rb = vl.ViLBlock(dim=DIM, direction=direction)
                rb.load_state_dict(blk.state_dict(), strict=True)
                real.append(rb)
            step = lambda: reference_hot_path_one_volume(ns, real[0], real[1], x, mus, lvs, gy, gzs)
            step()                                                   # proves the reference tree is usable before it is timed
            kind, what = "reference", "the reference's own vision_lstm.ViLBlock / ProductOfExperts / reparametrize / KL_divergence"
        except Exception as e:
            print(f"bench.py: reference tree unusable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
            step = None
    if step is None:
        blocks = [{k: blk.state_dict()[k].clone().requires_grad_() for k in keys} for blk in mirrors]
        step = lambda: cpu_hot_path_one_volume(blocks[0], blocks[1], x, mus, lvs, gy, gzs)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": round(steps / dt, 4), "unit": "volumes/s", "cores": cores, "threads": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} volume(s), one per step, same hot path fwd+bwd ({what}, fp32, torch CPU), {dt / steps:.2f} s/volume",
            "seconds_per_volume": round(dt / steps, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, args.warmup
    K = min(K, 6)                       # bounded: one volume per step, a few seconds each on the host cores
    cpu = cpu_reference_arm(steps=K, warmup=min(W, 1))
    line = {
        "impl": "reference",
        "metric": "128^3 4-modality volumes/sec fwd+bwd (ViL-mLSTM block pair + S-MVAE fusion hot path)",
        "value": cpu["value"], "unit": "volumes/s", "n_gpus": world, "steps": K, "warmup": min(W, 1),
        "ms_per_step": round(1e3 / cpu["value"], 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32 (reference precision)", "data": "synthetic (seeded, bottleneck statistics)",
        "config": {"workload": "configs[1]: bidirectional ViLBlock pair on bottleneck tokens (dim 32, S=4096, NH=4, DH=16) fwd+bwd "
                               "+ S-MVAE PoE/reparam/KL fwd+bwd over the 4 latent levels; bounded sample: 1 volume per step",
                   "volumes_per_step": 1},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="volumes per GPU per step")
    ap.add_argument("--cpu-steps", type=int, default=2, help="volumes timed for cpu_baseline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained loop")
    ap.add_argument("--streams", type=int, default=1, choices=[1, 2], help="2: S-MVAE fusion on a side stream next to the ViL pair")
    ap.add_argument("--no-eager-ref", action="store_true", help="skip timing the reference's PyTorch classes on the GPU")
    ap.add_argument("--no-conv-path", action="store_true", help="skip the K6 / K7 / K8 block (conv-path kernels alone)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
        run_gpu(args)


if __name__ == "__main__":
    main()
