#!/usr/bin/env python
"""BASELINE configs[3] and configs[2] at model level: the reference XLSTM_HVED with the hot path patched onto the sm_100a
kernels, on 1 / 2 / 4 / 8 GPUs (one process per GPU: `torchrun --nproc-per-node N tools/bench_train.py`, or plain python).

  train  (configs[3], train.py:207-268 without the GAN branch): model.train(), fp16 autocast +
         GradScaler as the reference, two forwards ([14] and one random subset, utils.subset_idx), loss = dice_f + dice_m +
         0.2 MSE(recon_m, x) + 0.2 mean_l KLD_l, backward, ONE flat-bucket NCCL all-reduce of all 422,588 gradients
         (xlstm_hved_b200.dist.FlatGradBucket; replaces nn.DataParallel's reduce, train.py:148-151), Adam(1e-4, wd 1e-5).
  infer  (configs[2], test.py:78-102 / evaluation.py:306-342): model.eval(), for each of the 15 missing-modality subsets
         zero the missing modalities and run model(x, [idx], valid=True); volumes sharded over the ranks (weak scaling).
  grads  stock-vs-patched gradient parity of one training step (same init, inputs, subset and noise), reduced over ranks.

Prints one JSON line per mode on rank 0 (device-timed with CUDA events, max over ranks).  Needs a reference tree
(baseline/_ref travels to the GPU box); the reference model is the harness here, the patched ViL / S-MVAE path the product.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xlstm_hved_b200 as xh                      # noqa: E402
from oracle import ref_loader                     # noqa: E402  (loads the reference harness; nothing of oracle/ is timed)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def train_step(ns, model, opt, scaler, bucket, x, mask, subset_index_list, amp=True):
    """One optimisation step of train.py:207-268 (generator side, GAN term left out: SURVEY 8d config 4)."""
    dice_loss, l2 = train_step.dice, torch.nn.functional.mse_loss
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp), quiet():
        f_out, _, f_rec = model(x, [14], recon=True)
        m_out, (mu, logvar), m_rec = model(x, subset_index_list, recon=True)
        m_rec = torch.cat(m_rec, dim=1)
        kld = sum(ns.loss.compute_KLD(mu[l], logvar[l], subset_index_list) for l in range(len(mu))) / len(mu)
        loss = dice_loss(f_out, mask) + dice_loss(m_out, mask) + 0.2 * l2(m_rec, x) + 0.2 * kld
    opt.zero_grad(set_to_none=True)
    scaler.scale(loss).backward()
    if bucket is not None:
        bucket.reduce()                              # one collective for every gradient of the model
    scaler.step(opt)
    scaler.update()
    return loss.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="train,infer", help="comma list of train, infer, grads")
    ap.add_argument("--batch", type=int, default=1, help="volumes per GPU per step (reference default 1, train.py:50)")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--stock", action="store_true", help="leave the reference's PyTorch ViL / PoE path in place (A/B)")
    ap.add_argument("--no-amp", action="store_true")
    ap.add_argument("--profile", action="store_true", help="train mode: print a torch.profiler kernel table of one step to stderr")
    ap.add_argument("--truth", action="store_true", help="grads mode: also run the stock model in fp64 and compare both fp32 paths with it")
    ap.add_argument("--no-tf32", action="store_true", help="cuDNN / cuBLAS in plain fp32 (PyTorch's default lets convolutions use TF32)")
    args = ap.parse_args()
    if args.no_tf32:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if ref_loader.find_reference() is None:
        if rank == 0:
            print(json.dumps({"unavailable": "no reference tree (baseline/_ref)"}))
        return
    ns = ref_loader.load_reference()
    train_step.dice = ns.loss.DiceLoss()
    model = ref_loader.build_model(f_maps=4, seed=1).to(dev)         # same seed on every rank: identical replicas
    patched = None if args.stock else xh.patch_model(model)
    B, S = args.batch, args.size
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.rand(B, 4, S, S, S, device=dev, generator=g)
    mask = (torch.rand(B, 3, S, S, S, device=dev, generator=g) > 0.5).float()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    common = {"n_gpus": world, "volumes_per_gpu": B, "size": S, "path": "stock reference PyTorch" if args.stock else "patched (sm_100a kernels)",
              "patched": patched, "model_params": n_params}
    modes = args.mode.split(",")

    if "grads" in modes:
        # identical init / inputs / subset / noise through the stock and the patched path; gradients averaged over the ranks
        def grads_of(tag):
            model.train()
            if tag == "patched":
                xh.patch_model(model)
            else:
                xh.unpatch_model(model)
            model.zero_grad(set_to_none=True)
            torch.manual_seed(77 + rank)
            with quiet():
                f_out, _, _ = model(x, [14], recon=True)
                m_out, (mu, logvar), m_rec = model(x, [5], recon=True)
                kld = sum(ns.loss.compute_KLD(mu[l], logvar[l], [5]) for l in range(len(mu))) / len(mu)
                loss = train_step.dice(f_out, mask) + train_step.dice(m_out, mask) + \
                    0.2 * torch.nn.functional.mse_loss(torch.cat(m_rec, 1), x) + 0.2 * kld
            loss.backward()
            bucket = xh.dist.FlatGradBucket(model.parameters(), average=True)
            flat = bucket.reduce().clone()
            none = sum(1 for p in bucket.params if p.grad is None)
            return flat, loss.item(), none
        gs, ls, n0 = grads_of("stock")
        gp, lp, n1 = grads_of("patched")
        truth = None
        if args.truth:
            # the same step through the STOCK model in fp64 (fp32 noise draws, widened): which of the two fp32 paths is nearer?
            import copy
            xh.unpatch_model(model)
            m64 = copy.deepcopy(model).double().train()
            ra = sys.modules["RA_HVED"]
            orig_rep = ra.reparametrize

            def rep64(mu, logvar, valid=False):
                if valid:
                    return mu
                std = logvar.mul(0.5).exp()
                return torch.empty(std.size(), dtype=torch.float32, device=std.device).normal_().double() * std + mu
            ra.reparametrize = rep64
            try:
                torch.manual_seed(77 + rank)
                with quiet():
                    f_out, _, _ = m64(x.double(), [14], recon=True)
                    m_out, (mu, logvar), m_rec = m64(x.double(), [5], recon=True)
                    kld = sum(ns.loss.compute_KLD(mu[l], logvar[l], [5]) for l in range(len(mu))) / len(mu)
                    loss = train_step.dice(f_out, mask.double()) + train_step.dice(m_out, mask.double()) + \
                        0.2 * torch.nn.functional.mse_loss(torch.cat(m_rec, 1), x.double()) + 0.2 * kld
                loss.backward()
            finally:
                ra.reparametrize = orig_rep
            g64 = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m64.parameters()])
            if world > 1:
                dist.all_reduce(g64)
                g64 /= world
            rl = lambda a: ((a.double() - g64).norm() / g64.norm()).item()
            truth = {"rel_l2_stock_fp32_vs_fp64": rl(gs), "rel_l2_patched_vs_fp64": rl(gp), "loss_fp64": loss.item()}
            del m64
        if args.stock:
            xh.unpatch_model(model)
        rel = ((gp.double() - gs.double()).norm() / gs.double().norm()).item()
        cos = torch.nn.functional.cosine_similarity(gp.double(), gs.double(), dim=0).item()
        if rank == 0:
            print(json.dumps({"mode": "grads", **common, "grad_elements": gs.numel(), "rel_l2_patched_vs_stock": rel, "cosine": cos,
                              "loss_stock": ls, "loss_patched": lp, "params_without_grad": [n0, n1], "fp64_truth": truth,
                              "tf32_convs": torch.backends.cudnn.allow_tf32,
                              "note": "one training step (subsets [14] and [5], seeded noise), fp32, all gradients of the model in one flat "
                                      "bucket averaged over the ranks; bf16 tensor-core operands in the patched ViL cell"}))

    if "train" in modes:
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)          # train.py:177
        scaler = torch.amp.GradScaler("cuda", enabled=not args.no_amp)
        bucket = xh.dist.FlatGradBucket(model.parameters(), average=True) if world > 1 else None
        rng = np.random.RandomState(7)

        def step():
            idx = [int(rng.choice(range(4, 10)))]                                        # utils.subset_idx([2]): one random pair
            return train_step(ns, model, opt, scaler, bucket, x, mask, idx, amp=not args.no_amp)
        for _ in range(args.warmup):
            step()
        if args.profile and rank == 0:
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
                step()
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=80), file=sys.stderr)
            print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=34,
                                                                     max_shapes_column_width=110), file=sys.stderr)
        ms = timed(step, args.steps)
        if rank == 0:
            print(json.dumps({"mode": "train", **common, "steps": args.steps, "ms_per_step": round(ms / args.steps, 2),
                              "volumes_per_s": round(world * B * args.steps / (ms * 1e-3), 3), "amp": not args.no_amp,
                              "allreduce_bytes": (n_params + len(list(model.parameters()))) * 4 if world > 1 else 0,
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)}))

    if "infer" in modes:
        model.eval()
        subsets = ns.RA_HVED.SUBSETS_MODALITIES

        def all_subsets():
            with torch.no_grad(), quiet():
                for idx, present in enumerate(subsets):
                    xm = x.clone()
                    for m in range(4):
                        if m not in present:
                            xm[:, m] = 0                                                 # evaluation.py:306-307
                    model(xm, [idx], valid=True)
        for _ in range(max(1, args.warmup - 1)):
            all_subsets()
        ms = timed(all_subsets, args.steps)
        if rank == 0:
            print(json.dumps({"mode": "infer", **common, "steps": args.steps, "ms_per_volume_all_15_subsets": round(ms / args.steps / B, 2),
                              "volumes_per_s": round(world * B * args.steps / (ms * 1e-3), 3),
                              "forwards_per_s": round(15 * world * B * args.steps / (ms * 1e-3), 2)}))
        # the same evaluation as ONE batch of 15 masked copies per volume (xlstm_hved_b200.driver)
        batched = lambda: xh.all_subsets_forward(model, x)
        batched()
        torch.cuda.reset_peak_memory_stats(dev)
        ms_b = timed(batched, args.steps)
        if rank == 0:
            print(json.dumps({"mode": "infer_batched", **common, "steps": args.steps, "ms_per_volume_all_15_subsets": round(ms_b / args.steps / B, 2),
                              "volumes_per_s": round(world * B * args.steps / (ms_b * 1e-3), 3), "speedup_vs_15_forwards": round(ms / ms_b, 3),
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)}))
        if not args.stock:
            # ... and replayed as one CUDA graph (xh.GraphedSubsetsForward)
            gfwd = xh.GraphedSubsetsForward(model)
            gfwd(x)
            ms_g = timed(lambda: gfwd(x), args.steps)
            one = lambda: gfwd(x, subsets=[14])
            one()
            ms_1 = timed(one, args.steps)
            if rank == 0:
                print(json.dumps({"mode": "infer_graphed", **common, "steps": args.steps,
                                  "ms_per_volume_all_15_subsets": round(ms_g / args.steps / B, 2),
                                  "volumes_per_s": round(world * B * args.steps / (ms_g * 1e-3), 3),
                                  "ms_per_single_forward_graphed": round(ms_1 / args.steps / B, 2)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
