#!/usr/bin/env python
"""mLSTM cell microbench: achieved tensor throughput of the chunkwise tcgen05 kernels at the reference head dim (16)
and at the scaled widths SURVEY.md 8d asks for (DH = 32 / 64 / 128), forward and backward, B*NH*S large.
Prints one JSON object per configuration; `useful` counts causal-half FLOP, `executed` counts the
full 128x128 tiles the tensor pipe actually runs."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlstm_hved_b200 import _lib, ops  # noqa: E402


def run(B, NH, S, DH, iters=10):
    lib = _lib.load_library()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v = [0.3 * torch.randn(B, NH, S, DH, device=dev, generator=g) for _ in range(3)]
    ig = torch.randn(B, NH, S, 1, device=dev, generator=g)
    fg = 2.0 + torch.randn(B, NH, S, 1, device=dev, generator=g)
    buf = ops.mlstm_pack_inputs(q, k, v, ig, fg)
    dh_tiles = torch.randn_like(buf.h.float()).to(torch.bfloat16)
    bwd = True
    for _ in range(2):
        ops.mlstm_fwd_tiles(buf)
        if bwd:
            ops.mlstm_bwd_tiles(buf, dh_tiles)
    nk = lib.xhved_profile_kernel_count()
    names = [lib.xhved_profile_kernel_name(i).decode() for i in range(nk)]
    ms, cnt = (ctypes.c_float * nk)(), (ctypes.c_int * nk)()
    lib.xhved_profile_enable(1)
    lib.xhved_profile_read(ms, cnt, nk)
    for _ in range(iters):
        ops.mlstm_fwd_tiles(buf)
        if bwd:
            ops.mlstm_bwd_tiles(buf, dh_tiles)
    lib.xhved_profile_read(ms, cnt, nk)
    lib.xhved_profile_enable(0)
    t = {names[i]: ms[i] / cnt[i] for i in range(nk) if cnt[i]}
    th = B * NH * S
    L, D = 128, buf.dhp
    useful = {"mlstm_chunk_out": th * (2 * L * DH + 2 * DH * DH), "mlstm_chunk_grad": th * (5 * L * DH + 6 * DH * DH),
              "mlstm_chunk_state": th * 2 * DH * DH, "mlstm_chunk_rstate": th * 2 * DH * DH}
    executed = {"mlstm_chunk_out": th * (4 * L * D + 2 * 2 * D * (D + 16)), "mlstm_chunk_grad": th * (2 * L * D + 2 * L * (D + 16) + 6 * L * D + 3 * 2 * 2 * D * (D + 16)),
                # one pass of the 128-row window (hi and lo halves ride in it) for dhp <= 64, two passes at dhp = 128
                "mlstm_chunk_state": th * (1 if D <= 64 else 2) * 2 * 128 * (D + 16),
                "mlstm_chunk_rstate": th * (1 if D <= 64 else 2) * 2 * 128 * (D + 16)}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    # HBM bytes the chunkwise algorithm has to move per launch (per token-head): bf16 operand / result tiles, fp32 gates and
    # row scalars, and the carried inter-chunk state (fp32-class precision: a bf16 hi/lo pair, dhp x (dhp+16) per chunk of 128)
    st = D * (D + 16) * 4 / L
    need_bytes = {"mlstm_chunk_out": th * (6 * D + 8 + st + 2 * D + 8), "mlstm_chunk_grad": th * (10 * D + 16 + 2 * st + 6 * D + 8),
                  "mlstm_chunk_state": th * (4 * D + 8 + st), "mlstm_chunk_rstate": th * (6 * D + 12 + st)}
    out = {"B": B, "NH": NH, "S": S, "DH": DH, "dhp": D, "kernels": {}}
    fwd_ms = sum(t.get(k_, 0) for k_ in ("mlstm_chunk_state", "mlstm_state_scan", "mlstm_chunk_out"))
    bwd_ms = sum(t.get(k_, 0) for k_ in ("mlstm_chunk_rstate", "mlstm_chunk_grad", "mlstm_gate_finish")) + (t.get("mlstm_state_scan", 0) if bwd else 0)
    for k_, v_ in t.items():
        e = {"ms": round(v_, 4)}
        if k_ in useful:
            e["useful_tflops"] = round(useful[k_] / (v_ * 1e-3) / 1e12, 2)
            e["executed_tflops"] = round(executed[k_] / (v_ * 1e-3) / 1e12, 2)
            e["executed_frac_of_bf16_burst_peak"] = round(e["executed_tflops"] / peaks["bf16_tflops"], 4)
            # which roof binds: executed FLOP per byte the algorithm must move, against the ridge of the measured peaks
            e["flop_per_byte"] = round(executed[k_] / need_bytes[k_], 1)
            e["hbm_frac_on_required_bytes"] = round(need_bytes[k_] / (v_ * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)
        out["kernels"][k_] = e
    out["fwd_ms"], out["bwd_ms"] = round(fwd_ms, 4), round(bwd_ms, 4) if bwd else None
    out["fwd_useful_tflops"] = round((useful["mlstm_chunk_out"] + useful["mlstm_chunk_state"]) / (fwd_ms * 1e-3) / 1e12, 2)
    out["ridge_flop_per_byte"] = round(peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9), 1)
    # third numerator: the FLOP the REFERENCE's O(S^2) parallel form spends on the same problem (causal halves of
    # QK^T and C.V: F = 2 B NH S^2 DH forward, 3.5 F forward + backward) divided by OUR time -- what the chunkwise algorithm is
    # worth in units of the reference's algorithm; it exceeds the executed rate by ~S / (2 L)
    F = 2.0 * B * NH * S * S * DH
    out["reference_form_equiv_tflops_fwd"] = round(F / (fwd_ms * 1e-3) / 1e12, 1)
    out["reference_form_equiv_frac_of_bf16_burst_peak_fwd"] = round(F / (fwd_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], 3)
    if bwd:
        out["reference_form_equiv_tflops_fwd_bwd"] = round(3.5 * F / ((fwd_ms + bwd_ms) * 1e-3) / 1e12, 1)
        out["reference_form_equiv_frac_of_bf16_burst_peak_fwd_bwd"] = round(3.5 * F / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peaks["bf16_tflops"], 3)
    return out


if __name__ == "__main__":
    for cfg in ((32, 4, 4096, 16), (8, 4, 32768, 16), (32, 4, 4096, 32), (16, 4, 4096, 64), (2, 4, 32768, 64), (16, 4, 4096, 128)):
        print(json.dumps(run(*cfg)))
