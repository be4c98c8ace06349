"""Per-entry-point CUDA-event timing of the C-ABI calls (used by bench.py for the roofline numbers).

`with KernelTimer() as kt: step()` brackets every liblecb call with CUDA events on the launching
stream and accumulates duration plus algorithmic FLOPs / bytes per entry point.  Instrumentation
only: nothing here is active during the timed throughput region."""
from __future__ import annotations

import torch

from . import _lib


# Set by callers whose launch carries structural zeros (pixel-pair packed stem convs: half the MACs of the
# packed problem are zero weights): algorithmic FLOPs = launched FLOPs * ALGO_FLOP_SCALE.
ALGO_FLOP_SCALE = 1.0


def _sig(name, a):
    if name == "lecb_gemm_bf16":
        return f"M={a[6]} N={a[7]} K={a[8]} flags={a[9]} res={int(bool(a[3]))}"
    if name == "lecb_gemm_bf16_dual":
        return f"M={a[7]} N={a[8]} K={a[1]}+{a[3]} flags={a[9]}"
    if name == "lecb_conv3x3_bf16":
        return f"B={a[4]} H={a[5]} W={a[6]} Cin={a[7]} Cout={a[8]}" + (" +pool" if a[9] & _lib.EPI_AVGPOOL2 else "")
    if name == "lecb_avgpool2x2":
        return f"B={a[2]} H={a[3]} W={a[4]} C={a[5]}"
    return ""


def _work(name, a):
    """(flops, algorithmic HBM bytes) of one call from its C arguments."""
    if name == "lecb_gemm_bf16":
        m, n, k, flags = a[6], a[7], a[8], a[9]
        out_b = 4 if flags & _lib.EPI_OUT_F32 else 2
        res_b = 0 if not a[3] else (4 if flags & _lib.EPI_RES_F32 else 2)
        return 2.0 * m * n * k, 2.0 * (m * k + n * k) + m * n * (out_b + res_b)
    if name == "lecb_gemm_bf16_dual":                 # [A1 | A2] . W^T: both operands and W read once, bf16 out, no residual
        k, m, n = a[1] + a[3], a[7], a[8]
        return 2.0 * m * n * k, 2.0 * (m * k + n * k) + 2.0 * m * n
    if name == "lecb_conv3x3_bf16":
        b, h, w, ci, co = a[4], a[5], a[6], a[7], a[8]
        out_px = b * h * w / (4.0 if a[9] & _lib.EPI_AVGPOOL2 else 1.0)      # fused 2x2 average pool writes a quarter
        return 2.0 * b * h * w * co * 9 * ci * ALGO_FLOP_SCALE, 2.0 * (b * h * w * ci + out_px * co + 9 * ci * co)
    if name == "lecb_avgpool2x2":
        b, h, w, c = a[2], a[3], a[4], a[5]
        return 1.0 * b * h * w * c, 2.0 * b * h * w * c * 1.25
    if name == "lecb_stem_conv1":
        b, h, w, co = a[4], a[5], a[6], a[7]
        return 2.0 * b * (h // 2) * (w // 2) * co * 27, 4.0 * b * 3 * h * w + 2.0 * b * (h // 2) * (w // 2) * co
    if name == "lecb_stem_conv1_u8":
        b, h, w, co = a[6], a[7], a[8], a[9]
        return 2.0 * b * (h // 2) * (w // 2) * co * 27, 1.0 * b * 3 * h * w + 2.0 * b * (h // 2) * (w // 2) * co
    if name == "lecb_head_aggregate":
        ldn, b, p, k, n_txt = a[1], a[7], a[8], a[9], a[10]
        maps = 2 if a[5] else 0
        cols_read = n_txt if a[5] else n_txt - 1          # the positive columns are only read for the pos_map output
        return 20.0 * b * p * k, 4.0 * b * p * (cols_read * k + maps * k) + 4.0 * b * k
    if name == "lecb_l2norm_rows":
        rows, d = a[2], a[3]
        return 3.0 * rows * d, rows * d * ((2 if a[4] else 4) + (2 if a[5] else 4))
    if name == "lecb_asl_fwd_bwd":
        return 25.0 * a[4] * a[5], 12.0 * a[4] * a[5]
    return 0.0, 0.0


class KernelTimer:
    def __init__(self):
        self.records = []          # (name, start_evt, end_evt, flops, bytes)
        self._saved = {}

    def __enter__(self):
        lib = _lib.lib
        for name in _lib.SIGNATURES:
            if name in ("lecb_abi_version", "lecb_last_error", "lecb_launch_count", "lecb_conv3x3_pool_fusable", "lecb_resize_ksize",
                        "lecb_resize_plan", "lecb_window_plan_size", "lecb_window_plan", "lecb_set_pair_gemm", "lecb_set_attn_poly"):      # host-only
                continue
            fn = getattr(lib, name)
            self._saved[name] = fn

            def wrapped(*args, _fn=fn, _name=name):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                stream = torch.cuda.current_stream()
                s.record(stream)
                r = _fn(*args)
                e.record(stream)
                fl, by = _work(_name, args)
                self.records.append((_name, s, e, fl, by, _sig(_name, args)))
                return r

            setattr(lib, name, wrapped)
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(_lib.lib, name, fn)
        torch.cuda.synchronize()
        return False

    def summary(self, steps=1):
        agg = {}
        for name, s, e, fl, by, _ in self.records:
            d = agg.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += by
            d["launches"] += 1
        for d in agg.values():
            for k in d:
                d[k] /= steps
        return agg

    def detail(self, steps=1):
        """Per (entry point, shape) rows sorted by time: the launch list bench.py writes under profiles/."""
        agg = {}
        for name, s, e, fl, by, sig in self.records:
            d = agg.setdefault((name, sig), {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += by
            d["launches"] += 1
        rows = []
        for (name, sig), d in agg.items():
            ms = d["ms"] / steps
            rows.append({"op": name, "shape": sig, "ms_per_step": round(ms, 4), "launches_per_step": d["launches"] / steps,
                         "tflops": round(d["flops"] / steps / max(ms, 1e-9) / 1e9, 1),
                         "gbs": round(d["bytes"] / steps / max(ms, 1e-9) / 1e6, 1)})
        rows.sort(key=lambda r: -r["ms_per_step"])
        return rows
