#!/usr/bin/env python
"""Measure the peaks bench.py divides by, on the box it runs on (one GPU):

  * int8 tensor pipe: back-to-back tcgen05.mma.kind::i8 from resident operands (no loads, no
    epilogue), M128 x N256 (SS), M128 x N128 (SS) and M128 x N128 with A in tensor memory - the
    form the batched Hamming scan issues;
  * POPC pipe: independent 32-bit POPC chains over a full grid;
  * shared memory: 128-bit loads over a full grid (the batched BM25 filter is bound by it);
  * SM clock while the probes run (nvidia-smi).

    python tools/peak_probe.py [out.json]      (default profiles/r2_peaks.json)

The kernels are csrc/probe.cu; nothing here is on the product path."""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from radiant_rag_b200 import _lib  # noqa: E402


def sm_clock_mhz():
    try:
        out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        return float(out[0]), float(out[1])
    except Exception:
        return None, None


def measure() -> dict:
    _lib.init(0)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    v = C.c_double()
    out = {}
    # warm the clocks with a long tensor probe, then sample the clock right after each probe
    lib.rr_probe_i8_mma(0, 20000, C.byref(v), st)
    for mode, name in ((0, "i8_mma_m128n256_ss"), (1, "i8_mma_m128n128_ss"), (2, "i8_mma_m128n128_ts")):
        rc = lib.rr_probe_i8_mma(mode, 40000, C.byref(v), st)
        if rc != 0:
            raise SystemExit(f"rr_probe_i8_mma({mode}): " + _lib.last_error())
        out[name + "_tops"] = v.value / 1e12
    clk, clk_max = sm_clock_mhz()
    rc = lib.rr_probe_popc(20000, C.byref(v), st)
    if rc != 0:
        raise SystemExit("rr_probe_popc: " + _lib.last_error())
    out["popc32_tera_per_s"] = v.value / 1e12
    rc = lib.rr_probe_smem(20000, C.byref(v), st)
    if rc != 0:
        raise SystemExit("rr_probe_smem: " + _lib.last_error())
    out["smem_load_tb_per_s"] = v.value / 1e12
    # random-row gather through the rescoring kernel's bulk-copy ring, arithmetic left out:
    # config 3's shape (1024 x 400 rows of 3 KB out of 1M) and config 2's (256 x 200)
    n, dim = 1_000_000, 768
    rows = torch.empty((n, dim), dtype=torch.float32, device="cuda")
    rows.normal_()
    for name, q, c in (("gather_3kb_rows_1024x400", 1024, 400), ("gather_3kb_rows_256x200", 256, 200)):
        cand = torch.randint(0, n, (q, c), dtype=torch.int64, device="cuda")
        scratch = torch.empty((q, c), dtype=torch.float32, device="cuda")
        rc = lib.rr_probe_gather(rows.data_ptr(), _lib.RR_F32, n, dim, cand.data_ptr(), q, c, scratch.data_ptr(),
                                 C.byref(v), st)
        if rc != 0:
            raise SystemExit("rr_probe_gather: " + _lib.last_error())
        out[name + "_tb_per_s"] = v.value / 1e12
    del rows
    sms = lib.rr_sm_count()
    out["sm_count"] = sms
    out["sm_mhz_after_probe"] = clk
    out["sm_max_mhz"] = clk_max
    if clk_max:
        hz = clk_max * 1e6
        out["per_sm_per_clk_at_max_clock"] = {
            "i8_mac_m128n256_ss": out["i8_mma_m128n256_ss_tops"] * 1e12 / 2 / sms / hz,
            "i8_mac_m128n128_ts": out["i8_mma_m128n128_ts_tops"] * 1e12 / 2 / sms / hz,
            "popc32": out["popc32_tera_per_s"] * 1e12 / sms / hz,
            "smem_bytes": out["smem_load_tb_per_s"] * 1e12 / sms / hz,
        }
    out["gpu"] = torch.cuda.get_device_name(0)
    out["how"] = ("csrc/probe.cu kernels timed with CUDA events (best of 5 after a warm-up); int8 ops = 2 x MAC; "
                  "the tensor probes issue 160000 MMAs per SM from resident operands")
    return out


if __name__ == "__main__":
    res = measure()
    path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r2_peaks.json"
    path.write_text(json.dumps(res, indent=1) + "\n")
    print(json.dumps(res))
