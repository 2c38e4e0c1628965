#!/usr/bin/env python
"""Summarise an `ncu --set full` report of one hybrid step (tools/step_profile.py) into the tracked
profile files: a markdown table per kernel and the dram traffic per launch keyed by kernel + shape
(profiles/r2_traffic.json, read by bench.py for `roofline.traffic`).

    ncu -i gpurun_out/r2/step_full2.ncu-rep --page raw --csv > /tmp/full2.csv
    python tools/ncu_summary.py /tmp/full2.csv profiles/r2_ncu_full_summary.md profiles/r2_traffic.json
"""
import csv
import json
import sys

SHAPES = {  # kernel substring -> (traffic key, shape key) at BASELINE config 3 on one GPU
    "tc_i8_search_kernel<0>": ("tc_i8_search_kernel_filter", "1000000x768x1024"),
    "bm25_fast_kernel<0>": ("bm25_fast_kernel_filter", "1000000x50000x1024"),
    "rescore_ring_kernel<0, 0>": ("rescore_ring_kernel_f32", "1024x400x768"),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    src, md_path, js_path = sys.argv[1:4]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, k, scale=True):
        try:
            v = float(r[ix[k]].replace(",", ""))
        except (KeyError, ValueError):
            return None
        return v * UNIT.get(units[ix[k]], 1.0) if scale else v

    cols = [("ms", "gpu__time_duration.sum", True), ("dram read MB", "dram__bytes_read.sum", True),
            ("dram write MB", "dram__bytes_write.sum", True), ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
            ("IPC", "sm__inst_executed.avg.per_cycle_elapsed", False),
            ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
            ("LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
            ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
            ("warp instr (M)", "smsp__inst_executed.sum", False), ("regs", "launch__registers_per_thread", False),
            ("grid x block", None, False)]
    out = ["# ncu --set full of one hybrid step on the real config-3 batch (round 2, one B200)", "",
           "`ncu --set full --clock-control none --import-source on -k regex:... --launch-skip 35 -c 13 python tools/step_profile.py 3`",
           "(third eager step; 1M docs x 768 dims, 1024 queries x 8 tokens, dense top-100 of 400 candidates, BM25 top-100, RRF top-10).",
           "Durations are ncu's (cold cache, serialised); the CUDA-event times of the same kernels are in `r2_bench_1gpu.json`.", "",
           "| kernel | " + " | ".join(c[0] for c in cols) + " |", "|---|" + "---|" * len(cols)]
    traffic = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        cells = []
        for label, key, scale in cols:
            if key is None:
                cells.append(f"{r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}")
                continue
            v = val(r, key, scale)
            if v is None:
                cells.append("")
            elif "MB" in label:
                cells.append(f"{v / 1e6:.1f}")
            elif label == "warp instr (M)":
                cells.append(f"{v / 1e6:.1f}")
            elif label == "ms":
                cells.append(f"{v:.4f}")
            else:
                cells.append(f"{v:.1f}")
        short = name.split("(")[0].replace("void ", "").replace("rr::", "")
        out.append(f"| `{short}` | " + " | ".join(cells) + " |")
        for sub, (tk, shape) in SHAPES.items():
            if sub in name:
                traffic.setdefault(tk, {})[shape] = int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    open(md_path, "w").write("\n".join(out) + "\n")
    open(js_path, "w").write(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(out))
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
