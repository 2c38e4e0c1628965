#!/usr/bin/env python
"""Turn the scratch output of tools/gpu_final.sh (gpurun_out/r2/final/) into the tracked files under profiles/:
bench lines, measured peaks, the exact-scan probe, the launch list with its share table, the ncu --set full
summary (step + exact tensor-core scan + the exchange kernels kept from the last capture of those) and the dram
traffic per launch that bench.py reads.  Runs here (no GPU): ncu only reads the reports.

    python tools/collect_profiles.py [gpurun_out/r2/final]"""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "r2" / "final"
PROF = ROOT / "profiles"
TMP = ROOT / "build"


def ncu_csv(rep: Path, out: Path) -> bool:
    if not rep.exists():
        return False
    res = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    out.write_text(res.stdout)
    return res.returncode == 0 and bool(res.stdout.strip())


def table_rows(md: str):
    return [line for line in md.splitlines() if line.startswith("| `")]


def main():
    TMP.mkdir(exist_ok=True)
    for name, dst in (("bench_1.json", "r2_bench_1gpu.json"), ("bench_2.json", "r2_bench_2gpu.json"),
                      ("bench_4.json", "r2_bench_4gpu.json"), ("bench_8.json", "r2_bench_8gpu.json"),
                      ("bench_ref.json", "r2_bench_reference_arm.json"), ("peaks.json", "r2_peaks.json"),
                      ("exact_probe.jsonl", "r2_exact_probe.jsonl"), ("launches_step.csv", "r2_launches_step.csv")):
        if (SRC / name).exists() and (SRC / name).stat().st_size:
            shutil.copy(SRC / name, PROF / dst)
    summ = PROF / "r2_ncu_full_summary.md"
    old = summ.read_text() if summ.exists() else ""
    exch = old[old.index("## Exchange kernels"):] if "## Exchange kernels" in old else ""
    if ncu_csv(SRC / "step_full.ncu-rep", TMP / "step_full.csv"):
        subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(TMP / "step_full.csv"),
                        str(TMP / "step_summary.md"), str(PROF / "r2_traffic.json")], check=True)
        new = (TMP / "step_summary.md").read_text().replace(
            "(third eager step;", "(a full step of the eager loop, both halves on ONE stream so that the launch order is fixed;")
        hdr = [line for line in new.splitlines() if line.startswith("| kernel")][0]
        sep = [line for line in new.splitlines() if line.startswith("|---")][0]
        exact_rows = []
        for rep, tmp in (("exact_tc_scan.ncu-rep", "exact_scan"), ("exact_tc_full.ncu-rep", "exact_full")):
            if ncu_csv(SRC / rep, TMP / f"{tmp}.csv"):
                subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(TMP / f"{tmp}.csv"),
                                str(TMP / f"{tmp}.md"), str(TMP / f"{tmp}.json")], check=True)
                rows = table_rows((TMP / f"{tmp}.md").read_text())
                exact_rows += [r for r in rows if tmp == "exact_scan" or "tc_tf32_scan" not in r]
        if exact_rows:
            new += ("\n## Exact float32 scan on the tensor cores (1M x 768, 64 queries, top-10)\n\n"
                    "`ncu --set full --clock-control none --import-source on -k regex:tc_tf32_scan_kernel --launch-skip 4 -c 2 "
                    "python tools/exact_tc_profile.py` (third call) and `-k regex:'tc_tf32|tx_refine|tc_select_lists|tx_qnorm|tau_keys' "
                    "--launch-skip 10 -c 5` for the small kernels.\n\n" + hdr + "\n" + sep + "\n" + "\n".join(exact_rows) + "\n\n"
                    "`tc_tf32_scan_kernel<0>` (filter pass) reads the 3,072 MB of float32 rows from DRAM once, at the copy bandwidth; "
                    "L2 -> SM traffic is twice that because the 128 x 768 float32 query tile is re-streamed from L2 for every row tile.  "
                    "CUDA-event time of the whole call (sample, tau, filter, refine, select): `r2_exact_probe.jsonl`.\n\n")
        summ.write_text(new + exch)
    # launch list -> share table
    lst = PROF / "r2_launches_step.csv"
    bench = PROF / "r2_bench_1gpu.json"
    if lst.exists() and bench.exists():
        rows = list(csv.reader(line for line in lst.open() if line.startswith('"')))
        hdr, out, tot = rows[0], [], 0.0
        for r in rows[1:]:
            d = dict(zip(hdr, r))
            ns = float(d["Metric Value"])
            tot += ns
            out.append((d["Kernel Name"].split("(")[0], ns, d["Grid Size"], d["Block Size"]))
        b = json.loads(bench.read_text())
        st = b["stages_ms"]
        md = ("# Launch list of one hybrid step (round 2 final, one B200)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none "
              "-k regex:... --launch-skip 35 -c 13 --csv python tools/step_profile.py 3`\n(one full step of the eager loop on the real "
              "config-3 batch; per-launch times are cold-cache and serialised - the SHARE is what must agree with `bench.py`).\n\n"
              "| kernel | us | share of the 13 launches | grid | block |\n|---|---|---|---|---|\n")
        for k, ns, g, bl in out:
            md += f"| `{k}` | {ns / 1e3:.1f} | {ns / tot * 100:.1f} % | {g} | {bl} |\n"
        stage_sum = st["dense_top100"] + st["bm25_top100"] + st["rrf_top10"]
        filt = [ns for k, ns, _, _ in out if "bm25_fast_kernel<0>" in k]
        md += (f"\nSum {tot / 1e6:.3f} ms.  `bench.py` (CUDA events, L2 flushed, kernels timed alone): BM25 filter pass "
               f"{st['bm25']['filter_pass']:.3f} ms = {st['bm25']['filter_pass'] / stage_sum * 100:.1f} % of the sum of the stage times "
               f"({st['dense_top100']:.3f} + {st['bm25_top100']:.3f} + {st['rrf_top10']:.3f} ms); here "
               f"{(filt[0] / tot * 100) if filt else float('nan'):.1f} %.  The graph replay runs the two halves on two streams: "
               f"{b['ms_per_step']:.3f} ms per step.\n")
        (PROF / "r2_launches_step_summary.md").write_text(md)
    sass = subprocess.run([sys.executable, str(ROOT / "tools" / "sass_summary.py")], capture_output=True, text=True)
    if sass.returncode == 0 and sass.stdout.strip():
        (PROF / "r2_sass_opcodes.md").write_text(sass.stdout)
    print("profiles updated from", SRC)


if __name__ == "__main__":
    main()
