#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` reports: DRAM bytes (read + write) per launch of the profiled kernels, keyed by
the tags bench.py uses for its per-kernel rooflines.
   python tools/ncu_traffic.py gpurun_out/r01_v52_kernels.ncu-rep gpurun_out/r01_v47_gemm.ncu-rep > profiles/traffic.json
The first launch of every kernel in a report is skipped when there is more than one (cold instruction / descriptor caches)."""
import csv
import io
import json
import re
import subprocess
import sys

TAGS = [  # (kernel-name regex, bench tag)
    (r"lstm_clk_fwd_kernel|lstm_seq_fwd_kernel", "lstm_seq_fwd"), (r"lstm_clk_bwd_kernel|lstm_seq_bwd_kernel", "lstm_seq_bwd"),
    (r"atten_fwd_tpar_kernel", "atten_fwd"), (r"atten_bwd_tpar_kernel", "atten_bwd"), (r"dec_atten_tma_kernel", "dec_step_fused"),
    (r"dec_cell_kernel", "dec_cell"), (r"ce_fwd_bwd", "ce_fwd_bwd"),
    (r"argmax_filter_kernel", "dec_argmax_filter"), (r"argmax_refine_kernel", "dec_argmax_refine"), (r"argmax_finalize_kernel", "dec_argmax"),
    (r"gemm_tc_kernel<128, 2, 5, 0, 0, 0, 2>", "dec_vocab_gemm1"),
]
DECODE_SPLIT_ORDER = ["dec_gate_gemm", "dec_qr_gemm"]   # the two 3xTF32 contractions of a decode step, in launch order (reports named *decode*)
GEMM_ORDER = ["gemm_vocab_fwd", "gemm_vocab_dx", "gemm_vocab_dw", "gemm_gates_in", "gemm_lstm_dw", "gemm_lstm_dx"]   # tools/prof_gemm.py, 2 launches each
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rd[2:]:
        if len(r) < len(hdr):
            continue
        val = lambda m: float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
        yield r[col["Kernel Name"]], val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")


def main(reps):
    acc = {}
    for rep in reps:
        gemm_i = 0
        split_i = 0
        seen = {}
        for name, rd, wr, us in rows(rep):
            if "decode" in rep and re.search(r"gemm_tc_kernel<\d+, 4, \d+, 0, 0, 1", name):
                tag = DECODE_SPLIT_ORDER[split_i % 2]
                first = split_i < 2
                split_i += 1
            elif "gemm_tc_kernel" in name and "gemm" in rep and "decode" not in rep:
                tag = GEMM_ORDER[gemm_i // 2] if gemm_i // 2 < len(GEMM_ORDER) else None
                first = gemm_i % 2 == 0
                gemm_i += 1
            else:
                tag = next((t for rx, t in TAGS if re.search(rx, name)), None)
                first = tag not in seen
                seen[tag] = True
            if tag is None:
                continue
            acc.setdefault(tag, {"all": [], "warm": [], "report": rep.split("/")[-1]})
            acc[tag]["all"].append((rd, wr, us))
            if not first:
                acc[tag]["warm"].append((rd, wr, us))
    out = {"_how": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full reports (tools/ncu_traffic.py); "
                   "first launch of a kernel skipped when a later one exists; under ncu every launch starts with cold L2"}
    for tag, v in acc.items():
        use = v["warm"] or v["all"]
        n = len(use)
        out[tag] = {"dram_bytes_per_launch": sum(a + b for a, b, _ in use) / n, "dram_read_bytes": sum(a for a, _, _ in use) / n,
                    "dram_write_bytes": sum(b for _, b, _ in use) / n, "us_under_ncu": sum(c for _, _, c in use) / n, "launches": n,
                    "report": v["report"]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
