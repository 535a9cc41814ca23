#!/usr/bin/env python3
"""profiles/traffic.json from the summary of a whole-step `ncu --set full` capture
(profiles/summarize_ncu.py output).  bench.py reads it for `roofline.traffic` and the L1 data-pipe
figures.

usage: python profiles/make_traffic.py 8 STAMP profiles/r2x_step_full_batch8.csv [more.csv ...] > profiles/traffic.json
(several summaries of the same batch size: a kernel takes its figures from the LAST file that captured it)

STAMP = bench.kernel_source_stamp() of the tree the capture was taken from; tools/capture_step.sh writes it on the GPU
box next to the report (gpurun_out/<tag>_stamp.txt).  bench.py leaves the ncu figures out when the tree's stamp differs.
"""
import csv
import json
import sys

GROUPS = {"k_resize_colsum": "resize_colsum", "k_colscan": "colscan", "k_integral_rows": "integral_rows",
          "k_tilt": "tilted", "k_cascade_tiles": "cascade_tiles"}


def one(path, batch):
    rows = list(csv.reader(open(path)))
    head = rows[0][2:]
    metric = {r[0]: (r[1], r[2:]) for r in rows[1:]}

    def val(name, i):
        unit, vals = metric[name]
        try:
            v = float(vals[i].replace(",", ""))
        except ValueError:   # "no data" for a launch that did not collect the metric
            return 0.0
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}.get(unit, 1)
        return v * scale

    out = {}
    for i, h in enumerate(head):
        name = h.split("|")[1].strip()
        key = next((g for k, g in GROUPS.items() if k in name), None)
        if key is None:
            continue
        o = out.setdefault(key, {"dram": 0.0, "ms": 0.0, "pipe_ms": 0.0, "wavefronts": 0.0, "eff_ms": 0.0, "l1": 0.0, "l2": 0.0})
        ms = val("gpu__time_duration.sum", i)
        o["dram"] += val("dram__bytes_read.sum", i) + val("dram__bytes_write.sum", i)
        o["ms"] += ms
        o["pipe_ms"] += ms * val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", i)
        if "SM_B.TriageCompute.l1tex__t_sectors.sum" in metric:   # 32-byte sectors through the L1 tag stage (global) / L2
            o["l1"] += 32 * val("SM_B.TriageCompute.l1tex__t_sectors.sum", i)
        if "lts__t_sectors.sum" in metric:
            o["l2"] += 32 * val("lts__t_sectors.sum", i)
        o["eff_ms"] += ms * val("smsp__thread_inst_executed_per_inst_executed.ratio", i) / 32.0
        # all data-pipe wavefronts (shared + global): the shared-memory count scaled by total pipe % / shared pipe %
        sh, sh_pct = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", i), val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", i)
        if sh_pct > 0:
            o["wavefronts"] += sh * val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", i) / sh_pct
    kernels = {}
    for k, o in out.items():
        kernels[k] = {"dram_bytes_per_frame": round(o["dram"] / batch), "l1_data_pipe_pct": round(o["pipe_ms"] / o["ms"], 1),
                      "l1_bytes_per_frame": round(o["l1"] / batch) or None, "l2_bytes_per_frame": round(o["l2"] / batch) or None,
                      "warp_execution_efficiency": round(o["eff_ms"] / o["ms"], 3),   # active threads per executed instruction / 32
                      "ncu_ms_batch%d" % batch: round(o["ms"], 4)}
        if k == "cascade_tiles":
            kernels[k]["l1_wavefronts_per_frame"] = round(o["wavefronts"] / batch)
            kernels[k]["l1_wavefronts_note"] = ("all L1 data-pipe wavefronts (shared + global) of both tile launches per frame: "
                                                "shared-memory wavefronts x (total pipe % / shared pipe %) from the same capture")
    for k in kernels:
        kernels[k]["capture"] = path
    return kernels


def main():
    batch, stamp, paths = int(sys.argv[1]), sys.argv[2], sys.argv[3:]
    kernels = {}
    for p in paths:
        kernels.update(one(p, batch))
    json.dump({"source": "%s (ncu --set full, bench.py --batch %d; per-frame figures = capture / %d)" % (", ".join(paths), batch, batch),
               "batch": batch, "kernel_source_stamp": stamp, "kernels": kernels}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
