"""Builds profiles/ncu_kernels.json from the raw pages of `ncu --set full` captures
(scripts/gpu_profile_all.sh): per kernel and PEAQ frame the DRAM traffic, the executed FP64
operations, pipe / issue utilisation -- the facts bench.py quotes in `roofline` -- together
with a hash of the kernel sources so that bench.py refuses a capture of other code.

  python scripts/ncu_to_json.py <dir with *.raw.csv> <frames per launch> <tag> [name=kernel ...]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            out[h] = v
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0,
                 "ns": 1e-9, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1.0)
        out[h] = x * scale
    return out


def facts(raw, frames, capture):
    cyc = raw.get("sm__cycles_elapsed.avg", 0.0)

    def per_frame(metric):   # "<op>.sum.per_cycle_elapsed" [inst/cycle] x elapsed cycles / frames
        return raw.get(metric, 0.0) * cyc / frames

    dfma = per_frame("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed")
    dmul = per_frame("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
    dadd = per_frame("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
    rd, wr = raw.get("dram__bytes_read.sum", 0.0), raw.get("dram__bytes_write.sum", 0.0)
    return {
        "capture": capture, "frames_in_launch": frames,
        "duration_ms_under_ncu": raw.get("gpu__time_duration.sum", 0.0) * 1e3,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_frame": (rd + wr) / frames,
        "algorithmic_bytes_per_frame": 16384,
        "fp64_thread_inst_per_frame": {"dfma": dfma, "dmul": dmul, "dadd": dadd},
        "fp64_flop_per_frame": 2 * dfma + dmul + dadd,
        "fp64_pipe_active_pct": raw.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": raw.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": raw.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "warp_inst_per_frame": raw.get("smsp__inst_executed.sum", 0.0) / frames,
        "registers_per_thread": raw.get("launch__registers_per_thread"),
        "lsu_data_pipe_pct": raw.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "shared_wavefronts_per_frame": raw.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0.0) / frames,
        "shared_bank_conflicts_per_frame": raw.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0.0) / frames,
    }


def main():
    import bench
    d, frames, tag = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    doc = {"source_hash": bench.kernel_source_hash(), "tag": tag,
           "how": "ncu --set full --clock-control none, one launch per kernel of scripts/profile_workload.py "
                  "(592 pairs x 10 s stereo); per-frame figures = launch totals / PEAQ frames in the launch",
           "kernels": {}}
    for spec in sys.argv[4:]:
        name, kernel = spec.split("=")
        p = os.path.join(d, name + ".raw.csv")
        if not os.path.exists(p) or sum(1 for _ in open(p)) < 3:
            print("missing", p, file=sys.stderr)
            continue
        doc["kernels"][kernel] = facts(read_raw(p), frames, "profiles/%s_%s_ncu.txt" % (tag, name))
    json.dump(doc, open(os.path.join(ROOT, "profiles", "ncu_kernels.json"), "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
