"""dev tool: turn .ncu-rep captures into the committed evidence under profiles/:
     python tools/ncu_summarize.py gpurun_out/attn_r2.ncu-rep profiles/r2_attn_ncu.md [shape-tag]
   writes a markdown table (one row per captured launch) and merges dram bytes / launch into profiles/ncu_traffic.json
   (key: "<entry point or kernel>/<shape-tag>"), which bench.py reads for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
    ("launch__registers_per_thread", "regs / thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{os.path.basename(rep)}` (`ncu --set full --clock-control none`; per launch)\n"]
    traffic = {}
    for r in data:
        name = r[idx["Kernel Name"]]
        lines.append(f"\n## `{name[:110]}`\n\n| metric | value |\n|---|---|")
        rd = wr = None
        for key, label in WANT:
            if key in idx:
                v, u = r[idx[key]], units[idx[key]]
                lines.append(f"| {label} (`{key}`) | {v} {u} |")
                if key == "dram__bytes_read.sum":
                    rd = (float(v.replace(",", "")), u)
                if key == "dram__bytes_write.sum":
                    wr = (float(v.replace(",", "")), u)
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1)
            short = name.split("(")[0].split("<")[0].split("::")[-1].replace("void ", "")
            traffic[f"{short}/{tag}" if tag else short] = tot
            lines.append(f"| **dram traffic / launch** | {tot / 1e6:.1f} MB |")
    open(out_md, "w").write("\n".join(lines) + "\n")
    tj = os.path.join(os.path.dirname(os.path.abspath(out_md)), "ncu_traffic.json")
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    cur.update(traffic)
    json.dump(cur, open(tj, "w"), indent=1, sort_keys=True)
    print("wrote", out_md, "and", tj, traffic)


if __name__ == "__main__":
    main()
