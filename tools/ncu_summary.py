"""Condense `ncu --page details --csv` of a multi-kernel capture (tools/ncu_kernels.sh) into one row per kernel:
    python tools/ncu_summary.py gpurun_out/ncu_kernels_r02.csv > profiles/r02_ncu_kernels_summary.csv
First launch of each kernel; the columns the north star asks for (occupancy, active threads per warp = warp divergence,
branch efficiency, IPC, DRAM and L2 figures)."""
import csv
import re
import sys

want = {
    "Duration": "dur_us", "Registers Per Thread": "regs", "Theoretical Occupancy": "occ_theo%", "Achieved Occupancy": "occ_ach%",
    "Avg. Active Threads Per Warp": "act_thr/warp", "Avg. Not Predicated Off Threads Per Warp": "npo_thr/warp",
    "Branch Efficiency": "branch_eff%", "Avg. Divergent Branches": "div_branches", "Executed Ipc Active": "ipc",
    "Warp Cycles Per Issued Instruction": "cyc/inst", "DRAM Throughput": "dram%", "L2 Hit Rate": "l2hit%",
    "Grid Size": "grid", "Block Size": "block",
}
rows = {}
order = []
with open(sys.argv[1]) as f:
    for r in csv.DictReader(l for l in f if not l.startswith("==")):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("scpr::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        key = (name, r["ID"])
        if name not in [k[0] for k in order]:
            order.append(key)
        if key not in order:
            continue  # later launches of a kernel already seen
        m = r["Metric Name"]
        if m in want and want[m] not in rows.setdefault(key, {}):
            v = r["Metric Value"].replace(",", "")
            if m == "Duration":
                u = r["Metric Unit"]
                v = float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                v = f"{v:.1f}"
            rows[key][want[m]] = v
cols = ["grid", "block", "dur_us", "regs", "occ_theo%", "occ_ach%", "act_thr/warp", "npo_thr/warp", "branch_eff%", "div_branches", "ipc", "cyc/inst", "dram%", "l2hit%"]
print("kernel," + ",".join(cols))
for key in order:
    print(key[0] + "," + ",".join(str(rows.get(key, {}).get(c, "")) for c in cols))
