"""Summarise an .ncu-rep (ncu --set full capture of one kernel) into profiles/<name>.json + .txt. Run here (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

rep, out_prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
summ = {"report": rep, "launches": []}
want = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__ops_path_tensor_src_fp64.sum": "tensor_fp64_ops",
    "sm__ops_path_tensor_src_fp64.avg.peak_sustained": "tensor_fp64_ops_per_clk_per_sm_peak",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_pct_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_sm",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum": "smem_ld_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
}
name_col = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = {"kernel": r[name_col]}
    for h, u, v in zip(hdr, units, r):
        if h in want:
            d[want[h]] = f"{v} {u}".strip()
    summ["launches"].append(d)

def num(s):
    return float(s.split()[0].replace(",", ""))

def to_bytes(s):
    v, u = s.split()[0], s.split()[1] if len(s.split()) > 1 else "byte"
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return float(v.replace(",", "")) * mult

if summ["launches"]:
    L = summ["launches"][0]
    summ["dram_bytes_per_launch"] = to_bytes(L["dram_read"]) + to_bytes(L["dram_write"])
# stall breakdown from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    idx = {h: i for i, h in enumerate(sh)}
    stall_cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0 for h in stall_cols}
    nsamp = 0
    ninstr = 0
    for r in srows[2:]:
        if r and r[0].startswith("0x") and len(r) == len(sh):
            ninstr += 1
            nsamp += int(r[idx["# Samples"]] or 0)
            for h in stall_cols:
                tot[h] += int(r[idx[h]] or 0)
    summ["sass_instructions"] = ninstr
    summ["stall_samples_pct"] = {h: round(100.0 * v / max(nsamp, 1), 2) for h, v in sorted(tot.items(), key=lambda x: -x[1]) if v}
json.dump(summ, open(out_prefix + ".json", "w"), indent=1)
print(json.dumps(summ, indent=1))
