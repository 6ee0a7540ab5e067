"""Print the metrics we quote from an .ncu-rep (run here: ncu reads reports without a GPU)."""
import csv, subprocess, sys, json

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']


def summarize(path, index=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[index]
    out = {"kernel": r[hdr.index("Kernel Name")], "launches_in_report": len(data)}
    for w in WANT:
        if w in hdr:
            out[w] = (r[hdr.index(w)] + " " + units[hdr.index(w)]).strip()
    stalls = {}
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0:
                stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = v
        if any(k in h for k in ("uhmma", "utcmma", "tmem", "subpipe_tc")) and "pct_of_peak" in h and ".avg." in h:
            out[h] = r[i]
    tot = sum(stalls.values()) or 1.0
    out["stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(json.dumps(summarize(p), indent=1))
