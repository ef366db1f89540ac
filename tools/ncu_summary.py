"""Condenses `ncu -i <report>.ncu-rep --page raw --csv` into the JSON kept under profiles/ (the .ncu-rep itself stays in
gpurun_out/, which is scratch).  usage: python tools/ncu_summary.py raw.csv out.json "<source line>" "<workload line>" """
import csv, json, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, out, source, workload = sys.argv[1:5]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    keep = KEEP + [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
    kernels = []
    for r in rows[2:]:
        k = {"name": r[col["Kernel Name"]]}
        for n in keep:
            if n in col and r[col[n]] != "":
                k[n] = {"value": float(r[col[n]].replace(",", "")), "unit": units[col[n]]}
        tr = 0.0
        for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tr += k[n]["value"] * TO_BYTES[k[n]["unit"]]
        k["traffic_bytes_per_launch"] = tr
        kernels.append(k)
    json.dump({"source": source, "workload": workload, "kernels": kernels}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
