"""Extract the per-kernel metrics the profile summaries under profiles/ hold from an .ncu-rep
(read on the authoring box: `ncu -i <rep> --page raw --csv`).

    python scripts/ncu_extract.py gpurun_out/prof_x.ncu-rep > profiles/r01_ncu_prof_x.csv
"""
import csv
import subprocess
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "launch__block_size", "launch__cluster_size", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "lts__t_sectors_lookup_miss.sum",
    "lts__t_sectors_srcunit_ltcfabric.sum", "lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum",
    "lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum", "lts__t_sectors_srcunit_tex_lookup_hit.sum",
    "lts__t_sectors_srcunit_tex_lookup_miss.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.max",
    "sm__cycles_elapsed.avg.per_second",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    print(f"kernel,{d['Kernel Name'].split('(')[0]}")
    for k in KEEP:
        if k in d and d[k] != "":
            print(f"{k},{u.get(k, '')},{d[k]}")
