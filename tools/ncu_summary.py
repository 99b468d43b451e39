"""Summarise an ncu report: `python tools/ncu_summary.py file.ncu-rep [regex]` (needs ncu on PATH)."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = re.compile(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|lts__t_bytes.sum$|lts__t_sectors.sum$|lts__t_sector_hit_rate.pct|"
                  r"l1tex__t_sector_hit_rate.pct|l1tex__t_(sectors|requests)_pipe_lsu_mem_global_op_ld.sum$|throughput.avg.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active.avg.pct_of_peak_sustained_active|launch__(registers_per_thread|occupancy_limit|grid_size|block_size|shared_mem_per_block_dynamic|waves)|"
                  r"sm__inst_executed.sum$|sm__inst_executed_pipe_(lsu|alu|fma|fmaheavy|xu|tensor|uniform).*sum$|smsp__inst_executed.avg.per_cycle_active|"
                  r"pipe_tensor.*pct|smsp__issue_active.avg.pct|l1tex__data_pipe_lsu_wavefronts(_mem_shared|_mem_lg)?.sum$|"
                  r"smsp__average_warp.*issue_stalled.*_per_warp_active.pct|smsp__pcsamp_warps_issue_stalled_[a-z_]+$|smsp__pcsamp_sample_buffer_full|sm__cycles_elapsed.max|l1tex__data_bank_conflicts_pipe_lsu.sum|achieved_occupancy|"
                  r"smsp__thread_inst_executed_per_inst_executed.ratio|lts__t_sectors_srcunit_tex_op_read.sum$|l1tex__m_xbar2l1tex_read_sectors.sum$")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:100])
    for i, h in enumerate(hdr):
        if KEEP.search(h) and (pat is None or pat.search(h)):
            print(f"  {h:95s} {units[i]:14s} {r[i]}")
