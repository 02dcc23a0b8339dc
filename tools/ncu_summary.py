"""Compact summary of an .ncu-rep: per-kernel headline metrics and the top stall sites from the source page."""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
M = {'dur_us': 'gpu__time_duration.sum', 'grid': 'launch__grid_size', 'dram_rd': 'dram__bytes_read.sum', 'dram_wr': 'dram__bytes_write.sum',
     'dram%': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'tensor%': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'warps%': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'issue%': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'xu%': 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'fma%': 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
     'lsu%': 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l2%': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
     'regs': 'launch__registers_per_thread', 'inst': 'smsp__inst_executed.sum', 'smem_conf': 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
     'st_sect/req': None, 'ld_sect/req': None}
st = [(h, i) for i, h in enumerate(hdr) if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio')]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')][:60]
    out = []
    for k, m in M.items():
        if m and m in hdr:
            v = r[hdr.index(m)]
            try:
                v = f'{float(v):.4g}'
            except ValueError:
                pass
            out.append(f'{k}={v}{units[hdr.index(m)] if k in ("dram_rd", "dram_wr") else ""}')
    def ratio(a, b):
        try:
            return float(r[hdr.index(a)]) / max(float(r[hdr.index(b)]), 1)
        except Exception:
            return float('nan')
    out.append(f"st_sect/req={ratio('l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum'):.1f}")
    out.append(f"ld_sect/req={ratio('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'):.1f}")
    vals = sorted(((float(r[i]), h) for h, i in st), reverse=True)[:6]
    print(f'## {name}\n   ' + ' '.join(out))
    print('   stalls/issue: ' + ', '.join(f'{h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")}={v:.2f}' for v, h in vals))
if topn > 0:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks = src.split('"Kernel Name",')
    seen = set()
    for blk in blocks[1:]:
        lines = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
        kname = lines[0][1][:60]
        if kname in seen:
            continue
        seen.add(kname)
        h = lines[1]
        isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
        sc = [(x, i) for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
        data = []
        for r in lines[2:]:
            try:
                data.append((int(r[isamp]), r))
            except Exception:
                pass
        tot = sum(s for s, _ in data) or 1
        print(f'## source hot spots: {kname} ({tot} samples)')
        for s, r in sorted(data, key=lambda x: -x[0])[:topn]:
            top = sorted(((int(r[i] or 0), x) for x, i in sc), reverse=True)[:2]
            print(f"   {100 * s / tot:5.1f}% ex={r[iex]:>9s} {r[isrc][:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
