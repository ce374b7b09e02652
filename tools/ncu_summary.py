"""Summarise an .ncu-rep (read here, no GPU):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name
Writes <out>.metrics.csv (selected raw metrics) and <out>.hot_sass.csv (instructions with the most stall samples)."""
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg ', 'launch__registers_per_thread ', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct',
        'sm__inst_executed_pipe_tensor_subpipe_dmma', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum ', 'sm__warps_active.avg.pct',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum.per_second', 'lts__t_sectors.sum ',
        'lts__t_sectors_lookup_hit.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum ', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'lrc__average_xbar2gpc_sectors_op_read.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pcsamp_warps_issue_stalled']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out + '.metrics.csv', 'w', newline='') as f:
        wr = csv.writer(f)
        wr.writerow(['launch', 'metric', 'unit', 'value'])
        for li, vals in enumerate(rows[2:]):
            for h, u, v in zip(hdr, units, vals):
                if h in ('Kernel Name',) or (any(k in h + ' ' for k in KEYS) and 'not_issued' not in h):
                    wr.writerow([li, h, u, v])
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    body = [r for r in srows[2:] if len(r) > 5 and (r[2] or '0').isdigit()]      # (a report with several launches repeats its header rows)
    tot = sum(int(r[2] or 0) for r in body) or 1
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][2] or 0))[:60]
    with open(out + '.hot_sass.csv', 'w', newline='') as f:
        wr = csv.writer(f)
        wr.writerow(['kernel', srows[0][1] if srows else ''])
        wr.writerow(['sass_index', 'stall_samples', 'pct_of_samples', 'instructions_executed', 'sass'])
        for i in sorted(ranked):
            wr.writerow([i, body[i][2], '%.2f' % (100.0 * int(body[i][2] or 0) / tot), body[i][5], re.sub(r'\s+', ' ', body[i][1]).strip()])
    print('wrote', out + '.metrics.csv', out + '.hot_sass.csv')


if __name__ == '__main__':
    main()
