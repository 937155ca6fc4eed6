"""Summarise an .ncu-rep (first kernel): key raw metrics + opcode mix + hottest SASS by stall samples.
Usage: python tools/ncu_summary.py file.ncu-rep [n_hot]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h in ('Kernel Name', 'Grid Size', 'Block Size') or h.startswith('smsp__average_warps_issue_stalled'):
            if h.startswith('smsp__average_warps_issue_stalled') and float(v) < 0.05:
                continue
            print('%s [%s] = %s' % (h, u, v))
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    hdr, data = rows[1], rows[2:]
    iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    data = [r for r in data if len(r) > max(iS, iE, iSm) and r[iE].isdigit()]  # (a multi-kernel report repeats its header)
    tot = sum(int(r[iE]) for r in data)
    tots = sum(int(r[iSm]) for r in data)
    byop, sm = collections.Counter(), collections.Counter()
    for r in data:
        t = r[iS].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        byop[op] += int(r[iE])
        sm[op] += int(r[iSm])
    print('# opcode mix: total warp instructions %d, SASS lines %d' % (tot, len(data)))
    for op, c in byop.most_common(22):
        print('%-10s %6.2f%% inst  %6.2f%% samples' % (op, 100.0 * c / tot, 100.0 * sm[op] / max(tots, 1)))
    print('# hottest SASS by stall samples')
    order = sorted(range(len(data)), key=lambda i: -int(data[i][iSm]))[:n_hot]
    for i in order:
        print('%5d %6.2f%%  %s' % (i, 100.0 * int(data[i][iSm]) / max(tots, 1), data[i][iS].strip()))


if __name__ == '__main__':
    main()
