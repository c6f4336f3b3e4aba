"""Turn ncu captures of tools/ncu_target.py into profiles/ncu_counters.json (what bench.py quotes for the rooflines) and
a readable summary per kernel.

    python tools/ncu_counters.py B=gpurun_out/stage_B.ncu-rep [C17=...] [D=...]  [--summary profiles/stage_r2_ncu.md]

The JSON carries a hash of the kernel sources (bench.kernel_source_hash): bench.py ignores counters captured for other
kernels than the ones it is timing.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def rows_of(path):
    if path.endswith('.csv'):  # already exported on the GPU box (tools/make_profiles.sh)
        out = open(path).read()
    else:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


def main():
    args = [a for a in sys.argv[1:] if '=' in a and not a.startswith('--')]
    summary_path = None
    if '--summary' in sys.argv:
        summary_path = sys.argv[sys.argv.index('--summary') + 1]
    commit = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    data = {'kernel_source_hashes': {k: bench.kernel_source_hash(k) for k in bench.KERNEL_SOURCES}, 'commit': commit, 'configs': {}}
    text = [f'# ncu --set full --clock-control none, one un-pipelined chunk per config (tools/ncu_target.py), commit {commit}, '
            f'kernel-source hashes {data["kernel_source_hashes"]}', '']
    for a in args:
        key, path = a.split('=', 1)
        hdr, units, rows = rows_of(path)
        col = {k: hdr.index(k) for k in KEYS if k in hdr}
        name_i = hdr.index('Kernel Name')
        entry = {'score_warp_instructions_per_launch': 0.0}
        text.append(f'## config {key}')
        for r in rows:
            name = r[name_i]
            text.append(f'### {name[:100]}')
            for k, i in col.items():
                text.append(f'  {k:85s} {r[i]:>18s} {units[i]}')
            inst = num(r[col['smsp__inst_executed.sum']])
            if 'hypothesis_kernel_t1' in name or 'frame_prep_kernel' in name:
                entry['score_warp_instructions_per_launch'] += inst
            if 'hypothesis_kernel_t1' in name:
                entry['hypothesis_issue_active_pct'] = num(r[col['smsp__issue_active.avg.pct_of_peak_sustained_active']])
                entry['hypothesis_fma_pipe_pct'] = num(r[col['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']])
                entry['hypothesis_us_under_ncu'] = num(r[col['gpu__time_duration.sum']])
            if 'decode_' in name:
                rd, wr = num(r[col['dram__bytes_read.sum']]), num(r[col['dram__bytes_write.sum']])
                ur, uw = units[col['dram__bytes_read.sum']], units[col['dram__bytes_write.sum']]
                scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
                entry['decode_dram_bytes_per_launch'] = rd * scale.get(ur, 1.0) + wr * scale.get(uw, 1.0)
                entry['decode_us_under_ncu'] = num(r[col['gpu__time_duration.sum']])
            if 'replay_kernel' in name:
                entry['replay_warp_instructions_per_launch'] = inst
                entry['replay_fp64_pipe_pct'] = num(r[col.get('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', col['smsp__inst_executed.sum'])])
        data['configs'][key] = entry
    json.dump(data, open(os.path.join(ROOT, 'profiles', 'ncu_counters.json'), 'w'), indent=1)
    if summary_path:
        open(summary_path, 'w').write('\n'.join(text) + '\n')
    print(json.dumps(data, indent=1))


if __name__ == '__main__':
    main()
