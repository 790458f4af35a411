"""Summaries of the ncu captures made by tools/ncu_step.sh (read here, on the CPU box):

    python tools/ncu_summarize.py            # writes profiles/r2_ncu_launch_summary.txt, r2_ncu_full_summary.txt, gemm_traffic.json

* launch list (`--metrics gpu__time_duration.sum`): per-kernel-class share of the step;
* `--set full` captures: per launch duration, DRAM bytes, tensor-pipe active %, DRAM throughput %, achieved occupancy,
  registers; for GEMM launches the algorithmic bytes of the call (operands + output + epilogue addend) next to the DRAM traffic.
"""
import csv
import gzip
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')


def rows_of(path):
    with open(path, newline='') as fh:
        lines = [ln for ln in fh if not ln.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    units = None
    out = []
    for r in rd:
        if len(r) != len(hdr):
            continue
        if units is None and r[0] == '':
            units = r
            continue
        out.append(dict(zip(hdr, r)))
    return hdr, units, out


def short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name.replace('vt::', '')


def num(v):
    try:
        return float(str(v).replace(',', ''))
    except ValueError:
        return float('nan')


def launch_summary():
    path = os.path.join(OUT, 'r2_launches.csv')
    if not os.path.exists(path):
        return None
    hdr, units, rows = rows_of(path)
    # long format: one row per (kernel, metric)
    per = {}
    order = []
    for r in rows:
        if 'gpu__time_duration' not in r.get('Metric Name', ''):
            continue
        ns = num(r['Metric Value'])
        unit = r.get('Metric Unit', 'ns')
        us = ns / 1e3 if unit in ('ns', 'nsecond') else (ns if unit in ('us', 'usecond') else ns * 1e3)
        k = short(r['Kernel Name'])
        order.append((k, us))
        c = per.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += us
    tot = sum(v[1] for v in per.values())
    lines = [f'# ncu --metrics gpu__time_duration.sum --clock-control none: one eager TimeSformer-B step (batch 8), {len(order)} launches, '
             f'{tot / 1e3:.2f} ms of kernel time (cold caches, serialised: shares are what counts)', '# kernel | launches | total us | share %']
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'{k:70s} {n:5d} {us:10.1f} {100 * us / tot:6.2f}')
    with open(os.path.join(PROF, 'r2_ncu_launch_summary.txt'), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    with gzip.open(os.path.join(PROF, 'r2_ncu_launches.csv.gz'), 'wt') as fh:
        fh.write(open(path).read())
    return per, tot


def find(hdr, *subs):
    for h in hdr:
        if all(s in h for s in subs):
            return h
    return None


def alg_bytes(c):
    esz = 4 if c['epi'] == 'f32' else 2
    b = c['M'] * c['K'] * 2 + c['N'] * c['K'] * 2 + c['M'] * c['N'] * esz
    if c['aux']:
        b += c['M'] * c['N'] * (4 if c['epi'] == 'f32' else 2)
    if c['epi'] == 'gelu':
        b += c['M'] * c['N'] * 2
    return b


try:
    BURST_TF = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops']
except Exception:
    BURST_TF = 1682.6


def full_summary():
    calls = []
    cj = os.path.join(OUT, 'r2_gemm_calls.json')
    if os.path.exists(cj):
        calls = json.load(open(cj))
    lines = ['# ncu --set full --clock-control none --import-source on: hot kernels of one eager TimeSformer-B step (batch 8)',
             '# fwd = first layers of the forward pass, bwd = backward of the last layers',
             '# pass | kernel | grid | duration us | DRAM read MB | DRAM write MB | sm__pipe_tensor_cycles_active % (blind to tcgen05) | DRAM throughput % | '
             'achieved occupancy % | regs | (GEMM: M N K epilogue, algorithmic MB, DRAM/algorithmic)']
    gemm_rows = []
    for tag, skip in (('fwd', 0), ('bwd', None)):
        path = os.path.join(OUT, f'r2_{tag}_raw.csv')
        if not os.path.exists(path):
            continue
        hdr, units, rows = rows_of(path)
        c_name = find(hdr, 'Kernel Name')
        c_dur = find(hdr, 'gpu__time_duration.sum')
        c_rd, c_wr = find(hdr, 'dram__bytes_read.sum'), find(hdr, 'dram__bytes_write.sum')
        # ncu 2025.2's `--set full` has no counter that sees tcgen05.mma (sm__ops_path_tensor_op_hmma_* reads 0,
        # sm__pipe_tensor_cycles_active_realtime a few percent for a GEMM running at half of peak): the column is printed as
        # collected, the GEMM rows add FLOP / duration against the measured burst peak instead
        c_tp = find(hdr, 'sm__pipe_tensor', 'cycles_active', 'pct') or find(hdr, 'sm__inst_executed_pipe_tensor', 'pct')
        c_dt = find(hdr, 'gpu__dram_throughput', 'pct')
        c_occ = find(hdr, 'sm__warps_active', 'pct')
        c_reg = find(hdr, 'launch__registers_per_thread')
        c_grid = find(hdr, 'launch__grid_size') or find(hdr, 'Grid Size')
        u = dict(zip(hdr, units)) if units else {}

        def scaled(r, col, want):
            v = num(r.get(col, 'nan'))
            un = u.get(col, '')
            if want == 'us':
                return v / 1e3 if un.startswith('n') else (v * 1e3 if un.startswith('m') else v)
            if want == 'MB':
                f = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(un, 1e-6)
                return v * f
            return v

        # GEMM launches in call order: forward capture starts at call 0; the backward capture is aligned from the end
        gidx = [i for i, r in enumerate(rows) if 'gemm' in r[c_name] and 'tcgen05' in r[c_name]]
        n_g = len(gidx)
        for i, r in enumerate(rows):
            name = short(r[c_name])
            dur = scaled(r, c_dur, 'us')
            rd, wr = scaled(r, c_rd, 'MB'), scaled(r, c_wr, 'MB')
            extra = ''
            if i in gidx and calls:
                pos = gidx.index(i)
                ci = pos if tag == 'fwd' else None
                if tag == 'bwd':
                    ci = None          # resolved below through the -s offset: not recoverable from the csv alone
                if ci is not None and ci < len(calls):
                    c = calls[ci]
                    ab = alg_bytes(c) / 1e6
                    tf = 2.0 * c['M'] * c['N'] * c['K'] / (dur * 1e-6) / 1e12
                    extra = (f" | M={c['M']} N={c['N']} K={c['K']} {c['epi']}{'+aux' if c['aux'] else ''} alg {ab:.1f} MB x{(rd + wr) / ab:.2f}"
                             f" | {tf:.0f} TFLOP/s = {100 * tf / BURST_TF:.0f}% of burst peak")
                    gemm_rows.append(dict(shape=[c['M'], c['N'], c['K']], epi=c['epi'], aux=c['aux'], dram_mb=rd + wr, alg_mb=ab,
                                          us=dur, tensor_pct=100 * tf / BURST_TF))
            lines.append(f"{tag} {name:52s} {r.get(c_grid, '?'):>7s} {dur:8.1f} {rd:8.1f} {wr:8.1f} {num(r.get(c_tp, 'nan')):6.1f} "
                         f"{num(r.get(c_dt, 'nan')):6.1f} {num(r.get(c_occ, 'nan')):6.1f} {r.get(c_reg, '?'):>4s}{extra}")
    with open(os.path.join(PROF, 'r2_ncu_full_summary.txt'), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    if gemm_rows:
        tot_us = sum(g['us'] for g in gemm_rows)
        traffic = {
            'kernel': 'gemm_tcgen05_kernel / gemm2_tcgen05_kernel',
            'dram_bytes_per_launch': 1e6 * sum(g['dram_mb'] for g in gemm_rows) / len(gemm_rows),
            'algorithmic_bytes_per_launch': 1e6 * sum(g['alg_mb'] for g in gemm_rows) / len(gemm_rows),
            'launches_averaged': len(gemm_rows),
            'time_weighted_pct_of_burst_bf16_peak': sum(g['us'] * g['tensor_pct'] for g in gemm_rows) / tot_us,
            'per_shape': gemm_rows[:14],
            'source': 'profiles/r2_ncu_full_summary.txt (ncu --set full, forward GEMM launches of the first layers of one step; '
                      'algorithmic bytes = operands + output + epilogue addend of the logged call)'}
        with open(os.path.join(PROF, 'gemm_traffic.json'), 'w') as fh:
            json.dump(traffic, fh, indent=1)
    return len(lines)


if __name__ == '__main__':
    print('launch summary:', 'ok' if launch_summary() else 'no r2_launches.csv')
    print('full summary lines:', full_summary())
