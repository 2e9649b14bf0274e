"""Turn the raw ncu artefacts of a gpurun call (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> <launches.csv | -> <name=report.ncu-rep> ..."""
import csv, io, json, os, subprocess, sys, collections, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
out_dir = os.path.join(ROOT, 'profiles')
UNIT = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1., 'Tbyte': 1e12}

rows = [] if launches == '-' else [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0].replace('void ', '')
    per.setdefault(name, []).append(float(r[-1]))
tot = sum(sum(v) for v in per.values())
if rows:
  with open(os.path.join(out_dir, tag + '_launches_summary.md'), 'w') as f:
      f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none): `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras`\n')
      f.write('per-launch times are cold-cache and serialised; what must agree with bench.py is each kernel\'s SHARE\n\n')
      f.write('| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|\n')
      for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
          f.write('| {} | {} | {:.3f} | {:.4f} | {:.2%} |\n'.format(k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6, sum(v) / tot))
  shutil.copy(launches, os.path.join(out_dir, tag + '_launches.csv'))

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__sass_inst_executed_op_global_ld.sum', 'smsp__sass_inst_executed_op_shared_ld.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sm__icc_request_hit_rate.pct', 'sm__icc_requests.sum']
traffic = {}
for spec in sys.argv[3:]:
    name, rep = spec.split('=')
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode()
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rr[0], rr[1], rr[-1]
    kname = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else name
    with open(os.path.join(out_dir, '{}_{}_ncu_summary.md'.format(tag, name)), 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on: {}\n\n| metric | value | unit |\n|---|---|---|\n'.format(kname))
        for w in WANT:
            if w in hdr:
                f.write('| {} | {} | {} |\n'.format(w, vals[hdr.index(w)], units[hdr.index(w)]))
        for w in ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
                  'SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg',
                  'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
                  'sm__cycles_elapsed.max', 'smsp__cycles_active.avg'):
            if w in hdr:
                f.write('| {} | {} | {} |\n'.format(w, vals[hdr.index(w)], units[hdr.index(w)]))
        f.write('\nwarp issue stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active.ratio):\n\n')
        st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(vals[i]))
              for i, h in enumerate(hdr) if 'average_warps_issue_stalled' in h and 'not_issued' not in h]
        for k, v in sorted(st, key=lambda kv: -kv[1]):
            if v > 0.005:
                f.write('- {}: {:.3f}\n'.format(k, v))
    try:
        ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        traffic[name] = dict(kernel=kname, dram_bytes_per_launch=float(vals[ir].replace(',', '')) * UNIT[units[ir]] + float(vals[iw].replace(',', '')) * UNIT[units[iw]])
    except Exception as e:
        print('traffic', name, e)
print(json.dumps(traffic))
json.dump(traffic, open(os.path.join(out_dir, tag + '_traffic.json'), 'w'), indent=1)
