"""Read an .ncu-rep here (no GPU): stall samples and executed instructions aggregated per source line and per opcode.
usage: ncu_lines.py report.ncu-rep [top_n] [units]   (units: divide instruction counts, e.g. the number of 8-chain leapfrogs)"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; cur = None; lines = {}; ops = collections.Counter(); osamp = collections.Counter(); kernels = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hdr = r; continue
    if not r or hdr is None: continue
    extra = len(r) - len(hdr)
    if r[0] != '':
        try: ln = int(r[0])
        except ValueError: continue
        try: s = int(r[6 + extra]); n = int(r[7 + extra])
        except ValueError: continue
        k = (cur, ln)
        o = lines.get(k, (0, 0, ''))
        lines[k] = (o[0] + s, o[1] + n, ','.join(r[1:2 + extra])[:100])
    elif len(r) > 7 and r[2].startswith('0x'):
        try: s = int(r[6]); n = int(r[7])
        except ValueError: continue
        tk = r[3].split()
        op = (tk[1] if tk[0].startswith('@') else tk[0]).split('.')[0]
        ops[op] += n; osamp[op] += s
ts = sum(v[0] for v in lines.values()) or 1; tn = sum(v[1] for v in lines.values()) or 1
print('total samples', ts, 'warp instructions', tn, ('per unit %.1f' % (tn / units)) if units else '')
print('--- opcodes')
for op, n in ops.most_common(30):
    print('%-10s instr %5.1f%% %s samples %5.1f%%' % (op, 100. * n / tn, ('%8.1f/unit' % (n / units)) if units else '', 100. * osamp[op] / ts))
print('--- lines by stall samples')
for (f, ln), (s, n, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%s:%d samp %5.2f%% instr %5.2f%%  %s' % (f, ln, 100. * s / ts, 100. * n / tn, src))
