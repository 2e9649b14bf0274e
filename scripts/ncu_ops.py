"""Opcode histogram (executed warp instructions, stall samples) of an .ncu-rep, SASS view (no double counting of inlined lines).
usage: ncu_ops.py report.ncu-rep [units]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); iss = hdr.index('# Samples')
ops = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= ia: continue
    try: n = int(r[ia]); s = int(r[iss])
    except ValueError: continue
    tk = r[isrc].split()
    op = (tk[1] if tk[0].startswith('@') else tk[0])
    op = op.split('.')[0] + ('.128' if '.128' in op else '')
    ops[op] += n; samp[op] += s; tot += n
ts = sum(samp.values()) or 1
print('warp instructions', tot, ('per unit %.1f' % (tot / units)) if units else '')
for op, n in ops.most_common(40):
    print('%-10s instr %5.1f%% %s samples %5.1f%%' % (op, 100. * n / tot, ('%8.1f/unit' % (n / units)) if units else '', 100. * samp[op] / ts))
