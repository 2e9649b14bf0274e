"""Static code size per source line of one kernel of an object file (nvdisasm --print-line-info).
usage: sass_lines.py file.o kernel_substring [top]"""
import sys, re, collections, subprocess, os, tempfile, glob
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, capture_output=True)
txt = subprocess.run(['nvdisasm', '--print-line-info', glob.glob(d + '/*.cubin')[0]], capture_output=True, text=True).stdout
cur = None; cnt = collections.Counter(); active = False; ops = collections.Counter()
for ln in txt.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m: active = pat in m.group(1); continue
    if ln.startswith('.section') or ln.startswith('\t.section'): active = active and ('.text.' in ln and pat in ln)
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m and cur: cnt[cur] += 1; ops[m.group(1)] += 1
tot = sum(cnt.values()); print('instructions', tot, '=', tot * 16 // 1024, 'KB')
print(ops.most_common(14))
for (f, l), n in cnt.most_common(top): print(n, f, l)
