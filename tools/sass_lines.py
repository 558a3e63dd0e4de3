"""Join an ncu SASS source page (per-instruction executed counts) with nvdisasm line info of the same build:
per-source-line share of executed warp instructions.
  ncu -i rep --page source --csv --print-source sass > sass.csv
  cuobjdump -xelf all lib.so; nvdisasm --print-line-info k_x.sm_100a.cubin > lines.txt
  python tools/sass_lines.py sass.csv lines.txt <kernel substring> <source dir>"""
import csv, re, collections, sys, os
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
data = [(r[isrc].strip(), int(r[ix])) for r in rows[2:] if len(r) > ix and r[ix].isdigit()]
txt = open(sys.argv[2]).read().split('\n')
start = next(i for i, l in enumerate(txt) if l.startswith('.text.') and sys.argv[3] in l)
cur, seq = None, []
for l in txt[start + 1:]:
    if l.startswith('.text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(1)))
n = min(len(seq), len(data))
agg = collections.Counter()
for k in range(n):
    agg[seq[k][0]] += data[k][1]
tot = sum(agg.values())
print(f"# {tot} warp instructions over {n} SASS instructions")
src = {}
for (key, nn) in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 30):
    line = ''
    if key:
        f = os.path.join(sys.argv[4], key[0])
        if key[0] not in src and os.path.exists(f):
            src[key[0]] = open(f).read().split('\n')
        if key[0] in src:
            line = src[key[0]][key[1] - 1].strip()[:100]
    print(f"{100 * nn / tot:5.1f}% {key} {line}")
