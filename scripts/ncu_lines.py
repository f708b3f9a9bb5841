"""Dev helper: aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, collections, sys
path, srcfile = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
hdr = rows[2]
idx = {}
for i, h in enumerate(hdr):
    idx.setdefault(h, i)
def num(x):
    try: return int(float(x))
    except Exception: return 0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(collections.Counter); inst = collections.Counter(); tot = collections.Counter()
for r in rows[3:]:
    if len(r) < len(hdr): continue
    line = r[0]
    agg[line]['samples'] += num(r[idx['# Samples']])
    inst[line] += num(r[idx['Instructions Executed']])
    for c in stall_cols:
        v = num(r[idx[c]]); agg[line][c] += v; tot[c] += v
total = sum(a['samples'] for a in agg.values())
print('total samples', total, 'total warp-inst', sum(inst.values()))
print({k: round(v / total, 3) for k, v in tot.most_common(10)})
src = open(srcfile).read().split('\n')
for line, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:topn]:
    top = [(k[6:], v) for k, v in a.most_common(4) if k != 'samples'][:3]
    l = int(line) if line.isdigit() else 0
    print(line, '%.1f%%' % (100 * a['samples'] / total), inst[line], top, '|', src[l - 1].strip()[:80] if l else '')
