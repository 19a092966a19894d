"""Summarise the source page of an ncu report: top SASS instructions by stall samples, grouped in address windows."""
import csv, sys
fn = sys.argv[1]
rows = list(csv.reader(open(fn)))
hdr = rows[1]
ia, isrc, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[isamp])
    except ValueError: continue
    data.append((r[ia], r[isrc].strip(), n, {hdr[i]: int(r[i] or 0) for i in stall_cols}))
total = sum(d[2] for d in data)
print('total samples', total, 'instructions', len(data))
# windows of 64 instructions
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for w0 in range(0, len(data), W):
    chunk = data[w0:w0 + W]
    n = sum(d[2] for d in chunk)
    if n < 0.01 * total: continue
    agg = {}
    for d in chunk:
        for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
    ops = {}
    for d in chunk:
        op = d[1].split()[0] if not d[1].startswith('@') else d[1].split()[1]
        ops[op] = ops.get(op, 0) + 1
    topops = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print('[{:5d}-{:5d}] {:5.1f}%  {}  | {}'.format(w0, w0 + len(chunk), 100. * n / total, ' '.join('{}={:.0f}%'.format(k[6:], 100. * v / max(n, 1)) for k, v in top), topops))
