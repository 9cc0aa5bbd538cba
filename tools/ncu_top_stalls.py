"""Print the most-sampled SASS instructions of an `ncu --page source --csv` export with their dominant stall reasons."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:3]
    print('%5d %6d %5.1f%% %-64s %s exec=%s' % (i, int(r[ix['# Samples']]), 100 * int(r[ix['# Samples']]) / tot, r[ix['Source']].strip()[:64],
                                        ' '.join('%s:%d' % (n, v) for v, n in st if v), r[ix['Instructions Executed']]))
