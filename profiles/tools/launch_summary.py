import csv, collections, sys
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if 'Kernel Name' in r: hdr = r
            continue
        if len(r) < len(hdr): continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum': continue
        name = d['Kernel Name'][:50] + '|' + d.get('Grid Size', '')
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        if u == 'ns': v /= 1e3
        elif u == 'ms': v *= 1e3
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f, "%.2f ms" % (tot / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"  {k:70s} n={v[0]:4d} tot={v[1]/1e3:9.2f}ms avg={v[1]/v[0]:9.1f}us {100*v[1]/tot:5.1f}%")
