"""Per-source-line warp-stall samples of an .ncu-rep taken with --import-source on (kernels built with -lineinfo); read here, no GPU needed.
    python profiles/source_hotspots.py gpurun_out/prof.ncu-rep [top=40] > profiles/rN_name_ncu_source_hotspots.txt"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr and r[0] not in ('', 'Line No') and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            smp = int(d['# Samples']); inst = int(d['Instructions Executed'])
        except (KeyError, ValueError):
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and not k.endswith('(Not Issued)') and v.isdigit() and int(v) > 0}
        lines.append((smp, inst, fname, int(r[0]), r[1].strip()[:110], stalls))
tot = sum(l[0] for l in lines)
print('total samples', tot)
agg = {}
for l in lines:
    for k, v in l[5].items():
        agg[k] = agg.get(k, 0) + v
print('stall reasons:', ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
for smp, inst, f, ln, src, st in sorted(lines, reverse=True)[:top]:
    s3 = ', '.join('%s %d' % (k, v) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print('%5.2f%%  %9d inst  %s:%d  %s   [%s]' % (100.0 * smp / tot, inst, f, ln, src, s3))
