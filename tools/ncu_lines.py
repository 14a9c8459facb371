"""Per-source-line hot spots of one kernel of an .ncu-rep captured with --import-source on (developer tool).

    python tools/ncu_lines.py REP KERNEL_SUBSTR [top]

Joins the SASS page of the report (stall samples, executed instructions per instruction) with the line table
nvdisasm prints for the same cubin of depth_correction_b200/libdcb200.so, in instruction order.
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(kernel):
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'depth_correction_b200', 'libdcb200.so')
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    for cub in sorted(glob.glob(os.path.join(tmp, '*.cubin'))):
        if 'libdcb200' in os.path.basename(cub):
            continue
        out = subprocess.run(['nvdisasm', '--print-line-info', '-c', cub], capture_output=True, text=True).stdout
        table, cur, inside = [], None, False
        for ln in out.splitlines():
            if ln.startswith('.text.'):
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*);', ln)
            if m:
                table.append((int(m.group(1), 16), cur, m.group(2).strip()))
        if table:
            return table
    raise SystemExit('kernel not found in the cubins')


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    # KERNEL_SUBSTR may be a mangled name (to pick one template instance in the cubin); ncu filters on the plain name
    plain = re.sub(r'^_Z\d+', '', kernel)
    plain = re.sub(r'I[A-Z].*$', '', plain)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + plain], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    ci = {n: hdr.index(n) for n in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed')}
    body = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith('0x'):
            break
        body.append(r)
    table = line_table(kernel)
    base = int(body[0][ci['Address']], 16)
    by_off = {off: (loc, txt) for off, loc, txt in table}
    agg = defaultdict(lambda: [0, 0, 0])
    tot_s = tot_i = 0
    for r in body:
        off = int(r[ci['Address']], 16) - base
        loc = by_off.get(off, (None, ''))[0]
        s, i, t = int(r[ci['# Samples']]), int(r[ci['Instructions Executed']]), int(r[ci['Thread Instructions Executed']])
        a = agg[loc]
        a[0] += s; a[1] += i; a[2] += t
        tot_s += s; tot_i += i
    src = {}
    print('total samples %d, warp instructions %d' % (tot_s, tot_i))
    print('%-22s %8s %6s %12s %6s %6s  %s' % ('file:line', 'samples', '%', 'warp-inst', '%', 'thr/w', 'source'))
    for loc, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ''
        if loc:
            f = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'depth_correction_b200', 'csrc', loc[0])
            if f not in src and os.path.exists(f):
                src[f] = open(f).read().splitlines()
            if f in src and loc[1] <= len(src[f]):
                text = src[f][loc[1] - 1].strip()[:90]
        print('%-22s %8d %6.2f %12d %6.2f %6.1f  %s' % ('%s:%d' % loc if loc else '?', s, 100.0 * s / max(tot_s, 1), i,
                                                      100.0 * i / max(tot_i, 1), t / max(i, 1), text))


if __name__ == '__main__':
    main()
