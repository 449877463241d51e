"""Aggregate an .ncu-rep's stall samples / executed instructions by CUDA source line (read here, no GPU):
python tools/ncu_lines.py file.ncu-rep [top]"""
import csv
import subprocess
import sys


def _i(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, top=30):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, curfile, hdr, out = None, '', None, {}
    for r in rows:
        if not r:
            continue
        if r[0] in ('Function Name', 'Kernel Name'):
            cur = r[1]
            out.setdefault(cur, [])
        elif r[0] in ('File Path', 'File Name'):
            curfile = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = r
        elif cur and hdr and len(r) >= 8 and r[0].isdigit():
            out[cur].append((curfile, r))
    for fn, lst in out.items():
        si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
        tot = sum(_i(r[si]) for _, r in lst) or 1
        toti = sum(_i(r[ii]) for _, r in lst) or 1
        print('==', fn[:100], ' samples', tot, ' warp-instr', toti)
        for f, r in sorted(lst, key=lambda x: -_i(x[1][si]))[:top]:
            print('%5.1f%% smp %5.1f%% ins  %s:%s  %s' % (100 * _i(r[si]) / tot, 100 * _i(r[ii]) / toti,
                                                      f, r[0], r[1].strip()[:105]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
