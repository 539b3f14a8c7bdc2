#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals from an .ncu-rep (source page)."""
import csv, subprocess, sys
def main(path, top=40):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source=cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file = ''
    hdr = None
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
        if len(r) == 2: continue
        if r and r[0] == 'Line No': hdr = r; continue
        if hdr and r and r[0] != '':
            try:
                lines.append((cur_file, int(r[0]), r[1].strip()[:100], int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')])))
            except ValueError:
                pass
    tot_i = sum(l[3] for l in lines); tot_s = sum(l[4] for l in lines)
    print('total inst', tot_i, 'samples', tot_s)
    print('--- by instructions')
    for l in sorted(lines, key=lambda x: -x[3])[:top]:
        print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100*l[3]/tot_i, 100*l[4]/max(tot_s,1), l[0], l[1], l[2]))
    print('--- by stall samples')
    for l in sorted(lines, key=lambda x: -x[4])[:top]:
        print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100*l[3]/tot_i, 100*l[4]/max(tot_s,1), l[0], l[1], l[2]))
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
