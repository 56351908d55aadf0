"""Join an `ncu --page source --csv` dump (per-SASS-instruction counters) with `nvdisasm -g -c` line info
to attribute executed instructions / stall samples to CUDA source lines.

usage: python profiles/attribute_lines.py src.csv kernel.sass <mangled-kernel-substring> [top_n]
"""
import collections
import csv
import re
import sys


def main():
    src_csv, sass, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    data = rows[2:]
    ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
    # walk the disassembly of the wanted function, tracking "//## File ..., line N" markers
    lines = open(sass).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip().startswith('.text.') and kname in l and l.strip().endswith(':'))
    cur = ('?', 0)
    inl = ''
    instr_lines = []
    for l in lines[start + 1:]:
        s = l.strip()
        if s.startswith('.text.') and s.endswith(':') and kname not in s:
            break
        m = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            inl = m.group(3)
            continue
        if re.match(r'/\*[0-9a-f]{4,}\*/', s):
            instr_lines.append(cur)
    if len(instr_lines) != len(data):
        print('warning: %d SASS instructions in nvdisasm vs %d in ncu' % (len(instr_lines), len(data)))
    n = min(len(instr_lines), len(data))
    by_line = collections.Counter()
    samp = collections.Counter()
    for k in range(n):
        by_line[instr_lines[k]] += int(data[k][ie])
        samp[instr_lines[k]] += int(data[k][isamp])
    tot, ts = sum(by_line.values()), max(sum(samp.values()), 1)
    print('total warp instructions %d' % tot)
    for (f, ln), c in by_line.most_common(top):
        print('%-22s %5d  %6.2f%% inst  %6.2f%% samples' % (f, ln, 100.0 * c / tot, 100.0 * samp[(f, ln)] / ts))
    by_file = collections.Counter()
    for (f, ln), c in by_line.items():
        by_file[f] += c
    print('--- by file')
    for f, c in by_file.most_common():
        print('%-30s %6.2f%%' % (f, 100.0 * c / tot))


if __name__ == '__main__':
    main()
