#!/usr/bin/env python3
"""Static SASS histogram of a line range of an `nvdisasm -g -c` listing, grouped by source line (inlining-aware:
the innermost '//## File' annotation wins).  usage: sass_by_line.py listing.txt first_line last_line"""
import collections
import re
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cur = "?"
by_line = collections.Counter()
by_op = collections.Counter()
ops_line = collections.defaultdict(collections.Counter)
for i, l in enumerate(open(path), 1):
    m = re.search(r'//## File "[^"]*/([^"/]*)", line (\d+)', l)
    if m:
        cur = "%s:%s" % (m.group(1), m.group(2))
        continue
    if i < lo or i > hi:
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
    if m:
        op = m.group(2)
        by_line[cur] += 1
        by_op[op] += 1
        ops_line[cur][op] += 1
tot = sum(by_op.values())
print("total", tot)
print("by op:", ", ".join("%s %d" % kv for kv in by_op.most_common(30)))
for k, v in by_line.most_common(45):
    print("%5d %-28s %s" % (v, k, " ".join("%s:%d" % kv for kv in ops_line[k].most_common(6))))
