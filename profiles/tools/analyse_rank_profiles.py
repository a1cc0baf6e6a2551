"""Summary of the per-rank, per-launch CSVs written by bench.py with SPLLT_BENCH_PROFILE_CSV
(un-graphed, event-timed pass; every rank).  usage: analyse_rank_profiles.py <prefix> <nranks>"""
import collections
import csv
import sys

prefix, n = sys.argv[1], int(sys.argv[2])
KIND = {0: "panel", 1: "tile_s", 2: "tile_l", 3: "push", 4: "wait"}
tot = []
for r in range(n):
    rows = list(csv.DictReader(open("%s.rank%d.csv" % (prefix, r))))
    by = collections.defaultdict(float)
    cnt = collections.Counter()
    for x in rows:
        k, tag, ms = int(x["kind"]), int(x["tag"]), float(x["ms"])
        phase = "top" if tag in (5, 7, 8, 9) or (k in (3, 4)) else None
        key = (KIND[k], tag)
        by[key] += ms
        cnt[key] += 1
    tot.append(by)
    s = sum(by.values())
    print("rank %d: total %.2f ms  " % (r, s) + "  ".join("%s/t%d %.2f(%d)" % (k[0], k[1], v, cnt[k]) for k, v in sorted(by.items())))
# per step (rank 0's view): chain time of the owner, wait of the others
rows0 = [list(csv.DictReader(open("%s.rank%d.csv" % (prefix, r)))) for r in range(n)]
steps = collections.defaultdict(lambda: collections.defaultdict(float))
for r in range(n):
    for x in rows0[r]:
        k, tag, ms, d = int(x["kind"]), int(x["tag"]), float(x["ms"]), int(x["depth"])
        if k in (3, 4) or tag in (5, 7):
            steps[d]["%s" % ("push" if k == 3 else "wait" if k == 4 else "upd")] += ms
        elif tag in (0, 4) and any(int(y["kind"]) == 3 and int(y["depth"]) == d for y in rows0[r][:0]):
            pass
print("steps: %d" % len(steps))
