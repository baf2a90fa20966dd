"""Pretty-print the SBTRACE lines of a CAVIT_SB_TRACE build (control warp + elementwise warp 0 of CTA 0)."""
import sys
ev = {0: [], 1: []}
for l in open(sys.argv[1]):
    if l.startswith('SBTRACE'):
        _, r, i, t, tag = l.split()
        ev[int(r)].append((int(t), int(tag)))
allv = sorted((t, 'C' if r == 0 else 'E', tag) for r in ev for (t, tag) in ev[r] if 0 < t < 10 ** 9)
names = {0: 'ctl: top', 1: 'ctl: sfree ok', 2: 'ctl: S issued, wait p', 3: 'ctl: p ok', 4: 'ctl: D issued', 5: 'ctl: D done',
         6: 'ctl: loads issued', 10: 'EW: top', 11: 'EW: bar_s ok', 12: 'EW: math done', 13: 'EW: bar_d ok',
         14: 'EW: drain done', 15: 'EW: stored+arrived'}
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 300, int(sys.argv[3]) if len(sys.argv) > 3 else 420
start = None
for t, who, tag in allv[lo:hi]:
    if start is None:
        start = t
    print(f"{t - start:7d} {who} {names[tag]}")
