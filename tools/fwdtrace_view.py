"""Pretty-print the FWDTRACE lines of a CAVIT_FWD_TRACE build of the attention forward kernel
(NVCC_EXTRA=-DCAVIT_FWD_TRACE python cross-attention-vit_b200/build.py --force; python tools/one_attn.py 2> trace.log):
clock64 time stamps of CTA 0's TMA producer, MMA issuer and the first softmax warp of each slot.
    python tools/fwdtrace_view.py trace.log [first_item] [n_items]"""
import sys

NAMES = {0: 'TMA: item top (wait qempty)', 1: 'TMA: qempty ok -> load Q', 2: 'TMA: kvempty ok -> load KV',
         10: 'MMA: sfree0 ok -> S0', 11: 'MMA: sfree1 ok -> S1', 12: 'MMA: pfull0 ok -> PV0', 13: 'MMA: pfull1 ok -> PV1',
         14: 'MMA: item top', 15: 'MMA: next qfull ok', 16: 'MMA: next kvfull ok',
         20: 'wait sfull', 21: 'sfull ok', 22: 'pass1 done', 23: 'max exchanged', 24: 'ofull(prev) ok', 25: 'pass2 done',
         26: 'wait last O', 27: 'last O ok', 28: 'last O folded', 29: 'O staged', 30: 'staging barrier', 31: 'O rows stored'}
WHO = {0: 'TMA ', 1: 'MMA ', 2: 'SM0 ', 3: 'SM1 '}


def main():
    ev = []
    for line in open(sys.argv[1]):
        if line.startswith('FWDTRACE'):
            _, r, _i, t, tag = line.split()
            if int(t) < 10 ** 12:
                ev.append((int(t), int(r), int(tag)))
    ev.sort()
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    tops = [t for t, r, tag in ev if r == 1 and tag == 14]
    print("item periods (cycles):", [b - a for a, b in zip(tops[first:first + 12], tops[first + 1:first + 13])])
    cnt, t0 = 0, None
    for t, r, tag in ev:
        if r == 1 and tag == 14:
            cnt += 1
        if first <= cnt < first + n:
            t0 = t if t0 is None else t0
            print(f"{t - t0:7d} {WHO[r]} {NAMES.get(tag, tag)}")


if __name__ == "__main__":
    main()
